def init_notebook_mode(*args, **kwargs):
    pass


def iplot(*args, **kwargs):
    raise NotImplementedError("plotting is not available in the oracle shim")
