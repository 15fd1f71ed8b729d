def get_ipython():
    return None
