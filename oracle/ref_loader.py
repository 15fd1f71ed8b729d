"""Import the UNMODIFIED upstream reference (build container only) -- test infrastructure.

`load_reference()` puts `oracle/ref_shim` (gym / plotting stand-ins) and `/root/reference/src` on
sys.path, restores `numpy.product` (removed in numpy 2, used by the reference at
modules/base/base_module.py:145) and returns the imported `pymgrid` package.

/root/reference does not exist on the GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py`
may call this.  It is used by `tests/golden/make_golden.py` (fixture generation) and by the CPU
tests that pin the C oracle against the live reference when the reference tree is present.
"""
import os
import sys
import warnings

REFERENCE_SRC = os.environ.get("PYMGRID_REFERENCE_SRC", "/root/reference/src")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "pymgrid"))


def load_reference():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod
    for p in (_SHIM, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pymgrid
    return pymgrid
