"""Canary for the gym / plotting stand-ins (build container only): run the reference's OWN unit tests, unmodified,
through oracle/ref_shim.  Expected: 368 passed, 9 568 subtests passed (tests/microgrid, tests/envs,
tests/control/test_rbc.py; the MPC / data-generation tests need cvxpy / statsmodels, which are absent).

    python oracle/run_reference_tests.py
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYMGRID_REFERENCE_ROOT", "/root/reference")

if __name__ == "__main__":
    sys.path[:0] = [os.path.join(HERE, "ref_shim"), os.path.join(REF, "src")]
    import numpy as np
    if not hasattr(np, "product"):
        np.product = np.prod
    import pytest
    with tempfile.TemporaryDirectory() as tmp:      # the reference tree is read-only: keep pytest's files elsewhere
        os.chdir(tmp)
        sys.exit(pytest.main([os.path.join(REF, "tests", "microgrid"), os.path.join(REF, "tests", "envs"),
                              os.path.join(REF, "tests", "control", "test_rbc.py"), "-q", "-p", "no:cacheprovider",
                              "--rootdir", tmp, "-W", "ignore"]))
