"""The facts the reference's TestMicrogridLoadPV family pins (tests/microgrid/test_microgrid.py:188-455), checked against
pymgrid_b200.Microgrid on the CPU: the composed path's C source built for the host (tests/hostsim).  The GPU run of the
same checks is in tests/test_zz_gpu_compose.py."""
import ctypes
import functools

from tests import hostsim
from tests.reference_suite_compose import checks


@functools.lru_cache(maxsize=1)
def _host_build():
    return ctypes.CDLL(hostsim.build())


for _fn in checks(_host_build):
    globals()[_fn.__name__] = _fn
del _fn
