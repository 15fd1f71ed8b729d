"""The reference's TestMicrogridLoadPV family (tests/microgrid/test_microgrid.py:188-421) against pymgrid_b200.Microgrid
on the CPU: the composed path's C source built for the host (tests/hostsim).  The GPU run of the same classes is in
tests/test_zz_gpu_compose.py."""
import ctypes

from tests import hostsim
from tests.reference_suite_compose import make_suite

for _cls in make_suite(ctypes.CDLL(hostsim.build())):
    globals()[_cls.__name__] = _cls
del _cls
