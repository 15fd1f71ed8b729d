"""tests/reference_suite_fused.py (the reference's TestMicrogrid / TestTrajectory / TestRBC on its fixture grid) on the CPU:
the engine is replaced by the oracle-backed stand-in of tests/oracle_engine.py, so what runs here is the package's Python
host layer.  The GPU run of the same classes is in tests/test_zz_gpu_dropin_more.py."""
import pytest

from tests.oracle_engine import install
from tests.reference_suite_fused import SUITES


@pytest.fixture(autouse=True)
def _oracle_backed_engine(monkeypatch):
    install(monkeypatch)


for _cls in SUITES:
    globals()[_cls.__name__] = _cls
del _cls
