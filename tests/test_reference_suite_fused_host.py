"""tests/reference_suite_fused.py (what the reference's TestMicrogrid / TestTrajectory / TestRBC pin on its fixture grid) on the
CPU: the engine is replaced by the oracle-backed stand-in of tests/oracle_engine.py, so what runs here is the package's
Python host layer.  The GPU run of the same functions is in tests/test_zz_gpu_dropin_more.py."""
import pytest

from tests.oracle_engine import install
from tests.reference_suite_fused import CHECKS


@pytest.fixture(autouse=True)
def _oracle_backed_engine(monkeypatch):
    install(monkeypatch)


for _fn in CHECKS:
    globals()[_fn.__name__] = _fn
del _fn
