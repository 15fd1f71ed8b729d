"""Build container only (`reference` marker): the fused path's drop-in classes side by side with the live, unmodified
reference on ALL 25 pymgrid25 scenarios -- Microgrid.run dicts, rewards, infos, get_log() frames, state_series(), reset(), and
DiscreteMicrogridEnv steps and logs -- with the engine replaced by the oracle-backed stand-in (tests/oracle_engine.py): what
is compared here is the package's Python host layer; the kernels are compared with the oracle in the GPU suites."""
import warnings

import numpy as np
import pytest

from tests.oracle_engine import install

pytestmark = pytest.mark.reference


@pytest.fixture(autouse=True)
def _oracle_backed_engine(monkeypatch):
    install(monkeypatch)


def same_nested(a, b, where):
    assert list(a.keys()) == list(b.keys()), (where, list(a.keys()), list(b.keys()))
    for k in a:
        assert len(a[k]) == len(b[k]), (where, k)
        for x, y in zip(a[k], b[k]):
            if isinstance(x, dict):
                assert dict(x) == dict(y), (where, k, x, y)
            else:
                assert np.array_equal(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)), (where, k)


def same_frame(a, b, where):
    assert [tuple(c) for c in a.columns] == [tuple(c) for c in b.columns], where
    assert np.array_equal(a.to_numpy(dtype=np.float64), b.to_numpy(dtype=np.float64), equal_nan=True), where
    assert np.array_equal(np.array(a.index), np.array(b.index)), where


@pytest.mark.parametrize("n", range(25))
def test_microgrid_surface_against_the_live_reference(n):
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    from pymgrid_b200.microgrid import Microgrid
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, ours = pymgrid.Microgrid.from_scenario(n), Microgrid.from_scenario(n)
    assert repr(ref) == repr(ours)
    assert len(ref) == len(ours) and ref.initial_step == ours.initial_step and ref.final_step == ours.final_step
    assert ref.get_empty_action() == ours.get_empty_action()
    np.random.seed(n)
    for k in range(12):
        state = np.random.get_state()
        a = ref.sample_action()
        np.random.set_state(state)
        b = ours.sample_action()                       # same draws from numpy's global generator
        same_nested(a, b, (n, k, "sample_action"))
        o1, r1, d1, i1 = ref.run(a, normalized=True)
        o2, r2, d2, i2 = ours.run(b, normalized=True)
        assert r1 == r2 and d1 == d2, (n, k)
        same_nested(o1, o2, (n, k, "obs"))
        same_nested(i1, i2, (n, k, "info"))
        assert ref.current_step == ours.current_step
    same_frame(ref.get_log(), ours.get_log(), (n, "log"))
    s1, s2 = ref.state_series(), ours.state_series()
    assert [tuple(map(str, i)) for i in s1.index] == [tuple(map(str, i)) for i in s2.index]
    assert np.array_equal(s1.to_numpy(dtype=np.float64), s2.to_numpy(dtype=np.float64))
    sd1, sd2 = ref.state_dict(), ours.state_dict()
    assert {k: [dict(d) for d in v] for k, v in sd1.items()} == {k: [dict(d) for d in v] for k, v in sd2.items()}
    c1, c2 = ref.get_cost_info(), ours.get_cost_info()
    assert {k: [dict(d) for d in v] for k, v in c1.items()} == {k: [dict(d) for d in v] for k, v in c2.items()}
    r1, r2 = ref.reset(), ours.reset()
    same_nested({k: v for k, v in r1.items() if k not in ("balance", "other")}, {k: v for k, v in r2.items() if k not in ("balance", "other")},
                (n, "reset"))
    assert list(r1.keys()) == list(r2.keys()) and len(ours.get_log()) == 0


@pytest.mark.parametrize("n", range(25))
def test_discrete_env_against_the_live_reference(n):
    from oracle.ref_loader import load_reference
    load_reference()
    from pymgrid.envs import DiscreteMicrogridEnv as RefEnv
    from pymgrid_b200.envs import DiscreteMicrogridEnv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, ours = RefEnv.from_scenario(n), DiscreteMicrogridEnv.from_scenario(n)
    assert ref.action_space.n == ours.action_space.n and ref.observation_space.shape == ours.observation_space.shape
    rows = lambda pls: [[(el.module, el.module_actions, el.action, el.marginal_cost) for el in pl] for pl in pls]      # noqa: E731
    assert rows(ref.actions_list) == rows(ours.actions_list)
    assert np.array_equal(ref.reset(), ours.reset())
    rng = np.random.default_rng(n)
    for k in range(10):
        a = int(rng.integers(0, ref.action_space.n))
        o1, r1, d1, i1 = ref.step(a)
        o2, r2, d2, i2 = ours.step(a)
        assert r1 == r2 and d1 == d2 and np.array_equal(o1, o2), (n, k)
        same_nested(i1, i2, (n, k, "info"))
    same_frame(ref.log, ours.log, (n, "env log"))
