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
    # normalized=True (what BaseMicrogridEnv._get_obs reads, envs/base/base.py:211-218): every module through its own space
    n1, n2 = ref.state_dict(normalized=True), ours.state_dict(normalized=True)
    assert {k: [{f: float(x) for f, x in d.items()} for d in v] for k, v in n1.items()} == \
           {k: [{f: float(x) for f, x in d.items()} for d in v] for k, v in n2.items()}, (n, "state_dict(normalized=True)")
    t1, t2 = ref.state_series(normalized=True), ours.state_series(normalized=True)
    assert [tuple(map(str, i)) for i in t1.index] == [tuple(map(str, i)) for i in t2.index]
    assert np.array_equal(t1.to_numpy(dtype=np.float64), t2.to_numpy(dtype=np.float64)), (n, "state_series(normalized=True)")
    c1, c2 = ref.get_cost_info(), ours.get_cost_info()
    assert {k: [dict(d) for d in v] for k, v in c1.items()} == {k: [dict(d) for d in v] for k, v in c2.items()}
    r1, r2 = ref.reset(), ours.reset()
    same_nested({k: v for k, v in r1.items() if k not in ("balance", "other")}, {k: v for k, v in r2.items() if k not in ("balance", "other")},
                (n, "reset"))
    assert list(r1.keys()) == list(r2.keys()) and len(ours.get_log()) == 0
    assert {k: [float(x) for x in v] for k, v in r1["balance"].items()} == r2["balance"] and r1["other"] == r2["other"] == {}


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


def draw_fused(rng, ns, T):
    """a random microgrid INSIDE the fused kernels' scope (one load, renewable, battery; genset / grid optional; one horizon)"""
    H = int(rng.choice([0, 1, 4, 23]))
    ts = dict(forecaster="oracle", forecast_horizon=H) if H else {}
    mx = float(rng.uniform(20, 400))
    mods = [ns.LoadModule(time_series=rng.uniform(5, 200) * rng.random(T), **ts),
            (str(rng.choice(["pv", "renewable", "PV"])), ns.RenewableModule(time_series=rng.uniform(5, 200) * np.clip(rng.random(T) - 0.3, 0, None), **ts)),
            ns.BatteryModule(min_capacity=float(rng.choice([0.0, 0.2 * mx])), max_capacity=mx, max_charge=float(rng.uniform(0.05, 1.2) * mx),
                             max_discharge=float(rng.uniform(0.05, 1.2) * mx), efficiency=float(rng.choice([1.0, rng.uniform(0.5, 0.99)])),
                             battery_cost_cycle=float(rng.uniform(0, 0.5)), init_soc=float(rng.uniform(0.3, 1.0)))]
    arch = int(rng.integers(0, 4))
    if arch in (0, 1):
        gmax = float(rng.uniform(20, 200))
        mods.append(ns.GensetModule(running_min_production=float(rng.choice([0.0, 0.3 * gmax])), running_max_production=gmax,
                                    genset_cost=float(rng.uniform(0, 1)), co2_per_unit=float(rng.uniform(0, 3)),
                                    cost_per_unit_co2=float(rng.uniform(0, 0.5)), start_up_time=int(rng.integers(0, 3)),
                                    wind_down_time=int(rng.integers(0, 3)), init_start_up=bool(rng.integers(0, 2))))
    if arch in (0, 2):
        g = np.stack([rng.uniform(0.05, 0.9, T), rng.uniform(0, 0.4, T), rng.uniform(0, 0.6, T), (rng.random(T) > 0.2).astype(float)], axis=1)
        mods.append(ns.GridModule(max_import=float(rng.uniform(10, 300)), max_export=float(rng.choice([0.0, rng.uniform(10, 300)])),
                                  time_series=g[:, :int(rng.choice([3, 4]))], cost_per_unit_co2=float(rng.uniform(0, 0.5)), **ts))
    order = rng.permutation(len(mods))
    return [mods[i] for i in order], dict(loss_load_cost=float(rng.uniform(1, 20)), overgeneration_cost=float(rng.uniform(0, 5)))


@pytest.mark.parametrize("g", range(30))
def test_random_fused_microgrids_built_from_modules_against_the_live_reference(g):
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    from pymgrid.algos import RuleBasedControl as RefRBC
    import pymgrid_b200
    from pymgrid_b200 import modules as M
    from pymgrid_b200.algos import RuleBasedControl
    import ctypes
    from pymgrid_b200.compose import in_fused_scope
    from tests import hostsim
    T = 40
    hostsim.select(ctypes.CDLL(hostsim.build()))      # lists outside the fused scope run on the host build of the composed kernel
    extra = lambda mods: {}      # noqa: E731
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m1, kw = draw_fused(np.random.default_rng(600 + g), R, T)
        m2, _ = draw_fused(np.random.default_rng(600 + g), M, T)
        ref, ours = pymgrid.Microgrid(m1, **kw), pymgrid_b200.Microgrid(m2, obs_order="container", **extra(m2), **kw)
    # a list that names the grid before the battery is dispatched in that order by the reference: it takes the composed path
    assert type(ours).__name__ in ("Microgrid", "ComposedMicrogrid") and repr(ref) == repr(ours)
    rng = np.random.default_rng(g)
    for k in range(25):
        normalized = k % 3 != 2
        a = {}
        for name, lst in ref.controllable.iterdict():
            mod = lst[0]
            n = mod.action_space.shape[0]
            if normalized:
                v = rng.random(n)
            else:
                lo, hi = np.atleast_1d(mod.min_act).astype(float), np.atleast_1d(mod.max_act).astype(float)
                v = lo - 0.3 * (hi - lo) + 1.6 * (hi - lo) * rng.random(n)
                if n == 2:
                    v[0], v[1] = rng.random(), max(v[1], 0.0)
            a[name] = [v if n > 1 else float(v[0])]
        outs = []
        for runner in (ref, ours):
            try:
                outs.append(runner.run(a, normalized=normalized))
            except Exception as exc:      # noqa: BLE001 -- e.g. an over-full battery asked to absorb: AssertionError on both sides
                outs.append(type(exc).__name__)
        if isinstance(outs[0], str) or isinstance(outs[1], str):
            assert outs[0] == outs[1], (g, k, outs)
            return
        (o1, r1, d1, i1), (o2, r2, d2, i2) = outs
        assert r1 == r2 and d1 == d2, (g, k)
        same_nested(o1, o2, (g, k, "obs"))
        same_nested(i1, i2, (g, k, "info"))
    same_frame(ref.get_log(), ours.get_log(), (g, "log"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rbc1 = RefRBC(pymgrid.Microgrid(draw_fused(np.random.default_rng(600 + g), R, T)[0], **kw))
        rbc2 = RuleBasedControl(pymgrid_b200.Microgrid(m2, obs_order="container", **extra(m2), **kw))
    rows = lambda pl: [(el.module, el.module_actions, el.action, el.marginal_cost) for el in pl]      # noqa: E731
    assert rows(rbc1.priority_list) == rows(rbc2.priority_list), g
    same_frame(rbc1.run(max_steps=15), rbc2.run(max_steps=15), (g, "rbc log"))


@pytest.mark.parametrize("g", range(20))
def test_random_fused_discrete_envs_against_the_live_reference(g):
    """DiscreteMicrogridEnv(modules): action lists, flat observation (gym-sorted), steps, env log, on random fused-scope grids"""
    import ctypes
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid.modules as R
    from pymgrid.envs import DiscreteMicrogridEnv as RefEnv
    from pymgrid_b200 import modules as M
    from pymgrid_b200.compose import in_fused_scope
    from pymgrid_b200.envs import DiscreteMicrogridEnv
    from tests import hostsim
    T = 40
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m1, kw = draw_fused(np.random.default_rng(800 + g), R, T)
        m2, _ = draw_fused(np.random.default_rng(800 + g), M, T)
        if any(isinstance(m, tuple) and m[0] == "PV" for m in m2) and in_fused_scope(m2):
            pytest.skip("the CPU stand-in has no 'PV'-first observation order (the engine has: MG_OBS_GYM_SORTED_PV_FIRST)")
        hostsim.select(ctypes.CDLL(hostsim.build()))
        extra = {}
        ref, ours = RefEnv(m1, **kw), DiscreteMicrogridEnv(m2, **extra, **kw)
    rows = lambda pls: [[(el.module, el.module_actions, el.action, el.marginal_cost) for el in pl] for pl in pls]      # noqa: E731
    assert rows(ref.actions_list) == rows(ours.actions_list) and ref.action_space.n == ours.action_space.n
    assert ref.observation_space.shape == ours.observation_space.shape
    assert np.array_equal(ref.reset(), ours.reset())
    rng = np.random.default_rng(g)
    for k in range(20):
        a = int(rng.integers(0, ref.action_space.n))
        outs = []
        for env in (ref, ours):
            try:
                outs.append(env.step(a))
            except Exception as exc:      # noqa: BLE001
                outs.append(type(exc).__name__)
        if isinstance(outs[0], str) or isinstance(outs[1], str):
            assert outs[0] == outs[1], (g, k, outs)
            return
        (o1, r1, d1, i1), (o2, r2, d2, i2) = outs
        assert r1 == r2 and d1 == d2 and np.array_equal(o1, o2), (g, k)
        same_nested(i1, i2, (g, k, "info"))
    same_frame(ref.log, ours.log, (g, "env log"))


ATTRS = ("max_production", "max_consumption", "min_production", "production_marginal_cost", "absorption_marginal_cost", "marginal_cost",
         "state", "min_obs", "max_obs", "min_act", "max_act", "is_source", "is_sink", "current_step", "initial_step", "final_step",
         "soc", "current_charge", "min_soc", "max_soc", "current_status", "goal_status", "import_price", "export_price", "co2_per_kwh",
         "grid_status", "forecast_horizon", "max_capacity", "efficiency", "running_max_production", "max_import", "loss_load_cost")


@pytest.mark.parametrize("n", range(25))
def test_module_views_and_normalisation_against_the_live_reference(n):
    """every attribute the reference's in-repo callers read from a module, after a few steps, and
    Microgrid.to_normalized / from_normalized, on all 25 scenarios"""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    from pymgrid_b200.microgrid import Microgrid
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, ours = pymgrid.Microgrid.from_scenario(n), Microgrid.from_scenario(n)
    rng = np.random.default_rng(n)
    for k in range(4):
        a = {name: [rng.random(2) if name == "genset" else rng.random()] for name in ref.get_empty_action()}
        ref.run(a), ours.run(a)
        for name, lst in ref.modules.iterdict():
            mine = ours.modules[name][0]
            for attr in ATTRS:
                try:
                    want = getattr(lst[0], attr)
                except Exception:      # noqa: BLE001 -- the reference module has no such attribute
                    continue
                if want is NotImplemented:
                    continue
                got = getattr(mine, attr)
                assert np.array_equal(np.asarray(want, dtype=np.float64), np.asarray(got, dtype=np.float64)), (n, k, name, attr, want, got)
            if name == "genset":
                for goal in (0, 1):
                    assert lst[0].next_status(goal) == mine.next_status(goal)
                    assert lst[0].next_max_production(goal) == mine.next_max_production(goal)
            assert lst[0].module_type == mine.module_type and tuple(lst[0].name) == tuple(mine.name)
            sd1, sd2 = lst[0].state_dict(), mine.state_dict()
            assert list(sd1) == list(sd2) and np.array_equal(np.array(list(sd1.values()), dtype=float), np.array(list(sd2.values()), dtype=float))
    act = {name: [rng.random(2) if name == "genset" else rng.random()] for name in ref.get_empty_action()}
    d1, d2 = ref.from_normalized(act, act=True), ours.from_normalized(act, act=True)
    same_nested(d1, d2, (n, "from_normalized"))
    same_nested(ref.to_normalized(d1, act=True), ours.to_normalized(d2, act=True), (n, "to_normalized"))
    obs = ref.reset()
    obs = {k: v for k, v in obs.items() if k not in ("balance", "other")}
    same_nested(ref.from_normalized(obs, obs=True), ours.from_normalized(obs, obs=True), (n, "obs denormalised"))
    assert ref.get_forecast_horizon() == ours.get_forecast_horizon()


@pytest.mark.parametrize("g", range(16))
def test_random_fused_microgrids_with_shaper_and_trajectory_against_the_live_reference(g):
    """the reference's built-in reward shapers (on the device here) and a deterministic trajectory window, on random grids"""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    from pymgrid.microgrid.reward_shaping import BatteryDischargeShaper, PVCurtailmentShaper
    from pymgrid.microgrid.trajectory import DeterministicTrajectory
    import pymgrid_b200
    from pymgrid_b200 import modules as M
    from pymgrid_b200.compose import in_fused_scope
    T = 40

    def named_pv(mods, ns=None):
        # the shapers address modules by NAME: 'pv' and 'unbalanced_energy' (reward_shaping/*.py)
        return [("pv", m[1]) if isinstance(m, tuple) else m for m in mods]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m1, kw = draw_fused(np.random.default_rng(900 + g), R, T)
        m2, _ = draw_fused(np.random.default_rng(900 + g), M, T)
        m1, m2 = named_pv(m1), named_pv(m2)
        if not in_fused_scope(m2):
            pytest.skip("this order runs on the composed path (Python shapers there; covered by compose.npz)")
        shaper = (PVCurtailmentShaper(), "pv_curtailment") if g % 2 else (BatteryDischargeShaper(), "battery_discharge")
        m1 = m1 + [("unbalanced_energy", R.UnbalancedEnergyModule(False, **kw))]
        m2 = m2 + [("unbalanced_energy", M.UnbalancedEnergyModule(False, **kw))]
        ref = pymgrid.Microgrid(m1, add_unbalanced_module=False, reward_shaping_func=shaper[0], trajectory_func=DeterministicTrajectory(5, 25))
        ours = pymgrid_b200.Microgrid(m2, add_unbalanced_module=False, reward_shaping_func=shaper[1], trajectory_func=lambda lo, hi: (5, 25))
    r1, r2 = ref.reset(), ours.reset()
    same_nested({k: v for k, v in r1.items() if k not in ("balance", "other")}, {k: v for k, v in r2.items() if k not in ("balance", "other")},
                (g, "reset"))
    assert ref.current_step == ours.current_step == 5
    rng = np.random.default_rng(g)
    for k in range(22):
        a = {name: [rng.random(2) if name == "genset" else rng.random()] for name in ref.get_empty_action()}
        outs = []
        for runner in (ref, ours):
            try:
                outs.append(runner.run(a))
            except Exception as exc:      # noqa: BLE001 -- the discharge shaper asserts its value lies in [-1, 1]
                outs.append(type(exc).__name__)
        if isinstance(outs[0], str) or isinstance(outs[1], str):
            assert outs[0] == outs[1], (g, k, outs)
            return
        (o1, s1, d1, i1), (o2, s2, d2, i2) = outs
        assert (s1 == s2 or (np.isnan(s1) and np.isnan(s2))) and d1 == d2, (g, k, s1, s2)
        assert d1 == (5 + k >= 24)
        same_nested(o1, o2, (g, k, "obs"))


def test_discrete_env_keeps_redundant_genset_lists_when_asked():
    """DiscreteMicrogridEnv(remove_redundant_gensets=False) (envs/discrete/discrete.py:60-80; priority_list.py:15-67): a genset
    whose running_min_production is 0 makes the genset-off lists redundant -- with the flag off they stay, so the action
    space, and what an action index means, match the reference's."""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    from pymgrid import envs as ref_envs
    import pymgrid_b200.modules as M
    from pymgrid_b200.envs import DiscreteMicrogridEnv

    def modules(lib):
        rng = np.random.default_rng(4)
        load, pv = 40 + 20 * rng.random(60), 30 * rng.random(60)
        return [lib.LoadModule(time_series=load), lib.RenewableModule(time_series=pv),
                lib.GensetModule(running_min_production=0, running_max_production=60, genset_cost=0.4),
                lib.BatteryModule(min_capacity=10, max_capacity=100, max_charge=30, max_discharge=30, efficiency=0.9, init_soc=0.5)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for flag in (True, False):
            ref = ref_envs.DiscreteMicrogridEnv(modules(pymgrid.modules), remove_redundant_gensets=flag)
            ours = DiscreteMicrogridEnv(modules(M), remove_redundant_gensets=flag)
            assert ref.action_space.n == ours.action_space.n, flag
            assert [[(e.module, e.module_actions, e.action) for e in a] for a in ref.actions_list] == \
                   [[(e.module, e.module_actions, e.action) for e in a] for a in ours.actions_list], flag
            ref.reset(), ours.reset()
            for a in range(ref.action_space.n):
                o1, r1, d1, _ = ref.step(a)
                o2, r2, d2, _ = ours.step(a)
                assert r1 == r2 and d1 == d2, (flag, a)
                assert np.array_equal(np.asarray(o1, dtype=np.float64), np.asarray(o2, dtype=np.float64)), (flag, a)


def test_single_env_raises_where_the_reference_raises():
    """One microgrid behind an env: stepping past the end of the series raises IndexError like the reference's env (a batch
    reports the event per env through `flags` and a NaN reward instead); observation_keys=None is the reference's default."""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    from pymgrid import envs as ref_envs
    import pymgrid_b200.modules as M
    from pymgrid_b200.envs import DiscreteMicrogridEnv

    def modules(lib):
        rng = np.random.default_rng(6)
        load, pv = 40 + 20 * rng.random(30), 30 * rng.random(30)
        return [lib.LoadModule(time_series=load), lib.RenewableModule(time_series=pv),
                lib.GensetModule(running_min_production=5, running_max_production=60, genset_cost=0.4),
                lib.BatteryModule(min_capacity=10, max_capacity=100, max_charge=30, max_discharge=30, efficiency=0.9, init_soc=0.5)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_envs.DiscreteMicrogridEnv(modules(pymgrid.modules))
        ours = DiscreteMicrogridEnv(modules(M), observation_keys=None)
    ref.reset(), ours.reset()
    for k in range(30):
        _, r1, d1, _ = ref.step(k % ref.action_space.n)
        _, r2, d2, _ = ours.step(k % ours.action_space.n)
        assert r1 == r2 and d1 == d2, k
    assert d1 and d2
    with pytest.raises(IndexError):
        ref.step(0)
    with pytest.raises(IndexError):
        ours.step(0)
