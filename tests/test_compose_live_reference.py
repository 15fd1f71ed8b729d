"""Build container only (`reference` marker: needs /root/reference): randomised COMPOSITIONS -- random module counts,
names, parameters (incl. the corner values: min_capacity 0, running_min_production 0 or = max, max_export 0, three-column
grids, slow gensets with and without abortion, horizons longer than the rest of the series, a window inside the series) --
stepped side by side through the live, unmodified reference, the Python oracle and the host build of the CUDA path's C
source.  Rewards, dones, observations, infos and battery / genset state must agree bit for bit, and so must the step at
which the reference raises."""
import ctypes
import warnings

import numpy as np
import pymgrid_b200  # noqa: F401
import pytest

from tests import hostsim

pytestmark = pytest.mark.reference
N_GRIDS = 60


def draw(rng, ns, T):
    """a random module list built with the classes of namespace `ns` (the reference's or ours: same constructor calls)"""
    calls = []
    H = lambda: dict(forecaster=None) if rng.random() < 0.3 else dict(forecaster="oracle", forecast_horizon=int(rng.choice([0, 1, 3, 7, 23, T + 5])))  # noqa: E731
    window = dict(initial_step=int(rng.choice([0, 0, 2])))
    final = int(rng.choice([-1, -1, T - 6]))
    for _ in range(int(rng.integers(0, 4))):
        ts = rng.uniform(5, 200) * rng.random(T)
        if rng.random() < 0.3:
            ts[int(rng.integers(0, T - 4)):][:3] = 0.0
        calls.append(("LoadModule", dict(time_series=ts, final_step=final, **window, **H())))
    for _ in range(int(rng.integers(0, 4))):
        ts = rng.uniform(5, 200) * np.clip(rng.random(T) - 0.3, 0, None)
        calls.append(("RenewableModule", dict(time_series=ts, final_step=final, **window, **H())))
    for _ in range(int(rng.integers(0, 3))):
        mx = float(rng.uniform(20, 400))
        mn = float(rng.choice([0.0, rng.uniform(0.05, 0.5) * mx]))
        calls.append(("BatteryModule", dict(min_capacity=mn, max_capacity=mx, max_charge=float(rng.uniform(0.05, 1.2) * mx),
                                            max_discharge=float(rng.uniform(0.05, 1.2) * mx),
                                            efficiency=float(rng.choice([1.0, rng.uniform(0.5, 0.999)])),
                                            battery_cost_cycle=float(rng.choice([0.0, rng.uniform(0.001, 0.8)])),
                                            init_soc=float(rng.uniform(mn / mx, 1.0)), **window)))
    for _ in range(int(rng.integers(0, 3))):
        gmax = float(rng.uniform(20, 200))
        gmin = float(rng.choice([0.0, gmax, rng.uniform(0.05, 0.9) * gmax], p=[0.25, 0.1, 0.65]))
        calls.append(("GensetModule", dict(running_min_production=gmin, running_max_production=gmax, genset_cost=float(rng.uniform(0, 1)),
                                           co2_per_unit=float(rng.choice([0.0, rng.uniform(0.1, 3)])),
                                           cost_per_unit_co2=float(rng.choice([0.0, rng.uniform(0.01, 0.5)])),
                                           start_up_time=int(rng.integers(0, 4)), wind_down_time=int(rng.integers(0, 4)),
                                           allow_abortion=bool(rng.integers(0, 2)), init_start_up=bool(rng.integers(0, 2)), **window)))
    for _ in range(int(rng.integers(0, 3))):
        cols = int(rng.choice([3, 4]))
        ts = np.stack([rng.uniform(0.05, 0.9, T), rng.choice([0.0, 1.0]) * rng.uniform(0.0, 0.4, T), rng.uniform(0.0, 0.6, T),
                       (rng.random(T) > rng.choice([0.0, 0.3])).astype(np.float64)], axis=1)[:, :cols]
        calls.append(("GridModule", dict(max_import=float(rng.uniform(10, 300)), max_export=float(rng.choice([0.0, rng.uniform(10, 300)])),
                                         time_series=ts, cost_per_unit_co2=float(rng.choice([0.0, rng.uniform(0.01, 0.5)])),
                                         final_step=final, **window, **H())))
    slack = dict(raise_errors=False, loss_load_cost=float(rng.uniform(0.5, 20)), overgeneration_cost=float(rng.uniform(0, 5)), **window)
    calls.append(("UnbalancedEnergyModule", slack))
    order = rng.permutation(len(calls))
    names = {"RenewableModule": [None, "pv", "PV", "wind"], "LoadModule": [None, "zload"], "BatteryModule": [None, "storage"]}
    picked = {cls: (opts[int(rng.integers(0, len(opts)))] if rng.random() < 0.4 else None) for cls, opts in names.items()}
    out = []
    for i in order:
        cls, kw = calls[i]
        m = getattr(ns, cls)(**kw)
        name = picked.get(cls)
        out.append((name, m) if name is not None else m)
    return out


def test_random_compositions_against_the_live_reference():
    from oracle.compose import ComposedOracle, Raised
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    from pymgrid_b200 import modules as M
    from pymgrid_b200.compose import ComposedMicrogrid
    lib = ctypes.CDLL(hostsim.build())
    hostsim.select(lib)
    checked_steps = raised = built = logs = 0
    kinds = set()
    for g in range(N_GRIDS):
        T = 40
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                ref = pymgrid.Microgrid(draw(np.random.default_rng(9000 + g), R, T), add_unbalanced_module=False)
            except Exception:      # noqa: BLE001 -- e.g. a final_step the reference's own modules disagree on: nothing to compare
                continue
            mods = draw(np.random.default_rng(9000 + g), M, T)
            orc = ComposedOracle(mods, add_unbalanced_module=False)
            ours = ComposedMicrogrid(mods, add_unbalanced_module=False, obs_order="container")
        built += 1
        order = [(name, j) for name, lst in ref.modules.iterdict() for j in range(len(lst))]
        assert order == [(m.name, m.index) for m in orc.listing] == [(s.name, s.index) for s in ours.composition.slots], g
        flat = lambda obs: np.concatenate([np.asarray(obs[n][j], dtype=np.float64).ravel() for n, j in order] + [np.zeros(0)])  # noqa: E731
        assert np.array_equal(flat(ref.reset()), flat(orc.reset())) and np.array_equal(flat(ref.reset()), flat(ours.reset())), g
        rng = np.random.default_rng(100 + g)
        for k in range(T + 2):
            if ours.current_step == T:          # the next step runs off the series: compare the whole log first
                l0, l2 = ref.get_log(), ours.get_log()
                assert [tuple(c) for c in l0.columns] == [tuple(c) for c in l2.columns], g
                assert np.array_equal(l0.to_numpy(dtype=np.float64), l2.to_numpy(dtype=np.float64), equal_nan=True), g
                assert np.array_equal(np.array(l0.index), np.array(l2.index)), g
                logs += 1
            normalized = k % 3 != 2
            control = {}
            for name, lst in ref.controllable.iterdict():
                vals = []
                for mod in lst:
                    n = mod.action_space.shape[0]
                    if normalized:
                        a = rng.random(n)
                    else:
                        lo, hi = np.atleast_1d(mod.min_act).astype(float), np.atleast_1d(mod.max_act).astype(float)
                        a = lo - 0.3 * (hi - lo) + 1.6 * (hi - lo) * rng.random(n)
                        if n == 2:
                            a[0], a[1] = rng.random(), max(a[1], 0.0)
                    vals.append(a if n > 1 else float(a[0]))
                control[name] = vals
            outs = []
            for runner in (ref, orc, ours):
                try:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        outs.append(runner.run({k_: list(v) for k_, v in control.items()}, normalized=normalized))
                except Raised as exc:
                    outs.append(exc.kind)
                except Exception as exc:      # noqa: BLE001
                    outs.append(type(exc).__name__)
            if isinstance(outs[0], str):
                assert outs[1] == outs[0] and outs[2] == outs[0], (g, k, outs)
                raised += 1
                kinds.add(outs[0])
                break
            (o0, r0, d0, i0), (o1, r1, d1, i1), (o2, r2, d2, i2) = outs
            assert r0 == r1 == r2 and d0 == d1 == d2, (g, k, r0, r1, r2)
            assert np.array_equal(flat(o0), flat(o1)) and np.array_equal(flat(o0), flat(o2)), (g, k)
            for n_, j in order:
                assert dict(i0[n_][j]) == dict(i1[n_][j]) == dict(i2[n_][j]), (g, k, n_, j)
            checked_steps += 1
    print(f"{built} of {N_GRIDS} compositions built, {checked_steps} steps compared, {logs} full logs compared, {raised} runs ended where the reference raised: {sorted(kinds)}")
    assert built > N_GRIDS // 2 and checked_steps > 15 * built and raised > 0 and logs > built // 2


def test_random_compositions_priority_lists_against_the_live_reference():
    """DiscreteMicrogridEnv action tables (with / without the redundant genset lists), the controls random actions expand
    to, rewards and observations, and RuleBasedControl's sorted list and run, on random compositions"""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    from pymgrid.algos import RuleBasedControl
    from pymgrid.envs import DiscreteMicrogridEnv
    import pymgrid_b200
    from pymgrid_b200 import modules as M
    from pymgrid_b200.compose import MAX_PRIORITY_ELEMENTS, ComposedDiscreteEnv, ComposedMicrogrid
    lib = ctypes.CDLL(hostsim.build())
    hostsim.select(lib)
    rows = lambda pls: [[(el.module[0], el.module[1], el.module_actions, el.action) for el in pl] for pl in pls]      # noqa: E731
    compared = steps = 0
    for g in range(40):
        T = 30
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_mods = draw(np.random.default_rng(7000 + g), R, T)
            n_el = sum(2 if type(m[1] if isinstance(m, tuple) else m).__name__ == "GensetModule" else 1
                       for m in ref_mods if type(m[1] if isinstance(m, tuple) else m).__name__ in ("GensetModule", "BatteryModule", "GridModule"))
            if not 1 <= n_el <= 6 or not any(type(m[1] if isinstance(m, tuple) else m).__name__ == "LoadModule" for m in ref_mods):
                continue
            try:
                ref_env = {f: DiscreteMicrogridEnv(draw(np.random.default_rng(7000 + g), R, T), add_unbalanced_module=False,
                                                   remove_redundant_gensets=f) for f in (False, True)}
            except Exception:      # noqa: BLE001
                continue
            ours_env = {f: ComposedDiscreteEnv(draw(np.random.default_rng(7000 + g), M, T), add_unbalanced_module=False,
                                               remove_redundant_gensets=f, obs_order="container") for f in (False, True)}
        assert n_el <= MAX_PRIORITY_ELEMENTS
        for f in (False, True):
            assert rows(ours_env[f].actions_list) == rows(ref_env[f].actions_list), (g, f)
        ref, ours = ref_env[False], ours_env[False]
        order = [(name, j) for name, lst in ref.modules.iterdict() for j in range(len(lst))]
        ref.reset(), ours.reset()
        rng = np.random.default_rng(300 + g)
        for k in range(20):
            a = int(rng.integers(0, ref.action_space.n))
            try:
                control = ref._get_action(a)
                o0, r0, d0, _ = pymgrid.Microgrid.run(ref, control, normalized=False)
            except Exception as exc:      # noqa: BLE001
                with pytest.raises(type(exc)):
                    ours.step(a)
                break
            o2, r2, d2, _ = ours.step(a)
            assert r0 == r2 and d0 == d2, (g, k, a, r0, r2)
            flat = np.concatenate([np.asarray(o0[n][j], dtype=np.float64).ravel() for n, j in order] + [np.zeros(0)])
            assert np.array_equal(flat, o2), (g, k)
            steps += 1
        # rule-based control on fresh copies
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_rbc = RuleBasedControl(pymgrid.Microgrid(draw(np.random.default_rng(7000 + g), R, T), add_unbalanced_module=False))
            ours_rbc = pymgrid_b200.algos.RuleBasedControl(ComposedMicrogrid(draw(np.random.default_rng(7000 + g), M, T),
                                                                            add_unbalanced_module=False))
        assert rows([ours_rbc.priority_list]) == rows([ref_rbc.priority_list]), g
        try:
            want = ref_rbc.run(max_steps=12)
        except Exception as exc:      # noqa: BLE001
            with pytest.raises(type(exc)):
                ours_rbc.run(max_steps=12)
        else:
            got = ours_rbc.run(max_steps=12)
            assert [tuple(c) for c in got.columns] == [tuple(c) for c in want.columns], g
            assert np.array_equal(got.to_numpy(dtype=np.float64), want.to_numpy(dtype=np.float64), equal_nan=True), g
        compared += 1
    print(f"{compared} compositions, {steps} discrete steps compared")
    assert compared >= 10 and steps > 100


def test_random_modules_on_their_own_against_the_live_reference():
    """BaseMicrogridModule.step on randomly parameterised modules of every kind, used without a Microgrid: observation,
    reward, done, info and the exception types, side by side with the live reference"""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid.modules as R
    import pymgrid_b200.compose as cp
    from pymgrid_b200 import modules as M
    hostsim.select(ctypes.CDLL(hostsim.build()))
    compared = raised = 0
    try:
        for g in range(25):
            T = 30
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref_mods, our_mods = draw(np.random.default_rng(5000 + g), R, T), draw(np.random.default_rng(5000 + g), M, T)
            rng = np.random.default_rng(g)
            for a, b in zip(ref_mods, our_mods):
                a, b = (a[1] if isinstance(a, tuple) else a), (b[1] if isinstance(b, tuple) else b)
                kind = type(a).__name__
                for k in range(T + 1):
                    normalized = bool(rng.integers(0, 2)) and kind != "UnbalancedEnergyModule"
                    if kind == "LoadModule":
                        act = np.array([])
                    elif kind == "GensetModule":
                        act = np.array([rng.random(), rng.random() if normalized else rng.uniform(0, 1.5 * a.running_max_production)])
                    else:
                        lo, hi = (0.0, 1.0) if normalized else ((-60.0, 60.0) if kind != "RenewableModule" else (0.0, 250.0))
                        act = float(rng.uniform(lo - 0.1 * (hi - lo) * (not normalized), hi))
                    outs = []
                    for mod in (a, b):
                        try:
                            with warnings.catch_warnings():
                                warnings.simplefilter("ignore")
                                outs.append(mod.step(act, normalized=normalized))
                        except Exception as exc:      # noqa: BLE001
                            outs.append(type(exc).__name__)
                    if isinstance(outs[0], str) or isinstance(outs[1], str):
                        assert outs[0] == outs[1], (g, kind, k, outs)
                        raised += 1
                        break
                    (o0, r0, d0, i0), (o1, r1, d1, i1) = outs
                    assert r0 == r1 and d0 == d1 and dict(i0) == dict(i1), (g, kind, k, r0, r1, i0, i1)
                    assert np.array_equal(np.asarray(o0, dtype=np.float64).ravel(), o1), (g, kind, k)
                    compared += 1
    finally:
        hostsim.select(None)
    print(f"{compared} module steps compared, {raised} runs ended where the reference raised")
    assert compared > 2000 and raised > 20


def test_composed_module_views_against_the_live_reference():
    """every attribute the reference's in-repo callers read from a module on random compositions, after a few
    steps; state dicts, cost info, normalisation helpers"""
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    from pymgrid_b200 import modules as M
    from pymgrid_b200.compose import ComposedMicrogrid
    attrs = ("max_production", "max_consumption", "min_production", "production_marginal_cost", "absorption_marginal_cost",
             "marginal_cost", "state", "min_obs", "max_obs", "min_act", "max_act", "is_source", "is_sink", "current_step",
             "initial_step", "final_step", "soc", "current_charge", "current_status", "goal_status", "import_price", "export_price",
             "co2_per_kwh", "grid_status", "forecast_horizon", "max_capacity", "efficiency", "running_max_production", "max_import",
             "loss_load_cost", "current_load", "current_renewable")
    lib = ctypes.CDLL(hostsim.build())
    hostsim.select(lib)
    compared = 0
    for g in range(30):
        T = 30
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                ref = pymgrid.Microgrid(draw(np.random.default_rng(3000 + g), R, T), add_unbalanced_module=False)
            except Exception:      # noqa: BLE001
                continue
            ours = ComposedMicrogrid(draw(np.random.default_rng(3000 + g), M, T), add_unbalanced_module=False)
        rng = np.random.default_rng(g)
        for k in range(3):
            a = {name: [rng.random(m.action_space.shape[0]) if m.action_space.shape[0] > 1 else rng.random() for m in lst]
                 for name, lst in ref.controllable.iterdict()}
            try:
                ref.run(a)
            except Exception:      # noqa: BLE001
                break
            ours.run(a)
            for name, lst in ref.modules.iterdict():
                for j, theirs in enumerate(lst):
                    mine = ours.modules[name][j]
                    for attr in attrs:
                        try:
                            want = getattr(theirs, attr)
                        except Exception:      # noqa: BLE001
                            continue
                        if want is NotImplemented:
                            continue
                        got = getattr(mine, attr)
                        same = np.array_equal(np.asarray(want, dtype=np.float64), np.asarray(got, dtype=np.float64), equal_nan=True)
                        assert same, (g, k, name, j, attr, want, got)
                        compared += 1
                    sd1, sd2 = theirs.state_dict(), mine.state_dict()
                    assert list(sd1) == list(sd2), (g, name, j)
                    assert np.array_equal(np.array(list(sd1.values()), dtype=float), np.array(list(sd2.values()), dtype=float))
            c1, c2 = ref.get_cost_info(), ours.get_cost_info()
            assert {k_: [dict(d) for d in v] for k_, v in c1.items()} == {k_: [dict(d) for d in v] for k_, v in c2.items()}
    assert compared > 3000
