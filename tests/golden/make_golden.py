"""Generate the golden fixtures under tests/golden/ from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference is imported through oracle/ref_loader.py (gym / plotting stand-ins; reference source untouched).
The fixtures pin the C oracle (tests/test_oracle_vs_golden.py) and the CUDA path (tests/test_gpu_parity.py)
to the reference's actual outputs:

  pymgrid25_steps.npz   all 25 scenarios: normalised random steps at the start of the year and across the end of
                        the series (forecast padding, done flag, last valid step), plus unnormalised steps
  pymgrid25_year.npz    scenarios 0, 1, 2 (the three architectures): full 8760-step year, reward checkpoints
  discrete.npz          DiscreteMicrogridEnv on all 25 scenarios (H=23) and on the 15 grid scenarios with H=24
                        (BASELINE config 4): action tables, expanded controls, rewards, flat observations
  genset_machine.npz    exhaustive genset status transitions, U, D in 0..4, both abortion settings
  custom.npz            small hand-built microgrids: slow gensets, weak grid, short series, no forecaster
  shaped.npz            the two built-in reward shapers (microgrid/reward_shaping/) on the three architectures
  noisy_forecast.npz    GaussianNoiseForecaster (forecast/forecaster.py:220-275) under the legacy numpy seed: flat
                        observations after reset and after each step, at the start and across the end of the series
"""
import itertools
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.simplefilter("ignore")

from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from pymgrid import Microgrid  # noqa: E402
from pymgrid.envs import DiscreteMicrogridEnv  # noqa: E402
from pymgrid.modules import BatteryModule, GensetModule, GridModule, LoadModule, RenewableModule  # noqa: E402

SORTED_KEYS = ("battery", "genset", "grid", "load", "pv", "renewable")   # alphabetical, as gym.spaces.Dict sorts
CONTROL_KEYS = ("genset", "battery", "grid")                            # container order of controllables
INFO_COLS = 16


def flat_obs(obs):
    return np.concatenate([np.asarray(x, dtype=np.float64).ravel() for k in SORTED_KEYS if k in obs for x in obs[k]])


def module_reward(m, name):
    """reward the module logged for the step just taken (ModularLogger, utils/logger.py:18-28)"""
    if not hasattr(m.modules, name):
        return 0.0
    return float(getattr(m.modules, name)[0].log_dict()["reward"][-1])


def info_vec(info, m=None):
    v = np.zeros(INFO_COLS)
    g = lambda name, key: float(info[name][0].get(key, 0.0)) if name in info else 0.0  # noqa: E731
    pv = "pv" if "pv" in info else "renewable"
    ub = "unbalanced_energy" if "unbalanced_energy" in info else "balancing"
    v[0] = g("load", "absorbed_energy")
    v[1], v[2] = g(pv, "provided_energy"), g(pv, "curtailment")
    v[3], v[4] = g(ub, "provided_energy"), g(ub, "absorbed_energy")
    v[5], v[6] = g("genset", "provided_energy"), g("genset", "co2_production")
    v[7], v[8] = g("battery", "provided_energy"), g("battery", "absorbed_energy")
    v[9], v[10], v[11] = g("grid", "provided_energy"), g("grid", "absorbed_energy"), g("grid", "co2_production")
    if m is not None:
        ub_name = "unbalanced_energy" if hasattr(m.modules, "unbalanced_energy") else "balancing"
        v[12], v[13], v[14], v[15] = (module_reward(m, "genset"), module_reward(m, "battery"), module_reward(m, "grid"),
                                      module_reward(m, ub_name))
    return v


def control_from_flat(m, flat):
    ctrl, i = {}, 0
    for k in CONTROL_KEYS:
        if hasattr(m.modules, k):
            if k == "genset":
                ctrl[k] = [np.array(flat[i:i + 2])]
                i += 2
            else:
                ctrl[k] = [float(flat[i])]
                i += 1
    return ctrl


def n_act(m):
    return sum((2 if k == "genset" else 1) for k in CONTROL_KEYS if hasattr(m.modules, k))


def state_vec(m):
    b = m.modules.battery[0]
    out = [m.current_step, b.current_charge]
    if hasattr(m.modules, "genset"):
        g = m.modules.genset[0]
        out += [g._current_status, g._goal_status, g._steps_until_up, g._steps_until_down]
    else:
        out += [0, 0, 0, 0]
    return np.array(out, dtype=np.float64)


def run_segment(m, actions, normalized=True):
    rewards, dones, obs, infos, states = [], [], [], [], []
    for a in actions:
        o, r, d, info = m.run(control_from_flat(m, a), normalized=normalized)
        rewards.append(r); dones.append(d); obs.append(flat_obs(o)); infos.append(info_vec(info, m)); states.append(state_vec(m))
    return (np.array(rewards), np.array(dones, dtype=np.uint8), np.stack(obs), np.stack(infos), np.stack(states))


def unnormalised_actions(m, rng, n):
    """Unnormalised controls spanning beyond the module limits (so every clip branch is hit)."""
    out = []
    for _ in range(n):
        row = []
        if hasattr(m.modules, "genset"):
            g = m.modules.genset[0]
            row += [float(rng.integers(0, 2)) if rng.random() < 0.7 else rng.random(),
                    rng.uniform(0.0, 1.3) * g.running_max_production]
        b = m.modules.battery[0]
        row += [rng.uniform(-1.5, 1.5) * b.max_charge]
        if hasattr(m.modules, "grid"):
            gr = m.modules.grid[0]
            row += [rng.uniform(-1.2, 1.2) * gr.max_import]
        out.append(row)
    return np.array(out)


def make_pymgrid25_steps():
    out = {}
    n_start, n_end, end_from = 48, 44, 8720
    for n in range(25):
        rng = np.random.default_rng(1000 + n)
        m = Microgrid.from_scenario(n)
        na = n_act(m)
        a0 = rng.random((n_start, na))
        r0, d0, o0, i0, s0 = run_segment(m, a0)
        # unnormalised segment continues from there
        au = unnormalised_actions(m, rng, 24)
        ru, du, ou, iu, su = run_segment(m, au, normalized=False)
        # across the end of the series: jump to end_from, run until the last valid step (t = 8759)
        m2 = Microgrid.from_scenario(n)
        m2.initial_step = end_from
        reset_obs = flat_obs(m2.reset())
        a1 = rng.random((8760 - end_from, na))[:n_end] if n_end < 8760 - end_from else rng.random((8760 - end_from, na))
        r1, d1, o1, i1, s1 = run_segment(m2, a1)
        for k, v in dict(a0=a0, r0=r0, d0=d0, o0=o0, i0=i0, s0=s0, au=au, ru=ru, du=du, ou=ou, iu=iu, su=su,
                         reset_obs=reset_obs, a1=a1, r1=r1, d1=d1, o1=o1, i1=i1, s1=s1).items():
            out[f"s{n}_{k}"] = v
        print("steps", n, r0[:2], d1[-3:], s1[-1][:2])
    out["end_from"] = np.array(end_from)
    np.savez_compressed(os.path.join(HERE, "pymgrid25_steps.npz"), **out)


def make_year():
    out = {}
    for n in (0, 1, 2):
        m = Microgrid.from_scenario(n)
        na = n_act(m)
        actions = np.random.default_rng(0).random((8760, na))     # regenerated identically in the tests
        rewards = np.empty(8760)
        dones = np.empty(8760, dtype=np.uint8)
        for t in range(8760):
            _, rewards[t], dones[t], _ = m.run(control_from_flat(m, actions[t]))
        out[f"s{n}_rewards"] = rewards
        out[f"s{n}_first_done"] = np.array(int(np.argmax(dones)))
        out[f"s{n}_final_state"] = state_vec(m)
        print("year", n, rewards.sum(), out[f"s{n}_first_done"], state_vec(m))
    # the legacy-seed known-answer of SURVEY.md 8c: np.random.seed(0), sample_action(strict_bound=True), scenario 0
    m = Microgrid.from_scenario(0)
    np.random.seed(0)
    a = m.sample_action(strict_bound=True)
    _, r, _, info = m.run(a)
    out["legacy_s0_action"] = np.array([a["battery"][0], a["grid"][0]])
    out["legacy_s0_reward"] = np.array(r)
    out["legacy_s0_info"] = info_vec(info, m)
    np.savez_compressed(os.path.join(HERE, "pymgrid25_year.npz"), **out)


MOD_ID = {"genset": 0, "battery": 1, "grid": 2}


def action_table(env):
    n, width = len(env.actions_list), max(len(pl) for pl in env.actions_list)
    mod = np.full((n, width), -1, dtype=np.int8)
    act = np.zeros((n, width), dtype=np.int8)
    for i, pl in enumerate(env.actions_list):
        for j, el in enumerate(pl):
            mod[i, j] = MOD_ID[el.module[0]]
            act[i, j] = el.action
    return mod, act


def flat_control(env, ctrl):
    row = []
    for k in CONTROL_KEYS:
        if k in ctrl:
            row += list(np.asarray(ctrl[k][0], dtype=np.float64).ravel())
    return np.array(row)


def make_discrete():
    out = {}
    n_steps = 40
    for horizon, scenarios in ((23, range(25)), (24, [n for n in range(25) if n in (0, 4, 6, 11, 12, 14, 16, 1, 8, 9, 10, 13, 18, 22, 24)])):
        for n in scenarios:
            m = Microgrid.from_scenario(n)
            if horizon != 23:
                m.set_forecaster("oracle", forecast_horizon=horizon)
            env = DiscreteMicrogridEnv.from_microgrid(m)
            rng = np.random.default_rng(3000 + n + horizon)
            mod, act = action_table(env)
            acts = rng.integers(0, env.action_space.n, n_steps)
            reset_obs = env.reset()
            ctrls, rewards, dones, obs = [], [], [], []
            for a in acts:
                ctrls.append(flat_control(env, env._get_action(int(a))))
                o, r, d, _ = env.step(int(a))
                rewards.append(r); dones.append(d); obs.append(np.asarray(o, dtype=np.float64))
            tag = f"h{horizon}_s{n}"
            out[f"{tag}_table_mod"], out[f"{tag}_table_act"] = mod, act
            out[f"{tag}_actions"] = acts.astype(np.int32)
            out[f"{tag}_controls"] = np.stack(ctrls)
            out[f"{tag}_rewards"] = np.array(rewards)
            out[f"{tag}_dones"] = np.array(dones, dtype=np.uint8)
            out[f"{tag}_obs"] = np.stack(obs)
            out[f"{tag}_reset_obs"] = np.asarray(reset_obs, dtype=np.float64)
            out[f"{tag}_obs_dim"] = np.array(env.observation_space.shape[0])
            print("discrete", tag, env.action_space.n, env.observation_space.shape, rewards[:2])
    np.savez_compressed(os.path.join(HERE, "discrete.npz"), **out)


def make_genset_machine():
    rows = []
    for U, D, abort, init in itertools.product(range(5), range(5), (True, False), (True, False)):
        for goals in itertools.product((0, 1), repeat=6):
            g = GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5,
                             start_up_time=U, wind_down_time=D, allow_abortion=abort, init_start_up=init)
            seq = []
            for goal in goals:
                pred = g.next_status(goal)
                g.update_status(goal)
                seq.append((g._current_status, g._goal_status, g._steps_until_up, g._steps_until_down, pred))
            rows.append((U, D, int(abort), int(init), goals, seq))
    params = np.array([r[:4] for r in rows], dtype=np.int8)
    goals = np.array([r[4] for r in rows], dtype=np.int8)
    states = np.array([r[5] for r in rows], dtype=np.int8)
    # fractional goals exercise Python's round-half-even
    frac = np.array([0.0, 0.25, 0.5, 0.5000001, 0.75, 1.0, 0.4999999])
    g = GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5)
    frac_cs = []
    for f in frac:
        g.update_status(f)
        frac_cs.append(g._current_status)
    np.savez_compressed(os.path.join(HERE, "genset_machine.npz"), params=params, goals=goals, states=states,
                        frac_goals=frac, frac_status=np.array(frac_cs, dtype=np.int8))
    print("genset machine", states.shape)


def custom_grids():
    """Hand-built microgrids (reference constructors) covering what pymgrid25 does not."""
    rng = np.random.default_rng(77)
    T = 60
    t = np.arange(T)
    load = 60 + 30 * np.sin(t / 5.0) + rng.random(T) * 5
    pv = np.clip(50 * np.sin(t / 7.0), 0, None)
    price = np.where((t // 6) % 2 == 0, 0.3, 0.7)
    status = (rng.random(T) > 0.25).astype(float)
    co2 = 0.2 + 0.1 * rng.random(T)
    grid_ts = np.stack([price, 0.1 * np.ones(T), co2, status], axis=1)
    specs = []
    for (U, D, abort, init, H, final_step, eff, weak, has_gen, has_grid) in [
        (2, 3, True, True, 4, -1, 0.9, True, True, True),
        (1, 1, False, False, 4, 50, 0.8, True, True, True),
        (3, 0, True, False, 0, -1, 1.0, False, True, False),
        (0, 2, True, True, 7, 40, 0.95, True, True, True),
        (0, 0, True, True, 23, -1, 0.9, False, False, True),
        (4, 4, False, True, 2, -1, 0.9, False, True, False),
    ]:
        specs.append(dict(U=U, D=D, abort=abort, init=init, H=H, final_step=final_step, eff=eff, weak=weak,
                          has_gen=has_gen, has_grid=has_grid))
    return load, pv, grid_ts, specs


def build_custom(load, pv, grid_ts, s):
    fc = "oracle" if s["H"] > 0 else None
    H = s["H"] if s["H"] > 0 else 23
    mods = [LoadModule(time_series=load, forecaster=fc, forecast_horizon=H, final_step=s["final_step"]),
            ("pv", RenewableModule(time_series=pv, forecaster=fc, forecast_horizon=H, final_step=s["final_step"]))]
    if s["has_gen"]:
        mods.append(GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5, co2_per_unit=1.5,
                                 cost_per_unit_co2=0.2, start_up_time=s["U"], wind_down_time=s["D"],
                                 allow_abortion=s["abort"], init_start_up=s["init"]))
    mods.append(BatteryModule(min_capacity=10, max_capacity=100, max_charge=40, max_discharge=45,
                              efficiency=s["eff"], battery_cost_cycle=0.05, init_soc=0.6))
    if s["has_grid"]:
        ts = grid_ts.copy()
        if not s["weak"]:
            ts[:, 3] = 1.0
        mods.append(GridModule(max_import=70, max_export=30, time_series=ts, forecaster=fc, forecast_horizon=H,
                               final_step=s["final_step"], cost_per_unit_co2=0.15))
    return Microgrid(mods, loss_load_cost=9.0, overgeneration_cost=1.5)


def make_custom():
    load, pv, grid_ts, specs = custom_grids()
    out = dict(load=load, pv=pv, grid_ts=grid_ts, n=np.array(len(specs)))
    for i, s in enumerate(specs):
        m = build_custom(load, pv, grid_ts, s)
        rng = np.random.default_rng(500 + i)
        T = len(load)
        a = rng.random((T, n_act(m)))
        if s["has_gen"]:
            a[:, 0] = np.where(rng.random(T) < 0.5, np.round(a[:, 0]), a[:, 0])   # mix of crisp and fractional goals
        reset_obs = flat_obs(m.reset())
        r, d, o, info, st = run_segment(m, a)
        out[f"c{i}_spec"] = np.array([s["U"], s["D"], int(s["abort"]), int(s["init"]), s["H"], s["final_step"],
                                      s["eff"], int(s["weak"]), int(s["has_gen"]), int(s["has_grid"])], dtype=np.float64)
        for k, v in dict(a=a, r=r, d=d, o=o, i=info, s=st, reset_obs=reset_obs).items():
            out[f"c{i}_{k}"] = v
        # reset mid-way keeps battery / genset state
        out[f"c{i}_after_reset_obs"] = flat_obs(m.reset())
        print("custom", i, r[:2], d.sum(), st[-1])
    np.savez_compressed(os.path.join(HERE, "custom.npz"), **out)


def make_log():
    """Microgrid.get_log() after 30 random steps + a reset + 5 more steps, scenarios 0 / 1 / 2 (the three architectures)."""
    out = {}
    for n in (0, 1, 2):
        m = Microgrid.from_scenario(n)
        rng = np.random.default_rng(900 + n)
        a = rng.random((30, n_act(m)))
        run_segment(m, a)
        df = m.get_log()
        out[f"s{n}_actions"] = a
        out[f"s{n}_columns"] = np.array(["|".join(map(str, c)) for c in df.columns])
        out[f"s{n}_values"] = df.values.astype(np.float64)
        out[f"s{n}_index"] = df.index.values
        ser = m.state_series()
        out[f"s{n}_state_series_index"] = np.array(["|".join(map(str, c)) for c in ser.index])
        out[f"s{n}_state_series_values"] = ser.values.astype(np.float64)
        m.reset()
        assert len(m.get_log()) == 0
        print("log", n, df.shape)
    np.savez_compressed(os.path.join(HERE, "log.npz"), **out)


def make_rbc():
    """RuleBasedControl (algos/rbc/rbc.py): the automatically chosen priority list and the rewards it earns."""
    from pymgrid.algos import RuleBasedControl
    out = {}
    for n in (0, 1, 2, 5, 9, 13):
        m = Microgrid.from_scenario(n)
        rbc = RuleBasedControl(m)
        out[f"s{n}_list_mod"] = np.array([MOD_ID[el.module[0]] for el in rbc.priority_list], dtype=np.int8)
        out[f"s{n}_list_act"] = np.array([el.action for el in rbc.priority_list], dtype=np.int8)
        steps = None if n == 0 else 300
        df = rbc.run(max_steps=steps)
        out[f"s{n}_rewards"] = df[("balance", 0, "reward")].values.astype(np.float64)
        out[f"s{n}_final_state"] = state_vec(rbc.microgrid)
        print("rbc", n, [(el.module[0], el.action, el.marginal_cost) for el in rbc.priority_list], len(df), df[("balance", 0, "reward")].sum())
    np.savez_compressed(os.path.join(HERE, "rbc.npz"), **out)


def make_generator():
    """Real MicrogridGenerator grids (MicrogridGenerator.py, convert/): their module parameters, the (profile, scale)
    factorisation of their load / PV series -- verified bit-exact here -- and short rollouts."""
    import glob
    import pandas as pd
    from pymgrid.MicrogridGenerator import MicrogridGenerator
    D = os.path.join(os.path.dirname(sys.modules["pymgrid"].__file__), "data")
    loads = [pd.read_csv(f).values[:, 0].astype(float) for f in sorted(glob.glob(D + "/load/*.csv"))]
    pvs = [pd.read_csv(f).values[:, 0].astype(float) for f in sorted(glob.glob(D + "/pv/*.csv"))]
    co2s = [pd.read_csv(f).values[:, 0].astype(float) for f in sorted(glob.glob(D + "/co2/*.csv"))]
    gen = MicrogridGenerator(nb_microgrid=24, random_seed=7).generate_microgrid()
    out = {"n": np.array(len(gen.microgrids))}
    for i, m in enumerate(gen.microgrids):
        names = [n for n, _ in m.modules.iterdict()]
        assert names[1] == "PV"
        L, P = m.modules.load[0].time_series[:, 0], m.modules["PV"][0].time_series[:, 0]
        lp = [k for k, b in enumerate(loads) if np.allclose(-L / b, (-L / b)[0], rtol=1e-9)]
        assert len(lp) == 1
        size = round(float((-L).max()))
        load_scale = size / loads[lp[0]].max()
        assert (-(loads[lp[0]] * load_scale) == L).all()
        pp = [k for k, b in enumerate(pvs) if np.allclose(P[b > 0] / b[b > 0], (P[b > 0] / b[b > 0])[0], rtol=1e-9)]
        assert len(pp) == 1
        ks = [k for k in range(30, 151) if (pvs[pp[0]] * ((-L).max() * (k / 100) / pvs[pp[0]].max()) == P).all()]
        assert len(ks) >= 1
        pv_scale = (-L).max() * (ks[0] / 100) / pvs[pp[0]].max()
        b = m.modules.battery[0]
        rec = [lp[0], load_scale, pp[0], pv_scale, b.min_capacity, b.max_capacity, b.max_charge, b.max_discharge, b.efficiency,
               b.battery_cost_cycle, b.current_charge]
        has_gen, has_grid = hasattr(m.modules, "genset"), hasattr(m.modules, "grid")
        if has_gen:
            g = m.modules.genset[0]
            rec += [1, g.running_min_production, g.running_max_production, g.genset_cost, g.co2_per_unit, g.cost_per_unit_co2]
        else:
            rec += [0, 0, 0, 0, 0, 0]
        if has_grid:
            g = m.modules.grid[0]
            ts = np.asarray(g.time_series, dtype=float)
            tariff = 1 if ts[:, 0].max() > 0.5 else 2
            cid = [k for k, c in enumerate(co2s) if np.array_equal(c, ts[:, 2])]
            assert len(cid) == 1 and (ts[:, 1] == 0).all()
            rec += [1, g.max_import, g.max_export, g.cost_per_unit_co2, tariff, cid[0]]
            out[f"g{i}_status"] = np.packbits(ts[:, 3].astype(np.uint8))
            out[f"g{i}_import_price"] = ts[:, 0]
        else:
            rec += [0, 0, 0, 0, 0, 0]
        ub = m.modules.unbalanced_energy[0]
        rec += [ub.loss_load_cost, ub.overgeneration_cost, m.modules.load[0].forecast_horizon, m.final_step]
        out[f"g{i}_rec"] = np.array(rec, dtype=np.float64)
        rng = np.random.default_rng(4000 + i)
        a = rng.random((60, n_act(m)))
        rewards, dones, obs_rows, states = [], [], [], []
        for row in a:
            o, r, d, info = m.run(control_from_flat(m, row))
            rewards.append(r); dones.append(d); states.append(state_vec(m))
            obs_rows.append(np.concatenate([np.asarray(x, dtype=np.float64).ravel() for k in ("PV", "battery", "genset", "grid", "load")
                                            if k in o for x in o[k]]))
        out[f"g{i}_a"], out[f"g{i}_r"], out[f"g{i}_d"] = a, np.array(rewards), np.array(dones, dtype=np.uint8)
        out[f"g{i}_o"], out[f"g{i}_s"] = np.stack(obs_rows), np.stack(states)
        print("generator", i, names, lp, pp, ks[:1], rewards[0])
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)


def make_shaped():
    """Microgrid(reward_shaping_func=...): run returns the shaped reward, the balance log keeps both (microgrid.py:316-319).

    BatteryDischargeShaper asserts its value is in [-1, 1] (battery_discharge_shaper.py:33) and the reference evaluates
    the shaper twice per step: once in the mid-step balance() before the flex modules (microgrid.py:277, loss load not
    yet known) and once at the end.  A failed assert leaves the microgrid half stepped, so each segment stops at the
    first one; `raised` marks it."""
    from pymgrid.microgrid.reward_shaping import BatteryDischargeShaper, PVCurtailmentShaper
    out = {}
    n_seg, seg_len = 6, 40
    for n in (0, 1, 2, 13):
        for tag, shaper in (("pv", PVCurtailmentShaper), ("bat", BatteryDischargeShaper)):
            rng = np.random.default_rng(7000 + n)
            for seg in range(n_seg):
                m = Microgrid.from_scenario(n)
                m.reward_shaping_func = shaper()
                m.initial_step = int(rng.integers(0, 8000))
                m.reset()
                a = rng.random((seg_len, n_act(m)))
                if tag == "bat" and seg % 2 == 0:      # keep the battery near idle so that long stretches pass the assert
                    a[:, -2 if hasattr(m.modules, "grid") else -1] = rng.uniform(0.35, 0.55, seg_len)
                r, s, raised = [], [], -1
                for k, row in enumerate(a):
                    try:
                        r.append(m.run(control_from_flat(m, row))[1])
                    except AssertionError:
                        raised = k
                        break
                    s.append(state_vec(m))
                log = m._balance_logger.to_dict()      # get_log cannot assemble a half-stepped microgrid
                key = f"s{n}_{tag}_{seg}"
                out[key + "_t0"], out[key + "_a"], out[key + "_r"] = np.array(m.initial_step), a, np.array(r)
                out[key + "_s"], out[key + "_raised"] = np.array(s).reshape(len(s), 6), np.array(raised)
                out[key + "_log_reward"] = np.array(log.get("reward", []), dtype=np.float64)
                out[key + "_log_shaped"] = np.array(log.get("shaped_reward", []), dtype=np.float64)
                print("shaped", n, tag, seg, len(r), raised, r[:2])
    # priority-list dispatch keeps the battery within the load, so BatteryDischargeShaper runs for long stretches
    for n in (0, 1, 2, 13):
        rng = np.random.default_rng(7100 + n)
        env = DiscreteMicrogridEnv.from_scenario(n)
        env.reward_shaping_func = BatteryDischargeShaper()
        env.initial_step = int(rng.integers(0, 8000))
        env.reset()
        acts = rng.integers(0, env.action_space.n, 150)
        ctrls, r, s, raised = [], [], [], -1
        for k, a in enumerate(acts):
            ctrls.append(flat_control(env, env._get_action(int(a))))
            try:
                r.append(env.step(int(a))[1])
            except AssertionError:
                raised = k
                break
            s.append(state_vec(env))
        log = env._balance_logger.to_dict()
        key = f"s{n}_batd"
        out[key + "_t0"], out[key + "_actions"], out[key + "_controls"] = np.array(env.initial_step), acts.astype(np.int32), np.stack(ctrls)
        out[key + "_r"], out[key + "_s"], out[key + "_raised"] = np.array(r), np.array(s).reshape(len(s), 6), np.array(raised)
        out[key + "_log_reward"] = np.array(log.get("reward", []), dtype=np.float64)
        print("shaped discrete", n, len(r), raised, r[:4])
    out["n_seg"] = np.array(n_seg)
    np.savez_compressed(os.path.join(HERE, "shaped.npz"), **out)


VIEW_SCALARS = ("marginal_cost", "production_marginal_cost", "absorption_marginal_cost", "max_production",
                "max_consumption", "min_production", "is_source", "is_sink")
VIEW_VECTORS = ("state", "min_obs", "max_obs", "min_act", "max_act")


def make_views():
    """Module attributes the reference's own callers read (RBC / priority lists / MPC / notebooks), after a few
    random steps: costs, limits, state vectors, spaces, the genset look-ahead, grid price columns, container sorting."""
    out = {}
    for n in (0, 1, 2):
        m = Microgrid.from_scenario(n)
        rng = np.random.default_rng(9000 + n)
        for row in rng.random((7, n_act(m))):
            m.run(control_from_flat(m, row))
        out[f"s{n}_state"] = state_vec(m)
        out[f"s{n}_actions"] = np.random.default_rng(9000 + n).random((7, n_act(m)))
        for name, lst in m.modules.iterdict():
            mod = lst[0]
            for a in VIEW_SCALARS:
                try:
                    out[f"s{n}_{name}_{a}"] = np.array(float(getattr(mod, a)))
                except (AttributeError, NotImplementedError, TypeError):
                    pass
            for a in VIEW_VECTORS:
                out[f"s{n}_{name}_{a}"] = np.atleast_1d(np.asarray(getattr(mod, a), dtype=np.float64)).ravel()
            out[f"s{n}_{name}_type"] = np.array(",".join(mod.module_type))
            out[f"s{n}_{name}_n_act"] = np.array(mod.action_space.shape[0])
            st = mod.state
            out[f"s{n}_{name}_norm_state"] = np.atleast_1d(np.asarray(mod.to_normalized(st, obs=True), dtype=np.float64)).ravel()
        if hasattr(m.modules, "genset"):
            g = m.modules.genset[0]
            out[f"s{n}_genset_next"] = np.array([g.next_status(0), g.next_status(1), g.next_max_production(0), g.next_max_production(1),
                                                 g.next_min_production(0), g.next_min_production(1)], dtype=np.float64)
        if hasattr(m.modules, "grid"):
            g = m.modules.grid[0]
            out[f"s{n}_grid_columns"] = np.stack([g.import_price, g.export_price, g.co2_per_kwh])
        b = m.modules.battery[0]
        out[f"s{n}_battery_socs"] = np.array([b.soc, b.min_soc, b.max_soc])
        for kind in ("fixed", "flex", "controllable"):
            c = getattr(m, kind)
            for sub in ("sources", "sinks", "source_and_sinks"):
                names = list(getattr(c, sub).keys()) if hasattr(c, sub) else []
                out[f"s{n}_{kind}_{sub}"] = np.array(",".join(names))
        out[f"s{n}_horizon"] = np.array(m.get_forecast_horizon())
        ctrl = control_from_flat(m, rng.random(n_act(m)))
        dn = m.from_normalized(ctrl, act=True)
        out[f"s{n}_denorm_in"] = np.concatenate([np.ravel(v[0]) for v in ctrl.values()])
        out[f"s{n}_denorm_out"] = np.concatenate([np.ravel(v[0]) for v in dn.values()])
        print("views", n, out[f"s{n}_state"], out[f"s{n}_controllable_source_and_sinks"], out[f"s{n}_denorm_out"])
    np.savez_compressed(os.path.join(HERE, "views.npz"), **out)


NOISE_CASES = (   # scenario, {module: (std, increase_uncertainty, relative_noise)}
    (0, dict(load=(40.0, False, False), pv=(0.15, True, True), grid=(0.05, True, False))),
    (2, dict(load=(0.1, True, True), pv=(25.0, False, False))),
    (1, dict(load=(0.2, False, True), pv=(0.3, True, True), grid=(0.02, False, True))),
)


def make_noisy_forecast():
    out = {}
    for ci, (n, spec) in enumerate(NOISE_CASES):
        for seg, t0 in enumerate((0, 8740)):
            m = Microgrid.from_scenario(n)
            m.initial_step = t0
            for name, (std, inc, rel) in spec.items():
                getattr(m.modules, name)[0].set_forecaster(std, forecast_horizon=23, forecaster_increase_uncertainty=inc,
                                                           forecaster_relative_noise=rel)
            rng = np.random.default_rng(8000 + n)
            a = rng.random((19 if t0 else 12, n_act(m)))
            np.random.seed(500 + ci)
            reset_obs = flat_obs(m.reset())
            obs, rewards = [], []
            for row in a:
                o, r, _, _ = m.run(control_from_flat(m, row))
                obs.append(flat_obs(o)); rewards.append(r)
            key = f"c{ci}_{seg}"
            out[key + "_t0"], out[key + "_a"], out[key + "_reset_obs"] = np.array(t0), a, reset_obs
            out[key + "_o"], out[key + "_r"] = np.stack(obs), np.array(rewards)
            out[key + "_noise_std"] = np.array([np.mean(getattr(m.modules, name)[0].forecaster.noise_std) for name in spec])
            print("noisy", n, t0, reset_obs[:3], out[key + "_noise_std"])
    np.savez_compressed(os.path.join(HERE, "noisy_forecast.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["views", "noisy", "steps", "year", "discrete", "genset", "custom", "log", "rbc", "generator", "shaped"]
    if "shaped" in which:
        make_shaped()
    if "noisy" in which:
        make_noisy_forecast()
    if "views" in which:
        make_views()
    if "generator" in which:
        make_generator()
    if "rbc" in which:
        make_rbc()
    if "log" in which:
        make_log()
    if "steps" in which:
        make_pymgrid25_steps()
    if "year" in which:
        make_year()
    if "discrete" in which:
        make_discrete()
    if "genset" in which:
        make_genset_machine()
    if "custom" in which:
        make_custom()
