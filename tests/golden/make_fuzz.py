"""Randomised-parameter golden vectors from the LIVE, UNMODIFIED reference -> tests/golden/fuzz.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_fuzz.py

pymgrid25 and the hand-built grids of custom.npz share a handful of parameter values (efficiency 0.9, instant
gensets, export price 0 ...).  This fixture draws every constructor argument of the modules on the hot path at
random -- including the corner values the reference treats specially (min_capacity 0, running_min_production 0
or equal to running_max_production, max_export 0, three-column grid series, forecast horizon longer than what is
left of the series, initial_step > 0, final_step inside the series, zero load / zero PV stretches) -- and records,
per grid:

  continuous path  reset observation; normalised steps; unnormalised steps reaching beyond every module limit
                   (rewards, done, flat observation, info block, state after each step)
  discrete path    DiscreteMicrogridEnv action table, the controls each action expands to, rewards, observations
                   (priority-list expansion with slow gensets: next_max_production / next_min_production)
  rule based       RuleBasedControl's automatically sorted priority list and the rewards of its run

tests/test_oracle_vs_golden.py (C oracle) and tests/test_gpu_parity.py (CUDA path) must reproduce all of it bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import make_golden as G  # noqa: E402  (loads the reference through oracle/ref_loader)
from make_golden import (BatteryModule, DiscreteMicrogridEnv, GensetModule, GridModule, LoadModule, Microgrid,  # noqa: E402
                         RenewableModule)

N_GRIDS = 40
SPEC_COLS = ("T", "H", "initial_step", "final_step", "b_min", "b_max", "b_charge", "b_discharge", "b_eff", "b_cost",
             "b_init_soc", "has_gen", "g_min", "g_max", "g_cost", "g_co2", "g_cco2", "g_U", "g_D", "g_abort", "g_init",
             "has_grid", "r_imp", "r_exp", "r_cco2", "r_cols", "llc", "ogc")


def draw_spec(rng, i):
    T = int(rng.integers(30, 80))
    s = dict(T=T)
    s["H"] = int(rng.choice([0, 1, 2, 5, 11, 23, 40]))
    s["initial_step"] = int(rng.choice([0, 0, 3, 7]))
    s["final_step"] = int(rng.choice([-1, -1, int(rng.integers(T // 2 + 8, T))]))
    s["b_max"] = float(rng.uniform(20, 400))
    s["b_min"] = float(rng.choice([0.0, rng.uniform(0.05, 0.5) * s["b_max"]]))
    s["b_charge"] = float(rng.uniform(0.05, 1.2) * s["b_max"])
    s["b_discharge"] = float(rng.uniform(0.05, 1.2) * s["b_max"])
    s["b_eff"] = float(rng.choice([1.0, rng.uniform(0.5, 0.999)]))
    s["b_cost"] = float(rng.choice([0.0, rng.uniform(0.001, 0.8)]))
    s["b_init_soc"] = float(rng.uniform(s["b_min"] / s["b_max"], 1.0))
    arch = i % 4                 # 0 genset + grid, 1 genset, 2 grid, 3 neither
    s["has_gen"] = int(arch in (0, 1))
    s["has_grid"] = int(arch in (0, 2))
    s["g_max"] = float(rng.uniform(20, 200))
    s["g_min"] = float(rng.choice([0.0, s["g_max"], rng.uniform(0.05, 0.9) * s["g_max"]], p=[0.25, 0.1, 0.65]))
    s["g_cost"] = float(rng.uniform(0.0, 1.0))
    s["g_co2"] = float(rng.choice([0.0, rng.uniform(0.1, 3.0)]))
    s["g_cco2"] = float(rng.choice([0.0, rng.uniform(0.01, 0.5)]))
    s["g_U"], s["g_D"] = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    s["g_abort"], s["g_init"] = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    s["r_imp"] = float(rng.uniform(10, 300))
    s["r_exp"] = float(rng.choice([0.0, rng.uniform(10, 300)]))
    s["r_cco2"] = float(rng.choice([0.0, rng.uniform(0.01, 0.5)]))
    s["r_cols"] = int(rng.choice([3, 4]))
    s["llc"], s["ogc"] = float(rng.uniform(0.5, 20)), float(rng.uniform(0.0, 5))
    return s


def draw_series(rng, s):
    T = s["T"]
    load = rng.uniform(5, 250) * rng.random(T)
    pv = rng.uniform(5, 250) * np.clip(rng.random(T) - 0.3, 0, None)
    k = int(rng.integers(0, T - 6))
    if rng.random() < 0.5:
        load[k:k + 4] = 0.0                      # nothing to serve: remaining_load <= 0 branch of the priority list
    if rng.random() < 0.5:
        pv[k + 1:k + 6] = 0.0
    grid = np.stack([rng.uniform(0.05, 0.9, T), rng.choice([0.0, 1.0]) * rng.uniform(0.0, 0.4, T), rng.uniform(0.0, 0.6, T),
                     (rng.random(T) > rng.choice([0.0, 0.3])).astype(np.float64)], axis=1)
    return load, pv, grid[:, :s["r_cols"]]


def build(s, load, pv, grid_ts):
    ts_kw = dict(forecaster="oracle" if s["H"] > 0 else None, forecast_horizon=s["H"] if s["H"] > 0 else 23,
                 final_step=s["final_step"])
    mods = [LoadModule(time_series=load, **ts_kw), ("pv", RenewableModule(time_series=pv, **ts_kw))]
    if s["has_gen"]:
        mods.append(GensetModule(running_min_production=s["g_min"], running_max_production=s["g_max"],
                                 genset_cost=s["g_cost"], co2_per_unit=s["g_co2"], cost_per_unit_co2=s["g_cco2"],
                                 start_up_time=s["g_U"], wind_down_time=s["g_D"], allow_abortion=bool(s["g_abort"]),
                                 init_start_up=bool(s["g_init"])))
    mods.append(BatteryModule(min_capacity=s["b_min"], max_capacity=s["b_max"], max_charge=s["b_charge"],
                              max_discharge=s["b_discharge"], efficiency=s["b_eff"], battery_cost_cycle=s["b_cost"],
                              init_soc=s["b_init_soc"]))
    if s["has_grid"]:
        mods.append(GridModule(max_import=s["r_imp"], max_export=s["r_exp"], time_series=grid_ts,
                               cost_per_unit_co2=s["r_cco2"], **ts_kw))
    m = Microgrid(mods, loss_load_cost=s["llc"], overgeneration_cost=s["ogc"])
    if s["initial_step"]:
        m.initial_step = s["initial_step"]      # the setter moves every module (microgrid.py:644-660); reset() jumps there
        m.reset()
    return m


def unnormalised(m, rng, n):
    out = []
    for _ in range(n):
        row = []
        if hasattr(m.modules, "genset"):
            g = m.modules.genset[0]
            row += [float(rng.integers(0, 2)) if rng.random() < 0.7 else rng.random(),
                    rng.uniform(0.0, 1.3) * g.running_max_production]
        b = m.modules.battery[0]
        row += [rng.uniform(-1.5, 1.5) * max(b.max_charge, b.max_discharge)]
        if hasattr(m.modules, "grid"):
            gr = m.modules.grid[0]
            row += [rng.uniform(-1.3, 1.3) * max(gr.max_import, gr.max_export)]
        out.append(row)
    return np.array(out).reshape(n, -1)


def run_segment(m, actions, normalized=True):
    """make_golden.run_segment that survives the reference's own assertions: returns the arrays of the steps that
    completed and the index of the step whose `run` raised AssertionError (-1: none).  Seen in practice: the battery's
    charge lands one ulp above max_capacity, max_consumption turns negative and base_module.py:272 fires on the next
    charge request (MG_FLAG_NEGATIVE_ABSORB on the batched path)."""
    rewards, dones, obs, infos, states = [], [], [], [], []
    err = -1
    for k, a in enumerate(actions):
        try:
            o, r, d, info = m.run(G.control_from_flat(m, a), normalized=normalized)
        except AssertionError:
            err = k
            break
        rewards.append(r); dones.append(d); obs.append(G.flat_obs(o)); infos.append(G.info_vec(info, m)); states.append(G.state_vec(m))
    width = len(G.flat_obs(m.state_dict())) if not obs else 0
    return (np.array(rewards, dtype=np.float64), np.array(dones, dtype=np.uint8),
            np.stack(obs) if obs else np.zeros((0, width)), np.stack(infos) if infos else np.zeros((0, G.INFO_COLS)),
            np.stack(states) if states else np.zeros((0, 6))), err


OVERFULL = dict(load=np.full(12, 1.0), pv=np.full(12, 5.0), min_capacity=0.0, max_capacity=100.0, max_charge=50.0,
                max_discharge=50.0, efficiency=0.9, battery_cost_cycle=0.01, max_import=30.0, max_export=30.0,
                grid_ts=np.stack([np.full(12, 0.2), np.full(12, 0.1), np.full(12, 0.3), np.ones(12)], axis=1))


def overfull_battery():
    """A battery whose charge sits one ulp ABOVE max_capacity (rounding of charge += e * efficiency gets there, see
    run_segment): max_consumption is negative.  The reference then refuses (a) any continuous charge request
    (AssertionError base_module.py:272) and (b) any priority list that reaches the battery while there is surplus energy
    (AssertionError priority_list.py:124); discharging works.  Recorded: which calls raised, and the step that still ran."""
    import traceback
    o = OVERFULL

    def build():
        m = Microgrid([LoadModule(time_series=o["load"], forecaster="oracle", forecast_horizon=3),
                       ("pv", RenewableModule(time_series=o["pv"], forecaster="oracle", forecast_horizon=3)),
                       BatteryModule(min_capacity=o["min_capacity"], max_capacity=o["max_capacity"], max_charge=o["max_charge"],
                                     max_discharge=o["max_discharge"], efficiency=o["efficiency"],
                                     battery_cost_cycle=o["battery_cost_cycle"], init_soc=1.0),
                       GridModule(max_import=o["max_import"], max_export=o["max_export"], time_series=o["grid_ts"],
                                  forecaster="oracle", forecast_horizon=3)], loss_load_cost=10.0, overgeneration_cost=1.0)
        m.modules.battery[0].current_charge = float(np.nextafter(o["max_capacity"], np.inf))
        return m

    def where(fn):
        try:
            fn()
        except AssertionError:
            tb = traceback.extract_tb(sys.exc_info()[2])[-1]
            return f"{os.path.basename(tb.filename)}:{tb.lineno}"
        return ""

    out = {"over_charge": np.array(np.nextafter(o["max_capacity"], np.inf))}
    m = build()
    assert m.modules.battery[0].max_consumption < 0
    out["over_continuous_charge_raised"] = np.array(where(lambda: m.run({"battery": [-10.0], "grid": [0.0]}, normalized=False)))
    m = build()
    obs, r, d, info = m.run({"battery": [5.0], "grid": [-9.0]}, normalized=False)      # discharging is fine
    out["over_discharge_reward"], out["over_discharge_obs"] = np.array(r), G.flat_obs(obs)
    out["over_discharge_state"] = G.state_vec(m)
    env = DiscreteMicrogridEnv.from_microgrid(build())
    mod, act = G.action_table(env)
    out["over_table_mod"], out["over_table_act"] = mod, act
    raised, rewards = [], []
    for a in range(env.action_space.n):
        env = DiscreteMicrogridEnv.from_microgrid(build())
        raised.append(where(lambda: rewards.append(env.step(a)[1])))
        if raised[-1]:
            rewards.append(np.nan)
    out["over_discrete_raised"], out["over_discrete_reward"] = np.array(raised), np.array(rewards, dtype=np.float64)
    for k in ("load", "pv", "grid_ts"):
        out[f"over_{k}"] = o[k]
    out["over_spec"] = np.array([o[k] for k in ("min_capacity", "max_capacity", "max_charge", "max_discharge", "efficiency",
                                                "battery_cost_cycle", "max_import", "max_export")])
    print("overfull battery:", out["over_continuous_charge_raised"], raised, float(out["over_discharge_reward"]))
    return out


def main():
    out = dict(n=np.array(N_GRIDS), spec_cols=np.array(SPEC_COLS))
    for i in range(N_GRIDS):
        rng = np.random.default_rng(9000 + i)
        s = draw_spec(rng, i)
        load, pv, grid_ts = draw_series(rng, s)
        tag = f"f{i}"
        out[f"{tag}_spec"] = np.array([s[k] for k in SPEC_COLS], dtype=np.float64)
        out[f"{tag}_load"], out[f"{tag}_pv"], out[f"{tag}_grid_ts"] = load, pv, grid_ts

        # ---- continuous path -----------------------------------------------------------------------------------------
        m = build(s, load, pv, grid_ts)
        last = (s["final_step"] if s["final_step"] > 0 else s["T"])           # steps t = initial .. last-1 are valid
        n_valid = last - s["initial_step"]
        n0 = n_valid // 2
        na = G.n_act(m)
        a0 = rng.random((n0, na))
        if s["has_gen"]:
            a0[:, 0] = np.where(rng.random(n0) < 0.6, np.round(a0[:, 0]), a0[:, 0])
        out[f"{tag}_reset_obs"] = G.flat_obs(m.reset())
        soc_before = [m.modules.battery[0].soc, m.state_dict()["battery"][0]["soc"]]
        seg0, err0 = run_segment(m, a0)
        # the battery's soc as the drop-in surface shows it: before any step (the init_soc it was constructed with), and
        # the logged column of the normalised segment (row 0 = state before the first update)
        log = m.get_log()
        out[f"{tag}_soc_before"] = np.array(soc_before, dtype=np.float64)
        out[f"{tag}_log_soc"] = log[("battery", 0, "soc")].values.astype(np.float64)
        out[f"{tag}_log_charge"] = log[("battery", 0, "current_charge")].values.astype(np.float64)
        out[f"{tag}_soc_after"] = np.array(m.modules.battery[0].soc)
        au = unnormalised(m, rng, n_valid - n0)
        segu, erru = run_segment(m, au, normalized=False) if err0 < 0 else (seg0, -1)
        for name, a, seg, err in (("n", a0, seg0, err0), ("u", au, segu, erru)):
            out[f"{tag}_{name}_a"] = a                      # all drawn actions; row `err` is the one the reference refused
            out[f"{tag}_{name}_err"] = np.array(err)
            for k, v in zip("rdois", seg):
                out[f"{tag}_{name}_{k}"] = v
        if err0 < 0 and erru < 0:
            assert segu[1][-1] == 1 and segu[1][:-1].sum() == 0 and seg0[1].sum() == 0

        # ---- discrete path -------------------------------------------------------------------------------------------
        env = DiscreteMicrogridEnv.from_microgrid(build(s, load, pv, grid_ts))
        mod, act = G.action_table(env)
        n_d = min(n_valid, 40)
        acts = rng.integers(0, env.action_space.n, n_d)
        out[f"{tag}_d_reset_obs"] = np.asarray(env.reset(), dtype=np.float64)
        ctrls, rewards, dones, obs = [], [], [], []
        for a in acts:                                      # an AssertionError here would abort the script: none seen
            ctrls.append(G.flat_control(env, env._get_action(int(a))))
            o, r, d, _ = env.step(int(a))
            rewards.append(r); dones.append(d); obs.append(np.asarray(o, dtype=np.float64))
        out[f"{tag}_d_table_mod"], out[f"{tag}_d_table_act"] = mod, act
        out[f"{tag}_d_actions"] = acts.astype(np.int32)
        out[f"{tag}_d_controls"] = np.stack(ctrls)
        out[f"{tag}_d_rewards"] = np.array(rewards)
        out[f"{tag}_d_dones"] = np.array(dones, dtype=np.uint8)
        out[f"{tag}_d_obs"] = np.stack(obs)
        out[f"{tag}_d_state"] = G.state_vec(env)

        # ---- rule based control --------------------------------------------------------------------------------------
        from pymgrid.algos import RuleBasedControl
        rbc = RuleBasedControl(build(s, load, pv, grid_ts))
        out[f"{tag}_rbc_list_mod"] = np.array([G.MOD_ID[el.module[0]] for el in rbc.priority_list], dtype=np.int8)
        out[f"{tag}_rbc_list_act"] = np.array([el.action for el in rbc.priority_list], dtype=np.int8)
        df = rbc.run()
        out[f"{tag}_rbc_rewards"] = df[("balance", 0, "reward")].values.astype(np.float64)
        out[f"{tag}_rbc_final_state"] = G.state_vec(rbc.microgrid)
        print(tag, {k: s[k] for k in ("T", "H", "initial_step", "final_step", "has_gen", "has_grid", "g_U", "g_D")},
              "err", err0, erru, "n_actions", env.action_space.n, "rbc", len(df), float(df[("balance", 0, "reward")].sum()))
    out.update(overfull_battery())
    np.savez_compressed(os.path.join(HERE, "fuzz.npz"), **out)


if __name__ == "__main__":
    main()
