"""Golden vectors for priority lists on COMPOSED microgrids from the LIVE, UNMODIFIED reference
-> tests/golden/compose_discrete.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_compose_discrete.py

DiscreteMicrogridEnv (envs/discrete/discrete.py:60-143) and RuleBasedControl (algos/rbc/rbc.py:7-140) work on any module
list through PriorityListAlgo (algos/priority_list/priority_list.py:15-167).  Recorded per composition: the action table
(with and without remove_redundant_gensets), for a sequence of random discrete actions the control each one expands to,
reward, done, observation and state; RuleBasedControl's automatically sorted priority list and the rewards of its run.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import make_compose as MC  # noqa: E402  (loads the reference)
from make_compose import Case, Microgrid, grid_series, split  # noqa: E402
from pymgrid.algos import RuleBasedControl  # noqa: E402
from pymgrid.envs import DiscreteMicrogridEnv  # noqa: E402


def make_cases():
    rng = np.random.default_rng(777)
    T = 50
    load, pv = 120 * rng.random(T), 150 * np.clip(rng.random(T) - 0.25, 0, None)
    load[7:10] = 0.0
    pv[8:12] = 0.0
    bat = dict(min_capacity=10.0, max_capacity=100.0, max_charge=40.0, max_discharge=35.0, efficiency=0.9,
               battery_cost_cycle=0.02, init_soc=0.5)
    bat2 = dict(min_capacity=0.0, max_capacity=55.5, max_charge=60.0, max_discharge=11.0, efficiency=1.0,
                battery_cost_cycle=0.3, init_charge=20.0)
    gen = dict(running_min_production=10.0, running_max_production=60.0, genset_cost=0.4, co2_per_unit=2.0, cost_per_unit_co2=0.1)
    cases = []
    cases.append(Case("two_batteries_grid", n_norm=40, n_unnorm=0)
                 .add("LoadModule", ts=load).add("RenewableModule", ts=pv)
                 .add("BatteryModule", **bat).add("BatteryModule", **bat2)
                 .add("GridModule", ts=grid_series(rng, T, weak=True), max_import=60.0, max_export=30.0, cost_per_unit_co2=0.1))
    c = Case("two_gensets_battery", n_norm=40, n_unnorm=0)
    for ts in split(rng, load + 1e-3, 2):
        c.add("LoadModule", ts=ts, forecaster="oracle", forecast_horizon=2)
    for ts in split(rng, pv + 1e-3, 2):
        c.add("RenewableModule", ts=ts, forecaster="oracle", forecast_horizon=3)
    c.add("GensetModule", running_min_production=0.0, running_max_production=45.0, genset_cost=0.55)
    c.add("GensetModule", start_up_time=2, wind_down_time=1, init_start_up=False, **gen)
    c.add("BatteryModule", **bat)
    cases.append(c)
    cases.append(Case("genset_grid_no_battery", n_norm=40, n_unnorm=0)
                 .add("LoadModule", ts=load).add("RenewableModule", ts=pv)
                 .add("GensetModule", start_up_time=1, wind_down_time=1, **gen)
                 .add("GridModule", ts=grid_series(rng, T, cols=3), max_import=40.0, max_export=0.0))
    cases.append(Case("one_of_each_own_horizons", n_norm=40, n_unnorm=0)
                 .add("LoadModule", ts=load, forecaster="oracle", forecast_horizon=3)
                 .add("RenewableModule", name="pv", ts=pv)
                 .add("BatteryModule", **bat).add("GensetModule", **gen)
                 .add("GridModule", ts=grid_series(rng, T), max_import=80.0, max_export=50.0, forecaster="oracle", forecast_horizon=5))
    return cases


def element_rows(pl):
    return [[el.module[0], int(el.module[1]), int(el.module_actions), int(el.action)] for el in pl]


def record(case, seed):
    rng = np.random.default_rng(seed)
    mods = case.build().modules.to_tuples()            # the same modules, named
    out = {"spec": np.array(case.spec())}
    for flag in (False, True):
        env = DiscreteMicrogridEnv(case.build().modules.to_tuples(), add_unbalanced_module=False, remove_redundant_gensets=flag)
        out[f"table_{int(flag)}"] = np.array(json.dumps([element_rows(pl) for pl in env.actions_list]))
    env = DiscreteMicrogridEnv(mods, add_unbalanced_module=False, remove_redundant_gensets=False)
    order = MC.listing(env)
    out["names"] = np.array(json.dumps([[n, j, type(x).__name__] for n, j, x in order]))
    acts, controls, rewards, dones, obs_rows, states = [], [], [], [], [], []
    env.reset()
    for k in range(case.n_norm):
        a = int(rng.integers(0, env.action_space.n))
        control = env._get_action(a)
        controls.append(MC.control_row(control, env))
        obs, reward, done, info = Microgrid.run(env, control, normalized=False)
        acts.append(a); rewards.append(reward); dones.append(bool(done))
        obs_rows.append(MC.flat_obs(obs, order)); states.append(MC.state_vec(order))
    out["actions"] = np.array(acts, dtype=np.int32)
    out["controls"] = np.array(controls, dtype=np.float64)
    out["rewards"], out["dones"] = np.array(rewards), np.array(dones, dtype=np.uint8)
    out["obs"], out["states"] = np.array(obs_rows), np.array(states, dtype=np.float64)
    # rule-based control on a fresh microgrid
    rbc = RuleBasedControl(case.build())
    out["rbc_list"] = np.array(json.dumps(element_rows(rbc.priority_list)))
    log = rbc.run(max_steps=30)
    out["rbc_rewards"] = log[("balance", 0, "reward")].to_numpy(dtype=np.float64)
    out["rbc_log_columns"] = np.array(json.dumps([list(c) for c in log.columns]))
    out["rbc_log_values"] = log.to_numpy(dtype=np.float64)
    return out


def main():
    data = {}
    cases = make_cases()
    for i, case in enumerate(cases):
        rec = record(case, 500 + i)
        for k, v in rec.items():
            data[f"c{i}_{k}"] = v
        for j, ts in enumerate(case.series):
            data[f"c{i}_ts{j}"] = ts
        print(i, case.label, "lists", len(json.loads(str(rec["table_0"]))), len(json.loads(str(rec["table_1"]))),
              "rbc", json.loads(str(rec["rbc_list"])), "sum", rec["rbc_rewards"].sum())
    data["n_cases"] = np.array(len(cases))
    path = os.path.join(HERE, "compose_discrete.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
