"""Golden vectors for COMPOSED microgrids (any number of modules of each kind) from the LIVE, UNMODIFIED reference
-> tests/golden/compose.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_compose.py

pymgrid25 / MicrogridGenerator grids all have one load, one renewable, one battery and at most one genset and one grid
(the fused kernel's scope).  `Microgrid.run` itself (microgrid/microgrid.py:227-325) dispatches over ANY module list, and
the reference's own balance tests build such lists (tests/microgrid/test_microgrid.py:188-455: load + PV only, two loads,
two PVs, 3-9 of each).  This fixture records, for a set of compositions -- that test family plus grids with several
batteries / gensets / grids, per-module forecast horizons, custom names, no slack module -- what the reference returns:

  reset observation; per step reward, done, observation (modules in listing order), per-module info (provided,
  absorbed, co2 / curtailment, logged module reward, acted-as-sink), battery and genset state, the balance log; the
  full get_log() frame; the step at which the reference raised and the exception type.

Each case carries a JSON spec (constructor arguments, series stored beside it) from which the tests rebuild the same
microgrid with pymgrid_b200.modules' classes.  tests/test_compose_oracle.py (Python oracle), tests/test_compose_host.py
(the kernel source compiled for the host + the Python host layer) and tests/test_zz_gpu_compose.py (CUDA path) must
reproduce all of it bit for bit.
"""
import json
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
from pymgrid.modules import (BatteryModule, GensetModule, GridModule, LoadModule, RenewableModule,  # noqa: E402
                             UnbalancedEnergyModule)

CLASSES = dict(LoadModule=LoadModule, RenewableModule=RenewableModule, BatteryModule=BatteryModule,
               GensetModule=GensetModule, GridModule=GridModule, UnbalancedEnergyModule=UnbalancedEnergyModule)
INFO_SLOTS = 5      # provided, absorbed, extra (co2_production / curtailment), module reward, acted as sink
BALANCE_COLS = ("reward", "shaped_reward", "overall_provided_to_microgrid", "overall_absorbed_from_microgrid",
                "controllable_provided_to_microgrid", "controllable_absorbed_from_microgrid",
                "fixed_provided_to_microgrid", "fixed_absorbed_from_microgrid")


# ---- case construction ----------------------------------------------------------------------------------------------
class Case:
    def __init__(self, label, microgrid_kwargs=None, n_norm=25, n_unnorm=15, shaper=None, trajectory=None, reset_at=()):
        self.shaper, self.trajectory, self.reset_at = shaper, trajectory, tuple(reset_at)
        self.label, self.mods, self.series = label, [], []
        self.microgrid_kwargs = dict(microgrid_kwargs or {})
        self.n_norm, self.n_unnorm = n_norm, n_unnorm

    def add(self, cls, name=None, ts=None, **kwargs):
        entry = dict(cls=cls, name=name, kwargs=kwargs, ts=None)
        if ts is not None:
            entry["ts"] = len(self.series)
            self.series.append(np.asarray(ts, dtype=np.float64))
        self.mods.append(entry)
        return self

    def build(self):
        out = []
        for e in self.mods:
            kw = dict(e["kwargs"])
            if e["ts"] is not None:
                kw["time_series"] = self.series[e["ts"]]
            m = CLASSES[e["cls"]](**kw)
            out.append((e["name"], m) if e["name"] is not None else m)
        from tests.compose_cases import callable_kwargs
        return Microgrid(out, **self.microgrid_kwargs, **callable_kwargs(dict(shaper=self.shaper, trajectory=self.trajectory)))

    def spec(self):
        return json.dumps(dict(label=self.label, modules=self.mods, microgrid_kwargs=self.microgrid_kwargs,
                               n_norm=self.n_norm, n_unnorm=self.n_unnorm, shaper=self.shaper, trajectory=self.trajectory,
                               reset_at=list(self.reset_at)))


def split(rng, total, n):
    """n positive series summing to `total` (the way test_microgrid.py:374-396 splits them)"""
    parts, remaining = [], total.copy()
    for _ in range(n - 1):
        parts.append(remaining * (1 - rng.random(total.shape)))
        remaining = remaining - parts[-1]
    parts.append(remaining)
    return parts


def grid_series(rng, T, cols=4, weak=False):
    ts = np.stack([rng.uniform(0.05, 0.9, T), rng.uniform(0.0, 0.4, T), rng.uniform(0.0, 0.6, T),
                   (rng.random(T) > (0.3 if weak else 0.0)).astype(np.float64)], axis=1)
    return ts[:, :cols]


def make_cases():
    rng = np.random.default_rng(20240229)
    cases = []
    T = 100
    # -- the reference's own balance-test family (tests/microgrid/test_microgrid.py:188-421); default forecaster=None
    for label, n_l, n_p, mode in (("load_pv", 1, 1, "same"), ("load_excess_pv", 1, 1, "pv"), ("pv_excess_load", 1, 1, "load"),
                                  ("two_loads", 2, 1, "same"), ("two_pv", 1, 2, "same"), ("two_each", 2, 2, "same"),
                                  ("many_each", 7, 9, "same"), ("many_each_excess_pv", 5, 8, "pv"),
                                  ("many_each_excess_load", 9, 4, "load")):
        base = 10 * rng.random(T)
        load, pv = (base, base) if mode == "same" else ((base, base + 5 * rng.random(T)) if mode == "pv"
                                                         else (base + 5 * rng.random(T), base))
        c = Case(label, n_norm=T, n_unnorm=0)
        for ts in split(rng, load, n_l):
            c.add("LoadModule", ts=ts, raise_errors=(n_l <= 2))
        for ts in split(rng, pv, n_p):
            c.add("RenewableModule", ts=ts)
        cases.append(c)
    T = 60
    load, pv = 120 * rng.random(T), 90 * np.clip(rng.random(T) - 0.25, 0, None)
    bat = dict(min_capacity=10.0, max_capacity=100.0, max_charge=40.0, max_discharge=35.0, efficiency=0.9,
               battery_cost_cycle=0.02, init_soc=0.5)
    gen = dict(running_min_production=10.0, running_max_production=60.0, genset_cost=0.4, co2_per_unit=2.0,
               cost_per_unit_co2=0.1)
    # -- one of each, every time-series module with its own forecast horizon
    cases.append(Case("per_module_horizons")
                 .add("LoadModule", ts=load, forecaster="oracle", forecast_horizon=3)
                 .add("RenewableModule", ts=pv)
                 .add("BatteryModule", **bat)
                 .add("GensetModule", **gen)
                 .add("GridModule", ts=grid_series(rng, T), max_import=80.0, max_export=50.0, cost_per_unit_co2=0.1,
                      forecaster="oracle", forecast_horizon=5))
    # -- several of everything: a slow genset without abortion, a weak grid, a three-column grid
    c = Case("several_of_each", microgrid_kwargs=dict(loss_load_cost=7.5, overgeneration_cost=1.25))
    for ts in split(rng, load, 2):
        c.add("LoadModule", ts=ts, forecaster="oracle", forecast_horizon=2)
    for ts in split(rng, pv + 1e-3, 3):
        c.add("RenewableModule", ts=ts, forecaster="oracle", forecast_horizon=4)
    c.add("BatteryModule", **bat)
    c.add("BatteryModule", min_capacity=0.0, max_capacity=55.5, max_charge=60.0, max_discharge=11.0, efficiency=1.0,
          battery_cost_cycle=0.0, init_charge=20.0)
    c.add("GensetModule", **gen)
    c.add("GensetModule", running_min_production=0.0, running_max_production=33.0, genset_cost=0.7, start_up_time=2,
          wind_down_time=1, allow_abortion=False, init_start_up=False)
    c.add("GridModule", ts=grid_series(rng, T, weak=True), max_import=30.0, max_export=20.0, cost_per_unit_co2=0.2,
          forecaster="oracle", forecast_horizon=23)
    c.add("GridModule", ts=grid_series(rng, T, cols=3), max_import=25.0, max_export=0.0)
    cases.append(c)
    # -- no battery
    cases.append(Case("load_pv_genset").add("LoadModule", ts=load).add("RenewableModule", ts=pv)
                 .add("GensetModule", start_up_time=1, wind_down_time=2, **gen))
    cases.append(Case("load_pv_grid").add("LoadModule", ts=load, forecaster="oracle", forecast_horizon=1)
                 .add("RenewableModule", ts=pv, forecaster="oracle", forecast_horizon=1)
                 .add("GridModule", ts=grid_series(rng, T, weak=True), max_import=100.0, max_export=100.0))
    # -- no renewable, no controllable at all
    cases.append(Case("load_only").add("LoadModule", ts=load))
    # -- caller-chosen names (they decide dict keys, log columns and the gym-sorted observation order)
    cases.append(Case("custom_names", microgrid_kwargs=dict(add_unbalanced_module=False))
                 .add("LoadModule", name="zload", ts=load, forecaster="oracle", forecast_horizon=2)
                 .add("RenewableModule", name="PV", ts=pv, forecaster="oracle", forecast_horizon=2)
                 .add("RenewableModule", name="wind", ts=0.5 * pv[::-1], forecaster="oracle", forecast_horizon=2)
                 .add("BatteryModule", name="Abat", **bat)
                 .add("UnbalancedEnergyModule", name="unbalanced_energy", raise_errors=False, loss_load_cost=3.0,
                      overgeneration_cost=0.5))
    # -- no slack module: the reference raises RuntimeError as soon as the modules cannot balance (microgrid.py:321-323)
    # (balanced while pv covers the load exactly, then short)
    cases.append(Case("no_slack", microgrid_kwargs=dict(add_unbalanced_module=False), n_norm=12, n_unnorm=0)
                 .add("LoadModule", ts=load + 1.0)
                 .add("RenewableModule", ts=np.where(np.arange(T) < 6, 1.0, 0.5) * (load + 1.0)))
    # -- more than eight energy entries with batteries on both sides: np.sum switches to its unrolled pairwise order
    c = Case("pairwise_sums")
    for ts in split(rng, load, 6):
        c.add("LoadModule", ts=ts)
    for ts in split(rng, pv + 1e-3, 5):
        c.add("RenewableModule", ts=ts)
    for k in range(6):
        c.add("BatteryModule", min_capacity=float(k), max_capacity=40.0 + 7 * k, max_charge=9.0 + k,
              max_discharge=12.0 - k, efficiency=1.0 - 0.03 * k, battery_cost_cycle=0.01 * k, init_soc=0.6)
    for k in range(3):
        c.add("GridModule", ts=grid_series(rng, T), max_import=5.0 + k, max_export=4.0 + k)
    cases.append(c)
    # -- a window inside the series, run past final_step and past the end of the data (IndexError, load_module.py:111)
    T2 = 14
    cases.append(Case("past_the_end", microgrid_kwargs=dict(add_unbalanced_module=False), n_norm=T2 + 2, n_unnorm=0)
                 .add("LoadModule", ts=load[:T2], forecaster="oracle", forecast_horizon=4, initial_step=2, final_step=9)
                 .add("RenewableModule", ts=pv[:T2], forecaster="oracle", forecast_horizon=4, initial_step=2, final_step=9)
                 .add("BatteryModule", initial_step=2, **bat)
                 .add("UnbalancedEnergyModule", raise_errors=False, initial_step=2))
    # -- a Python reward shaper (the reference's TestMicrogridRewardShaping function) on a grid with everything
    c = Case("python_reward_shaper", shaper="marginal_cost_total", n_norm=20, n_unnorm=10)
    for ts in split(rng, load, 2):
        c.add("LoadModule", ts=ts)
    c.add("RenewableModule", ts=pv).add("BatteryModule", **bat).add("BatteryModule", **bat).add("GensetModule", **gen)
    c.add("GridModule", ts=grid_series(rng, T), max_import=30.0, max_export=20.0, cost_per_unit_co2=0.2)
    cases.append(c)
    # -- a trajectory function: every reset() moves the episode window (microgrid.py:221-225); run to done, reset, run on
    c = Case("trajectory_window", trajectory=[3, -40], n_norm=30, n_unnorm=0, reset_at=(0, 17))
    for ts in split(rng, load, 2):
        c.add("LoadModule", ts=ts, forecaster="oracle", forecast_horizon=2)
    c.add("RenewableModule", ts=pv).add("BatteryModule", **bat).add("BatteryModule", **bat)
    cases.append(c)
    return cases


# ---- recording ------------------------------------------------------------------------------------------------------
def listing(m):
    """[(name, index, module)] in the container's listing order (module_container.py:40-95)"""
    return [(name, j, mod) for name, mods in m.modules.iterdict() for j, mod in enumerate(mods)]


def flat_obs(obs, order):
    parts = [np.asarray(obs[name][j], dtype=np.float64).ravel() for name, j, _ in order]
    return np.concatenate(parts) if parts else np.zeros(0)


def control_for(m, rng, normalized):
    """{name: [action per module]} for every controllable module: U[0,1) normalised, or unnormalised values that reach
    beyond every limit"""
    control = {}
    for name, mods in m.controllable.iterdict():
        vals = []
        for mod in mods:
            n = mod.action_space.shape[0]
            if normalized:
                a = rng.random(n)
            else:
                lo, hi = np.atleast_1d(mod.min_act).astype(float), np.atleast_1d(mod.max_act).astype(float)
                a = lo - 0.3 * (hi - lo) + 1.6 * (hi - lo) * rng.random(n)
                if n == 2:
                    a[0] = rng.random()      # the genset goal is never denormalised and must stay in [0, 1]
                    a[1] = max(a[1], 0.0)    # a genset cannot act as a sink (AssertionError, genset_module.py:208)
            vals.append(a if n > 1 else float(a[0]))
        control[name] = vals
    return control


def control_row(control, m):
    row = []
    for name, mods in m.controllable.iterdict():
        for j, _ in enumerate(mods):
            row.extend(np.atleast_1d(control[name][j]).tolist())
    return row


def state_vec(order):
    v = []
    for _, _, mod in order:
        cls = type(mod).__name__
        if cls == "BatteryModule":
            v += [mod.current_charge, mod.soc]
        elif cls == "GensetModule":
            v += [float(mod.current_status), float(mod.goal_status), float(mod._steps_until_up), float(mod._steps_until_down)]
    return v


def record(case, seed):
    rng = np.random.default_rng(seed)
    m = case.build()
    order = listing(m)
    out = {"spec": np.array(case.spec()), "names": np.array(json.dumps([[n, j, type(x).__name__] for n, j, x in order]))}
    reset_obs = m.reset()
    out["reset_keys"] = np.array(json.dumps(list(reset_obs.keys())))
    out["obs_reset"] = flat_obs(reset_obs, order)
    out["state0"] = np.array(state_vec(order))
    rewards, dones, obs_rows, infos, states, actions, norm_flags, steps_after = [], [], [], [], [], [], [], []
    raised_at, raised_type = -1, ""
    run_keys = None
    reset_obs_rows = []
    for k in range(case.n_norm + case.n_unnorm):
        if k in case.reset_at:
            reset_obs_rows.append(flat_obs(m.reset(), order))
        normalized = k < case.n_norm
        control = control_for(m, rng, normalized)
        try:
            obs, reward, done, info = m.run(control, normalized=normalized)
        except Exception as exc:       # noqa: BLE001 -- recorded: the engine must flag / raise at the same step
            raised_at, raised_type = k, type(exc).__name__
            break
        run_keys = list(obs.keys())
        actions.append(control_row(control, m)); norm_flags.append(int(normalized))
        rewards.append(reward); dones.append(bool(done)); obs_rows.append(flat_obs(obs, order)); steps_after.append(m.current_step)
        row = np.zeros((len(order), INFO_SLOTS))
        for i, (name, j, mod) in enumerate(order):
            inf = info[name][j]
            row[i, 0] = inf.get("provided_energy", 0.0)
            row[i, 1] = inf.get("absorbed_energy", 0.0)
            row[i, 2] = inf.get("co2_production", inf.get("curtailment", 0.0))
            row[i, 3] = mod.log_dict()["reward"][-1]
            row[i, 4] = float("absorbed_energy" in inf)
        infos.append(row); states.append(state_vec(order))
    n = len(rewards)
    out["run_keys"] = np.array(json.dumps(run_keys))
    out["actions"] = np.array(actions, dtype=np.float64).reshape(n, len(actions[0]) if n else 0)
    out["normalized"] = np.array(norm_flags, dtype=np.int32)
    out["rewards"], out["dones"] = np.array(rewards, dtype=np.float64), np.array(dones, dtype=np.uint8)
    out["obs"] = np.array(obs_rows, dtype=np.float64).reshape(n, len(out["obs_reset"]))
    out["info"] = np.array(infos, dtype=np.float64).reshape(n, len(order), INFO_SLOTS)
    out["states"] = np.array(states, dtype=np.float64).reshape(n, len(out["state0"]))
    out["raised_at"], out["raised_type"] = np.array(raised_at), np.array(raised_type)
    out["reset_obs_rows"] = np.array(reset_obs_rows, dtype=np.float64).reshape(len(reset_obs_rows), len(out["obs_reset"]))
    out["steps_after"] = np.array(steps_after, dtype=np.int64)
    try:
        log = m.get_log()
        out["log_raises"] = np.array("")
    except ValueError as exc:
        # with a trajectory_func the reference indexes the frame from the MICROGRID's initial_step while the modules log
        # from the window's (microgrid.py:452-475 vs :221-225): get_log() raises a length mismatch.  Recorded, not mirrored.
        out["log_raises"] = np.array(type(exc).__name__)
        log = None
    if log is not None:
        out["log_columns"] = np.array(json.dumps([list(c) for c in log.columns]))
        out["log_values"] = log.to_numpy(dtype=np.float64)
        out["log_index"] = np.array(log.index, dtype=np.int64)
        out["balance"] = log["balance"][0][list(BALANCE_COLS)].to_numpy(dtype=np.float64) if n else np.zeros((0, 8))
    out["current_step"] = np.array(m.current_step)
    # microgrid-level views the reference's callers read
    out["empty_action"] = np.array(json.dumps({k: len(v) for k, v in m.get_empty_action().items()}))
    ss = m.state_series()
    out["state_series_index"] = np.array(json.dumps([list(map(str, i)) for i in ss.index]))
    out["state_series_values"] = ss.to_numpy(dtype=np.float64)
    return out


def main():
    data = {}
    cases = make_cases()
    for i, case in enumerate(cases):
        rec = record(case, 1000 + i)
        for k, v in rec.items():
            data[f"c{i}_{k}"] = v
        for j, ts in enumerate(case.series):
            data[f"c{i}_ts{j}"] = ts
        print(f"{i:2d} {case.label:24s} steps {len(rec['rewards']):3d} raised {int(rec['raised_at'])} {rec['raised_type']}"
              f"  obs {rec['obs'].shape[1] if len(rec['obs']) else 0}  log cols {rec['log_values'].shape[1] if 'log_values' in rec else rec['log_raises']}")
    data["n_cases"] = np.array(len(cases))
    path = os.path.join(HERE, "compose.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
