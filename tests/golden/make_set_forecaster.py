"""Microgrid.set_forecaster (microgrid/microgrid.py:477-546) on the LIVE reference -> tests/golden/set_forecaster.npz.
Build container only:  python tests/golden/make_set_forecaster.py

BASELINE config 4 is "DiscreteMicrogridEnv with forecast_horizon=24": `set_forecaster('oracle', forecast_horizon=24)` on a
built microgrid.  Recorded for pymgrid25 scenarios 0, 1, 2 (the fused module set) and for the quick-start notebook's
two-battery grid (the composed path): a few steps with the constructed forecaster, set_forecaster('oracle', 24), steps,
set_forecaster({name: ...}) -- a silent no-op in the reference (the dict branch calls the method on the module LIST and
swallows the AttributeError) --, steps, set_forecaster(None), steps; rewards, observations per phase, the final log.
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
from pymgrid.modules import BatteryModule, GridModule, LoadModule, RenewableModule  # noqa: E402

PHASES = (("keep", 4), ("oracle24", 4), ("dict", 2), ("none", 3))
SORTED = ("battery", "genset", "grid", "load", "pv")


def apply(m, phase, names):
    if phase == "oracle24":
        m.set_forecaster("oracle", forecast_horizon=24)
    elif phase == "dict":
        m.set_forecaster({names[0]: None}, forecast_horizon=7)
    elif phase == "none":
        m.set_forecaster(None)


def flow(m, rng, names, data, prefix):
    for phase, n in PHASES:
        apply(m, phase, names)
        rewards, rows, acts = [], [], []
        for _ in range(n):
            a = {name: [rng.random(2) if name == "genset" else rng.random() for _ in mods] for name, mods in m.controllable.iterdict()}
            obs, r, d, _ = m.run(a)
            acts.append(np.concatenate([np.atleast_1d(x) for name in a for x in a[name]]))
            rewards.append(r)
            rows.append(np.concatenate([np.asarray(x).ravel() for name in sorted(obs) for x in obs[name]]))
        data[f"{prefix}_{phase}_actions"], data[f"{prefix}_{phase}_rewards"] = np.array(acts), np.array(rewards)
        data[f"{prefix}_{phase}_obs"] = np.array(rows)
        if phase == "dict":     # the longest horizon so far: later a shorter one makes the reference's get_log() raise
            log = m.get_log()
            data[f"{prefix}_log_columns"] = np.array(json.dumps([list(c) for c in log.columns]))
            data[f"{prefix}_log_values"] = log.to_numpy(dtype=np.float64)
    try:
        m.get_log()
        data[f"{prefix}_final_log_raises"] = np.array("")
    except ValueError as exc:   # columns that stopped being logged are shorter than the index (utils/logger.py:18-28)
        data[f"{prefix}_final_log_raises"] = np.array(type(exc).__name__)


def quickstart_modules(ns):
    rng = np.random.default_rng(0)
    load, pv = 100 + 100 * rng.random(200), 200 * rng.random(200)
    return [ns["BatteryModule"](10, 100, 50, 50, 0.9, init_soc=0.2), ns["BatteryModule"](10, 1000, 10, 10, 0.7, init_soc=0.2),
            ("pv", ns["RenewableModule"](time_series=pv)), ns["LoadModule"](time_series=load),
            ns["GridModule"](100, 100, [0.2, 0.1, 0.5] * np.ones((200, 3)))]


def main():
    data = {}
    for n in (0, 1, 2):
        flow(Microgrid.from_scenario(n), np.random.default_rng(40 + n), ["load"], data, f"s{n}")
    flow(Microgrid(quickstart_modules(globals())), np.random.default_rng(50), ["load"], data, "quick")
    path = os.path.join(HERE, "set_forecaster.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), {k: v.shape for k, v in data.items() if k.endswith("_obs")})


if __name__ == "__main__":
    main()
