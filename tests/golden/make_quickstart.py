"""The reference's quick-start notebook (notebooks/quick-start.ipynb, cells 1-43) run against the LIVE reference
-> tests/golden/quickstart.npz.  Build container only:  python tests/golden/make_quickstart.py

Two batteries, a renamed renewable ('pv'), a load and a three-column grid; one hand-computed control step
(cells 23-32), then ten `microgrid.run(microgrid.sample_action(strict_bound=True))` steps (cell 41) under
np.random.seed(0).  tests/test_compose_host.py / tests/test_zz_gpu_compose.py replay the same cells against
pymgrid_b200 and must produce the same log, value for value.
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


def notebook(Microgrid, BatteryModule, LoadModule, RenewableModule, GridModule, **microgrid_kw):
    """the notebook's cells; returns what it displays"""
    np.random.seed(0)
    small_battery = BatteryModule(min_capacity=10, max_capacity=100, max_charge=50, max_discharge=50, efficiency=0.9, init_soc=0.2)
    large_battery = BatteryModule(min_capacity=10, max_capacity=1000, max_charge=10, max_discharge=10, efficiency=0.7, init_soc=0.2)
    load_ts = 100 + 100 * np.random.rand(24 * 90)
    pv_ts = 200 * np.random.rand(24 * 90)
    load = LoadModule(time_series=load_ts)
    pv = RenewableModule(time_series=pv_ts)
    grid_ts = [0.2, 0.1, 0.5] * np.ones((24 * 90, 3))
    grid = GridModule(max_import=100, max_export=100, time_series=grid_ts)
    microgrid = Microgrid([small_battery, large_battery, ('pv', pv), load, grid], **microgrid_kw)
    out = {"repr": repr(microgrid), "same_object": microgrid.modules.grid is microgrid.modules['grid'],
           "controllable": list(microgrid.controllable.to_dict().keys()), "empty_action": microgrid.get_empty_action()}
    microgrid.reset()
    ss = microgrid.state_series()
    out["state_series_index"], out["state_series"] = [list(map(str, i)) for i in ss.index], ss.to_numpy(dtype=np.float64)
    load_now = -1.0 * microgrid.modules.load.item().current_load
    pv_now = microgrid.modules.pv.item().current_renewable
    net_load = load_now + pv_now
    if net_load > 0:
        net_load = 0.0
    battery_0_discharge = min(-1 * net_load, microgrid.modules.battery[0].max_production)
    net_load += battery_0_discharge
    battery_1_discharge = min(-1 * net_load, microgrid.modules.battery[1].max_production)
    net_load += battery_1_discharge
    grid_import = min(-1 * net_load, microgrid.modules.grid.item().max_production)
    control = {"battery": [battery_0_discharge, battery_1_discharge], "grid": [grid_import]}
    out["control"] = np.array([battery_0_discharge, battery_1_discharge, grid_import])
    obs, reward, done, info = microgrid.run(control, normalized=False)
    out["reward0"], out["done0"] = reward, done
    out["obs0"] = np.concatenate([np.asarray(x).ravel() for name in obs for x in obs[name]])
    out["obs0_keys"] = list(obs.keys())
    sampled = []
    for _ in range(10):
        a = microgrid.sample_action(strict_bound=True)
        sampled.append([a["battery"][0], a["battery"][1], a["grid"][0]])
        microgrid.run(a)
    out["sampled"] = np.array(sampled, dtype=np.float64)
    log = microgrid.log
    out["log_columns"], out["log"] = [list(c) for c in log.columns], log.to_numpy(dtype=np.float64)
    import pandas as pd
    out["load_slice_cols"] = [list(c) for c in log.loc[:, pd.IndexSlice['load', 0, :]].columns]
    out["battery_slice_cols"] = [list(map(str, c)) for c in log.loc[:, 'battery'].columns]
    plot = log[[('load', 0, 'load_met'), ('pv', 0, 'renewable_used'), ('balancing', 0, 'loss_load')]].droplevel(axis=1, level=1)
    out["plot_frame"] = plot.to_numpy(dtype=np.float64)
    return out


def main():
    from pymgrid import Microgrid
    from pymgrid.modules import BatteryModule, GridModule, LoadModule, RenewableModule
    out = notebook(Microgrid, BatteryModule, LoadModule, RenewableModule, GridModule)
    data = {k: (v if isinstance(v, np.ndarray) else np.array(json.dumps(v))) for k, v in out.items()}
    path = os.path.join(HERE, "quickstart.npz")
    np.savez_compressed(path, **data)
    print(out["repr"], out["controllable"], out["empty_action"], out["reward0"], out["log"].shape)


if __name__ == "__main__":
    main()
