"""Scenario ingestion: pymgrid25 benchmark grids -> `MicrogridParams`.

Two sources, same result (tests/test_scenario.py checks they agree bit for bit):

* `read_reference_scenario(dir)`: our own reader of the reference's on-disk scenario format
  (`!Microgrid` YAML + `!NDArray` csv.gz side files; reference: utils/serialize.py:91-112,
  data/scenario/pymgrid25/microgrid_N/microgrid_N.yaml).  No reference code is imported.
* `load_pymgrid25(n)`: the bundled `data/pymgrid25.npz`, produced from those files by
  `tools/convert_pymgrid25.py`, so that the benchmark scenarios exist on machines that do not have the
  reference checkout (the GPU boxes).
"""
import gzip
import io
import os
from functools import lru_cache

import numpy as np
import yaml

from .params import BatteryParams, GensetParams, GridParams, MicrogridParams

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
N_PYMGRID25 = 25


class _Tagged(dict):
    tag = None


class _ScenarioLoader(yaml.SafeLoader):
    # private constructor table: only the standard YAML tags, so that tag handlers another library (e.g. the
    # reference itself) registered on yaml.SafeLoader never leak into this reader
    yaml_constructors = {k: v for k, v in yaml.SafeLoader.yaml_constructors.items()
                         if k is None or str(k).startswith("tag:yaml.org")}


def _construct_tagged(loader, suffix, node):
    if isinstance(node, yaml.MappingNode):
        out = _Tagged(loader.construct_mapping(node, deep=True))
        out.tag = suffix
        return out
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return ("!" + suffix, loader.construct_scalar(node))


_ScenarioLoader.add_multi_constructor("!", _construct_tagged)


def _read_series(path):
    """csv(.gz) with an index column and a header row, as written by the reference's NDArray dumper."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        text = f.read()
    arr = np.genfromtxt(io.StringIO(text), delimiter=",", skip_header=1, dtype=np.float64)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 2)
    return np.ascontiguousarray(arr[:, 1:])


def _resolve(value, base_dir):
    if isinstance(value, tuple) and value and value[0] == "!NDArray":
        return _read_series(os.path.join(base_dir, value[1]))
    return value


def read_reference_scenario(yaml_path):
    """Parse one `microgrid_N.yaml` of the reference's scenario format into `MicrogridParams`."""
    base_dir = os.path.dirname(os.path.abspath(yaml_path))
    with open(yaml_path) as f:
        doc = yaml.load(f, Loader=_ScenarioLoader)
    modules = {}
    for name, mod in doc["modules"]:
        cls_params = {k: _resolve(v, base_dir) for k, v in mod["cls_params"].items()}
        modules[name] = (mod.tag, cls_params, mod.get("state", {}))

    def by_tag(tag):
        hits = [m for m in modules.values() if m[0] == tag]
        if len(hits) > 1:
            raise NotImplementedError(f"more than one {tag} per microgrid is outside the batched engine's scope")
        return hits[0] if hits else None

    load, pv = by_tag("LoadModule"), by_tag("RenewableModule")
    bat, gen, grid, unb = by_tag("BatteryModule"), by_tag("Genset"), by_tag("GridModule"), by_tag("UnbalancedEnergyModule")
    if load is None or pv is None or bat is None:
        raise NotImplementedError("the batched engine needs a load, a renewable and a battery module")
    bp, bs = bat[1], bat[2]
    battery = BatteryParams(min_capacity=float(bp["min_capacity"]), max_capacity=float(bp["max_capacity"]),
                            max_charge=float(bp["max_charge"]), max_discharge=float(bp["max_discharge"]),
                            efficiency=float(bp["efficiency"]), battery_cost_cycle=float(bp["battery_cost_cycle"]))
    # reference restores soc first, then current_charge (serializable_state_attributes order), so the
    # serialised charge wins; without a state block fall back to the constructor rule (battery_module.py:97-107)
    if "current_charge" in bs:
        battery.current_charge = float(bs["current_charge"])
    elif bp.get("init_charge") is not None:
        battery.current_charge = float(bp["init_charge"])
    else:
        battery.soc = float(bp["init_soc"])
        battery.current_charge = battery.soc * battery.max_capacity
    genset = None
    if gen is not None:
        gp, gs = gen[1], gen[2]
        genset = GensetParams.with_init(
            init_start_up=bool(gp.get("init_start_up", True)),
            running_min_production=float(gp["running_min_production"]),
            running_max_production=float(gp["running_max_production"]),
            genset_cost=float(gp["genset_cost"]), co2_per_unit=float(gp.get("co2_per_unit", 0.0)),
            cost_per_unit_co2=float(gp.get("cost_per_unit_co2", 0.0)),
            start_up_time=int(gp.get("start_up_time", 0)), wind_down_time=int(gp.get("wind_down_time", 0)),
            allow_abortion=bool(gp.get("allow_abortion", True)))
        if "_current_status" in gs:
            genset.current_status = int(gs["_current_status"])
            genset.goal_status = int(gs["_goal_status"])
            genset.steps_until_up = int(gs["_steps_until_up"])
            genset.steps_until_down = int(gs["_steps_until_down"])
    gridp = None
    if grid is not None:
        gp = grid[1]
        gridp = GridParams(max_import=float(gp["max_import"]), max_export=float(gp["max_export"]),
                           time_series=gp["time_series"], cost_per_unit_co2=float(gp.get("cost_per_unit_co2", 0.0)))
    lp = load[1]
    horizon = int(lp.get("forecast_horizon", 23)) if lp.get("forecaster") is not None else 0
    if lp.get("forecaster") not in (None, "oracle"):
        raise NotImplementedError("only the oracle forecaster (and None) is on the batched path (SURVEY.md section 2)")
    return MicrogridParams(
        battery=battery, genset=genset, grid=gridp,
        load_ts=lp["time_series"][:, 0], pv_ts=pv[1]["time_series"][:, 0],
        loss_load_cost=float(unb[1]["loss_load_cost"]) if unb else 10.0,
        overgeneration_cost=float(unb[1]["overgeneration_cost"]) if unb else 2.0,
        forecast_horizon=horizon, initial_step=int(doc.get("initial_step", 0)),
        final_step=int(doc.get("final_step", -1)), current_step=int(load[2].get("_current_step", 0)),
        name=os.path.splitext(os.path.basename(yaml_path))[0])


def reference_scenario_path(n, data_root):
    return os.path.join(data_root, "scenario", "pymgrid25", f"microgrid_{n}", f"microgrid_{n}.yaml")


@lru_cache(maxsize=1)
def _bundle():
    path = os.path.join(_DATA_DIR, "pymgrid25.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing; run tools/convert_pymgrid25.py where the reference data is present")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def load_pymgrid25(n):
    """`Microgrid.from_scenario(n)` equivalent (reference: microgrid.py:958-980): benchmark grid n of pymgrid25."""
    if not 0 <= n < N_PYMGRID25:
        raise ValueError(f"pymgrid25 has scenarios 0..{N_PYMGRID25 - 1}")
    z = _bundle()
    s = z["scalars"][n]
    (min_cap, max_cap, max_ch, max_dis, eff, cc, charge, has_gen, rmin, rmax, gcost, co2u, gco2c, U, D, abort,
     cs, gs, up, dn, has_grid, gimp, gexp, grco2c, llc, ogc, H, init_step, final_step) = s
    battery = BatteryParams(min_cap, max_cap, max_ch, max_dis, eff, cc, charge)
    genset = None
    if has_gen:
        genset = GensetParams(rmin, rmax, gcost, co2u, gco2c, int(U), int(D), bool(abort), int(cs), int(gs), int(up), int(dn))
    grid = None
    if has_grid:
        gi = int(z["grid_index"][n])
        ts = np.stack([z["grid_import_price"][gi], z["grid_export_price"][gi],
                       z["co2_profiles"][int(z["grid_co2_profile"][gi])],
                       np.unpackbits(z["grid_status_bits"][gi])[:z["load"].shape[1]].astype(np.float64)], axis=1)
        grid = GridParams(gimp, gexp, ts, grco2c)
    return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=z["load"][n], pv_ts=z["pv"][n],
                           loss_load_cost=llc, overgeneration_cost=ogc, forecast_horizon=int(H),
                           initial_step=int(init_step), final_step=int(final_step), name=f"microgrid_{n}")


SCALAR_FIELDS = 29


def pack_scalars(p):
    """Inverse of the unpacking in `load_pymgrid25` (used by tools/convert_pymgrid25.py)."""
    b, g, gr = p.battery, p.genset, p.grid
    row = [b.min_capacity, b.max_capacity, b.max_charge, b.max_discharge, b.efficiency, b.battery_cost_cycle,
           b.current_charge, float(g is not None)]
    if g is not None:
        row += [g.running_min_production, g.running_max_production, g.genset_cost, g.co2_per_unit,
                g.cost_per_unit_co2, g.start_up_time, g.wind_down_time, float(g.allow_abortion),
                g.current_status, g.goal_status, g.steps_until_up, g.steps_until_down]
    else:
        row += [0.0] * 12
    row += [float(gr is not None)]
    row += [gr.max_import, gr.max_export, gr.cost_per_unit_co2] if gr is not None else [0.0] * 3
    row += [p.loss_load_cost, p.overgeneration_cost, p.forecast_horizon, p.initial_step, p.final_step]
    assert len(row) == SCALAR_FIELDS
    return np.array(row, dtype=np.float64)
