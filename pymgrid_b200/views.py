"""Pure host-side conversions between the engine's flat rows and the reference's Python types (numpy only, no GPU).

The reference returns nested dicts (`Microgrid.run`: microgrid/microgrid.py:227-325, microgrid/utils/step.py) and a
MultiIndex DataFrame log (`Microgrid.get_log`: microgrid.py:434-475).  The B=1 drop-in (`microgrid.py`) and the
env wrappers (`envs.py`) build those from the engine's `[D_obs]` observation row, `[16]` info row and flag word.
"""
import warnings
from collections import OrderedDict

import numpy as np

from ._cabi import FLAG_BATTERY_SINK, FLAG_EXCESS, FLAG_GRID_SINK

SORTED_ORDER = ("battery", "genset", "grid", "load", "pv")          # gym.spaces.Dict sorts keys
CONTAINER_ORDER = ("load", "pv", "genset", "battery", "grid")       # module listing order (module_container.py:355-413)
DISPATCH_ORDER = ("load", "genset", "battery", "grid", "pv", "unbalanced_energy")   # insertion order of run()'s dicts
CONTROL_ORDER = ("genset", "battery", "grid")                        # Microgrid.controllable iteration order


def module_widths(p):
    rows = 1 + p.forecast_horizon
    w = {"battery": 2, "load": rows, "pv": rows}
    if p.has_genset:
        w["genset"] = 4
    if p.has_grid:
        w["grid"] = 4 * rows
    return w


def obs_slices(p, order="gym_sorted"):
    """name -> slice of the flat observation row (envs/base/base.py:211-223 flatten order)."""
    w = module_widths(p)
    out, start = OrderedDict(), 0
    for name in (SORTED_ORDER if order == "gym_sorted" else CONTAINER_ORDER):
        if name in w:
            out[name] = slice(start, start + w[name])
            start += w[name]
    return out


def obs_row_to_dict(row, p, order="gym_sorted"):
    """Flat normalised observation -> the dict `Microgrid.run` / `reset` return ({name: [np.ndarray]})."""
    sl = obs_slices(p, order)
    out = OrderedDict()
    for name in DISPATCH_ORDER:
        if name == "unbalanced_energy":
            out[name] = [np.array([])]
        elif name in sl:
            out[name] = [module_obs(row[sl[name]])]
    return out


def module_obs(values):
    """one module's normalised state as the reference returns it: an ndarray, or a Python float when the state has a single
    element (a time-series module without a forecast) -- ModuleSpace.normalize hands back `.item()` for those"""
    arr = np.array(values, dtype=np.float64)
    return float(arr[0]) if arr.shape == (1,) else arr


def info_row_to_dict(info, flags, p):
    """Engine info row + flag word -> the reference's info dict (step.py:22-39; one list entry per module)."""
    f = int(flags)
    out = OrderedDict()
    out["load"] = [{"absorbed_energy": float(info[0])}]
    if p.has_genset:
        out["genset"] = [{"provided_energy": float(info[5]), "co2_production": float(info[6])}]
    out["battery"] = [{"absorbed_energy": float(info[8])} if f & FLAG_BATTERY_SINK else {"provided_energy": float(info[7])}]
    if p.has_grid:
        out["grid"] = [{"absorbed_energy": float(info[10]), "co2_production": 0.0} if f & FLAG_GRID_SINK
                       else {"provided_energy": float(info[9]), "co2_production": float(info[11])}]
    out["pv"] = [{"provided_energy": float(info[1]), "curtailment": float(info[2])}]
    out["unbalanced_energy"] = [{"absorbed_energy": float(info[4])} if f & FLAG_EXCESS else {"provided_energy": float(info[3])}]
    return out


def control_names(p):
    return [n for n in CONTROL_ORDER if (n == "battery" or (n == "genset" and p.has_genset) or (n == "grid" and p.has_grid))]


def control_dict_to_row(control, p, act_cols):
    """The reference's control dict ({name: [value]} or {name: value}; genset value = [goal, energy]) -> action row.
    Missing modules raise ValueError, extra keys warn (microgrid.py:262-284)."""
    row = np.zeros(p.n_act)
    control = dict(control)
    for name in control_names(p):
        try:
            v = control.pop(name)
        except KeyError:
            raise ValueError(f'Control for module "{name}" not found. Available controls:\n\t{control.keys()}')
        if isinstance(v, (list, tuple)) and len(v) == 1:      # {name: [value]} -> value (one module per name)
            v = v[0]
        arr = np.asarray(v, dtype=np.float64).reshape(-1)
        col = act_cols[name]
        if name == "genset":
            if arr.size != 2:
                raise ValueError(f"Bad action {v}")
            row[col:col + 2] = arr
        else:
            if arr.size != 1:
                raise ValueError(f"Bad action {v}")
            row[col] = arr[0]
    if control:
        warnings.warn(f'\nIgnoring the following keys in passed control:\n {list(control.keys())}')
    return row


# ---- unnormalised state / log (microgrid.py:434-475, base_module.py:276-290) -------------------------------------------
def series_bounds(ts, pull_zero):
    lo, hi = ts.min(axis=0), ts.max(axis=0)
    if pull_zero:   # base_timeseries_module.py:81-88
        lo, hi = (0.0 if lo > 0 else lo), (hi if lo > 0 else (0.0 if hi < 0 else hi))
    return lo, hi


def series_state(ts, t, horizon, low, high):
    """[current, forecast_0 .. forecast_{H-1}] rows of a time-series module, unnormalised, with end padding."""
    ts = ts.reshape(len(ts), -1)
    fill = (np.asarray(high) + np.asarray(low)) / 2
    rows = []
    for k in range(horizon + 1):
        idx = t + k
        rows.append(ts[idx] if (idx < len(ts) and t < len(ts)) else np.broadcast_to(fill, ts.shape[1:]))
    return np.array(rows, dtype=np.float64).reshape(-1)


def state_dict(p, t, charge, genset, soc=None):
    """Unnormalised state of every module at step t, keyed like the reference's `state_dict()`s.  `soc`: the battery's
    stored `_soc` when it is not charge / max_capacity (before its first update, battery_module.py:89, 125-130)."""
    H = p.forecast_horizon
    out = OrderedDict()

    def ts_entries(name, labels, ts, pull):
        ts2 = ts.reshape(len(ts), -1)
        lo, hi = (series_bounds(ts2[:, 0], True) if pull else (ts2.min(axis=0), ts2.max(axis=0)))
        vals = series_state(ts2, t, H, lo, hi)
        d = OrderedDict()
        for c, lab in enumerate(labels):
            d[f"{lab}_current"] = vals[c]
        for j in range(H):
            for c, lab in enumerate(labels):
                d[f"{lab}_forecast_{j}"] = vals[(j + 1) * len(labels) + c]
        out[name] = d
    ts_entries("load", ["load"], p.load_ts, True)
    ts_entries("pv", ["renewable"], p.pv_ts, True)
    out["unbalanced_energy"] = OrderedDict()
    if p.has_genset:
        cs, gs, up, dn = genset
        out["genset"] = OrderedDict(current_status=int(cs), goal_status=int(gs), steps_until_up=int(up), steps_until_down=int(dn))
    out["battery"] = OrderedDict(soc=charge / p.battery.max_capacity if soc is None else soc, current_charge=charge)
    if p.has_grid:
        ts_entries("grid", ["import_price", "export_price", "co2_per_kwh", "grid_status"], p.grid.time_series, False)
    return out


def log_row(p, pre_state, info, reward, genset_after=None):
    """One row of `Microgrid.get_log()`: {(module, 0, field): value} in the reference's column order.
    `pre_state` is `state_dict()` taken before the step -- except for the genset: the reference updates the genset
    status BEFORE it snapshots the state it logs (genset_module.py:148-149 -> base_module.py:152), so the logged
    genset tuple is the one AFTER the step (`genset_after`)."""
    row = OrderedDict()
    if p.has_genset and genset_after is not None:
        cs, gs, up, dn = genset_after
        pre_state = OrderedDict(pre_state)
        pre_state["genset"] = OrderedDict(current_status=int(cs), goal_status=int(gs), steps_until_up=int(up), steps_until_down=int(dn))

    def put(name, fields):
        for k, v in fields.items():
            row[(name, 0, k)] = v
        for k, v in pre_state[name].items():
            row[(name, 0, k)] = v
    put("load", OrderedDict(reward=0.0, load_met=float(info[0])))
    put("pv", OrderedDict(reward=0.0, curtailment=float(info[2]), renewable_used=float(info[1])))
    put("unbalanced_energy", OrderedDict(reward=float(info[15]), loss_load=float(info[3]), overgeneration=float(info[4])))
    if p.has_genset:
        put("genset", OrderedDict(reward=float(info[12]), co2_production=float(info[6]), genset_production=float(info[5])))
    put("battery", OrderedDict(reward=float(info[13]), discharge_amount=float(info[7]), charge_amount=float(info[8])))
    if p.has_grid:
        put("grid", OrderedDict(reward=float(info[14]), co2_production=float(info[11]), grid_import=float(info[9]),
                                grid_export=float(info[10])))
    # balance block (microgrid.py:259-260, 281, 317-319): sums in the reference's list order
    # (np.sum over the info lists is sequential from 0.0; adding the 0.0 of an absent entry changes nothing)
    fixed_absorbed = 0.0 + float(info[0])
    provided = 0.0
    for v in ([float(info[5])] if p.has_genset else []) + [float(info[7])] + ([float(info[9])] if p.has_grid else []):
        provided += v
    consumed = fixed_absorbed + float(info[8])
    if p.has_grid:
        consumed += float(info[10])
    ctrl_provided = provided - 0.0
    ctrl_absorbed = consumed - fixed_absorbed
    overall_provided = provided + float(info[1]) + float(info[3])
    overall_absorbed = consumed + float(info[4])
    # with a reward shaper, `reward` (what run() returned) is the shaped value and the balance log keeps the plain sum of
    # the module rewards beside it, added in dispatch order (utils/step.py:18, microgrid.py:316-319)
    unshaped = reward
    if getattr(p, "reward_shaper", None) is not None:
        unshaped = 0.0
        for col in ([12] if p.has_genset else []) + [13] + ([14] if p.has_grid else []) + [15]:
            unshaped += float(info[col])
    row[("balance", 0, "reward")] = unshaped
    row[("balance", 0, "shaped_reward")] = reward
    row[("balance", 0, "overall_provided_to_microgrid")] = overall_provided
    row[("balance", 0, "overall_absorbed_from_microgrid")] = overall_absorbed
    row[("balance", 0, "controllable_provided_to_microgrid")] = ctrl_provided
    row[("balance", 0, "controllable_absorbed_from_microgrid")] = ctrl_absorbed
    row[("balance", 0, "fixed_provided_to_microgrid")] = 0.0
    row[("balance", 0, "fixed_absorbed_from_microgrid")] = fixed_absorbed
    return row


def log_frame(rows, stop, drop_singleton_key=False):
    """`Microgrid.get_log()` (microgrid.py:434-475) from per-step row dicts {(module, number, field): value}.  Columns keep
    the reference's order: modules in the order of the first row, each module's fields in first-logged order -- a field
    that appears later (a longer forecast after set_forecaster) is appended to ITS module's block and is NaN before
    (ModularLogger.log, utils/logger.py:18-28).  A field that stops being logged is NaN afterwards (the reference's
    get_log() raises a length mismatch in that situation)."""
    import pandas as pd
    blocks = OrderedDict()
    for r in rows:
        for col in r:
            blocks.setdefault(col[:2], OrderedDict()).setdefault(col, None)
    cols = [c for b in blocks.values() for c in b]
    uniform = all(len(r) == len(cols) for r in rows)
    data = [list(r.values()) if uniform else [r.get(c, np.nan) for c in cols] for r in rows]
    df = pd.DataFrame(data, columns=pd.MultiIndex.from_tuples(cols, names=["module_name", "module_number", "field"]) if cols else None,
                      index=pd.RangeIndex(start=stop - len(rows), stop=stop))
    if drop_singleton_key and cols:
        df.columns = df.columns.remove_unused_levels()
    return df


def drop_stale_forecasts(row, stale):
    """The first step after `set_forecaster` logs the forecast the module computed BEFORE the change: the reference's
    `_state_dict` zips the new keys with the stale `_current_forecast` (base_timeseries_module.py:332-338, :99-101), so
    that row carries only the first min(old, new) forecast rows.  `stale`: {(module, number): old horizon}."""
    def keep(col):
        old = stale.get(col[:2])
        if old is None or "_forecast_" not in col[2]:
            return True
        return int(col[2].rsplit("_", 1)[1]) < old
    return OrderedDict((c, v) for c, v in row.items() if keep(c))


def caller_names(d, p):
    """engine keys ('pv', 'unbalanced_energy') -> the caller's module names (`p.renewable_name`, `p.unbalanced_name`), for
    dicts keyed by module name or by (module name, number, field)"""
    if p.renewable_name == "pv" and p.unbalanced_name == "unbalanced_energy":
        return d
    nm = lambda k: {"pv": p.renewable_name, "unbalanced_energy": p.unbalanced_name}.get(k, k)      # noqa: E731
    return type(d)(((nm(k) if not isinstance(k, tuple) else (nm(k[0]),) + k[1:]), v) for k, v in d.items())


def flushed_balance_log(rows):
    """what Microgrid.reset() returns under 'balance': the balance log it has just flushed, {field: [value per step]}
    (microgrid.py:205-219: `self._balance_logger.flush()`)"""
    out = OrderedDict()
    for r in rows:
        for (name, _, field), v in r.items():
            if name == "balance":
                out.setdefault(field, []).append(v)
    return dict(out)
