"""Heterogeneous microgrids drawn from MicrogridGenerator's parameter distributions (BASELINE config 5).

The reference builds one NonModularMicrogrid per grid in Python (37 ms each) and converts it to modules
(MicrogridGenerator.py:443-603, convert/get_module.py).  For a million grids this module samples the same
distributions vectorised with a seeded numpy Generator and hands the result to the engine in array form:

  architecture   u < .33 genset | u < .66 grid | else both (:417-435); a weak grid forces a genset (:535-538)
  load           one of 5 profiles scaled to a peak of randint(100, 100001) (:437-441, :458): series = profile * (peak / max)
  PV             one of 5 profiles scaled to peak load * randint(30, 151) / 100 (:357, :493)
  battery        capacity ceil(randint(3, 6) * mean load), power ceil(capacity / 4), eta .9, soc_min .2,
                 soc_0 = clip(randn(), .2, 1), cycle cost .02 (:230-243, :382-386)
  genset         rated ceil(peak / .9), running min / max = .05 / .9 * rated, fuel .4, co2 2 (:214-228, :372-379)
  grid           import = export = int(2 * peak), tariff 1 | 2 (:253-285), one of 2 CO2 profiles, weak-grid outage
                 profile (:321-340) when weak, cost_co2 .1, loss load 10, overgeneration 1, horizon 23, 'PV' module name

A grid's series are (profile id, scale): the engine multiplies on the fly, bit-identical to the reference's
`profile * ratio` (tests/test_generator.py checks real MicrogridGenerator grids).  `to_params(i)` materialises one grid
as a `MicrogridParams` (explicit form) for the oracle / the B=1 surface.
"""
import os
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .params import BatteryParams, GensetParams, GridParams, MicrogridParams
from .priority_list import priority_lists

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "generator_profiles.npz")
T = 8760
HORIZON = 23


def load_profiles():
    with np.load(_DATA) as z:
        return {k: z[k] for k in ("load", "pv", "co2")}


def tariff_import(scenario):
    """MicrogridGenerator._get_electricity_tariff (MicrogridGenerator.py:253-285)."""
    h = np.arange(T) % 24
    if scenario == 1:   # PG&E A-6 TOU
        return np.where((h >= 12) & (h < 18), 0.59, np.where((h < 8) | (h >= 21), 0.22, 0.29))
    return np.where(((h >= 0) & (h < 5)) | ((h >= 14) & (h < 17)), 0.08, 0.11)


def weak_grid_profile(rng, outage_per_day, duration):
    """MicrogridGenerator._generate_weak_grid_profile (MicrogridGenerator.py:321-340), timestep 1."""
    u = rng.random(T + 1)
    ts = (u >= outage_per_day / 24).astype(np.uint8)
    zeros = np.nonzero(ts == 0)[0]
    for j in range(1, int(duration)):
        idx = zeros - j
        ts[idx[idx > 0]] = 0
    return ts[:T]


@dataclass
class GeneratorBatch:
    n: int
    has_genset: np.ndarray
    has_grid: np.ndarray
    load_profile: np.ndarray
    load_scale: np.ndarray
    pv_profile: np.ndarray
    pv_scale: np.ndarray
    bat_capacity: np.ndarray
    bat_power: np.ndarray
    bat_soc0: np.ndarray
    gen_rated: np.ndarray
    grid_power: np.ndarray
    grid_weak: np.ndarray
    tariff: np.ndarray
    co2_profile: np.ndarray
    status: np.ndarray            # [n, T] uint8 (ones where the grid is not weak / absent)
    profiles: dict

    def to_params(self, i):
        """Explicit single-grid form of grid i (what convert/get_module.py would build)."""
        pr = self.profiles
        cap = float(self.bat_capacity[i])
        battery = BatteryParams(min_capacity=cap * 0.2, max_capacity=cap, max_charge=float(self.bat_power[i]),
                                max_discharge=float(self.bat_power[i]), efficiency=0.9, battery_cost_cycle=0.02,
                                current_charge=float(self.bat_soc0[i]) * cap, soc=float(self.bat_soc0[i]))
        genset = grid = None
        if self.has_genset[i]:
            r = float(self.gen_rated[i])
            genset = GensetParams.with_init(running_min_production=0.05 * r, running_max_production=0.9 * r, genset_cost=0.4,
                                            co2_per_unit=2, cost_per_unit_co2=0.1)
        if self.has_grid[i]:
            ts = grid_table(pr, int(self.tariff[i]), int(self.co2_profile[i]))
            grid = GridParams(max_import=float(self.grid_power[i]), max_export=float(self.grid_power[i]), time_series=ts,
                              cost_per_unit_co2=0.1, status=self.status[i].astype(np.float64))
        return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=pr["load"][self.load_profile[i]],
                               pv_ts=pr["pv"][self.pv_profile[i]], load_scale=float(self.load_scale[i]),
                               pv_scale=float(self.pv_scale[i]), loss_load_cost=10.0, overgeneration_cost=1.0,
                               forecast_horizon=HORIZON, final_step=-1, renewable_name="PV")


_GRID_TABLES = {}


def grid_table(profiles, tariff, co2_id):
    key = (tariff, co2_id)
    if key not in _GRID_TABLES:
        _GRID_TABLES[key] = np.ascontiguousarray(np.stack([tariff_import(tariff), np.zeros(T), profiles["co2"][co2_id], np.ones(T)], axis=1))
    return _GRID_TABLES[key]


def sample(n, seed=0):
    """n grids from MicrogridGenerator's distributions (vectorised; not the reference's RNG stream)."""
    rng = np.random.default_rng(seed)
    pr = load_profiles()
    u = rng.random(n)
    has_genset = (u < 0.33) | (u >= 0.66)
    has_grid = u >= 0.33
    weak = has_grid & (rng.integers(0, 2, n) == 1)
    has_genset = has_genset | weak
    tariff = rng.integers(1, 3, n)
    load_profile = rng.integers(0, pr["load"].shape[0], n)
    pv_profile = rng.integers(0, pr["pv"].shape[0], n)
    co2_profile = rng.integers(0, pr["co2"].shape[0], n)
    size_load = rng.integers(100, 100001, n)
    load_max, load_mean = pr["load"].max(axis=1), pr["load"].mean(axis=1)
    load_scale = size_load / load_max[load_profile]                       # size / df_ts.max()
    peak = load_max[load_profile] * load_scale                            # load.max() of the scaled series
    pv_size = peak * (rng.integers(30, 151, n) / 100)
    pv_scale = pv_size / pr["pv"].max(axis=1)[pv_profile]
    bat_capacity = np.ceil(rng.integers(3, 6, n) * (load_mean[load_profile] * load_scale)).astype(np.int64)
    bat_power = np.ceil(bat_capacity / 4).astype(np.int64)
    bat_soc0 = np.clip(rng.standard_normal(n), 0.2, 1.0)
    gen_rated = np.ceil(peak / 0.9).astype(np.int64)
    grid_power = (peak * 2).astype(np.int64)
    status = np.ones((n, T), dtype=np.uint8)
    for i in np.nonzero(weak)[0]:
        status[i] = weak_grid_profile(rng, rng.standard_normal() * 3 / 4 + 0.25, rng.integers(1, 8))
    return GeneratorBatch(n=n, has_genset=has_genset, has_grid=has_grid, load_profile=load_profile, load_scale=load_scale,
                          pv_profile=pv_profile, pv_scale=pv_scale, bat_capacity=bat_capacity, bat_power=bat_power,
                          bat_soc0=bat_soc0, gen_rated=gen_rated, grid_power=grid_power, grid_weak=weak, tariff=tariff,
                          co2_profile=co2_profile, status=status, profiles=pr)


def engine_from_batch(gb, device=None, obs_order="gym_sorted", with_info=False, with_flags=True, action_order=None,
                      env_slice=None, obs_dtype=None):
    """Array-form construction of a BatchedMicrogrid from a GeneratorBatch (one config record per env, built with numpy).
    env_slice: (lo, hi) to build only this rank's contiguous shard of the batch."""
    from .engine import BatchedMicrogrid
    lo, hi = (0, gb.n) if env_slice is None else env_slice
    n = hi - lo
    pr = gb.profiles
    cfg, tables, plist_rows, load_tab, pv_tab, grid_np = batch_config_records(gb, lo, hi)
    sel = slice(lo, hi)
    hg, hr = gb.has_genset[sel], gb.has_grid[sel]
    cap = gb.bat_capacity[sel].astype(np.float64)
    status = gb.status[sel]
    bm = BatchedMicrogrid.__new__(BatchedMicrogrid)
    bm.configs = None
    bm.generator_batch = gb
    bm._setup(cfg_np=cfg, plist_np=np.concatenate(plist_rows).view(np.uint8), action_tables=tables,
              env_config=np.arange(n), cfg_arch=np.stack([hg.astype(np.int64), hr.astype(np.int64), np.full(n, HORIZON)], axis=1),
              cfg_step=np.zeros(n, dtype=np.int32), cfg_charge=gb.bat_soc0[sel] * cap, cfg_soc=gb.bat_soc0[sel],
              cfg_genset=np.where(hg, 0x0101, 0).astype(np.int32), load_np=load_tab, pv_np=pv_tab, grid_np=grid_np,
              cfg_status=status, device=device, obs_order="gym_sorted_pv_first" if obs_order == "gym_sorted" else obs_order,
              with_info=with_info, with_flags=with_flags, action_order=action_order,
              **({} if obs_dtype is None else {"obs_dtype": obs_dtype}))
    bm.global_env_ids = np.arange(lo, hi)
    return bm


def batch_config_records(gb, lo=0, hi=None):
    """The MgConfig records of grids [lo, hi) of a GeneratorBatch, built with numpy (no device needed), plus the shared
    tables they index: (cfg, action tables per architecture, priority-list rows, load table, pv table, grid tables).
    Field for field the same as engine.config_record(gb.to_params(i), ...) (tests/test_generator.py)."""
    hi = gb.n if hi is None else hi
    sel = slice(lo, hi)
    n = hi - lo
    pr = gb.profiles
    cfg = np.zeros(n, dtype=np.dtype(_cabi.MgConfig))
    cap = gb.bat_capacity[sel].astype(np.float64)
    power = gb.bat_power[sel].astype(np.float64)
    eff = 0.9
    cfg["bat_min_capacity"], cfg["bat_max_capacity"] = cap * 0.2, cap
    cfg["bat_max_charge"] = cfg["bat_max_discharge"] = power
    cfg["bat_efficiency"], cfg["bat_cost_cycle"] = eff, 0.02
    act_lo, act_hi = -power / eff, power * eff
    cfg["bat_act_low"] = act_lo
    sp = act_hi - act_lo
    cfg["bat_act_spread"] = np.where(sp == 0, 1.0, sp)
    min_soc = cfg["bat_min_capacity"] / cap
    cfg["bat_soc_low"] = min_soc
    sp = 1.0 - min_soc
    cfg["bat_soc_spread"] = np.where(sp == 0, 1.0, sp)
    sp = cap - cfg["bat_min_capacity"]
    cfg["bat_charge_spread"] = np.where(sp == 0, 1.0, sp)
    hg, hr = gb.has_genset[sel], gb.has_grid[sel]
    rated = gb.gen_rated[sel].astype(np.float64)
    cfg["gen_running_min"] = np.where(hg, 0.05 * rated, 0.0)
    cfg["gen_running_max"] = np.where(hg, 0.9 * rated, 0.0)
    cfg["gen_cost"], cfg["gen_co2_per_unit"], cfg["gen_cost_per_unit_co2"] = np.where(hg, 0.4, 0.0), np.where(hg, 2.0, 0.0), np.where(hg, 0.1, 0.0)
    sp = cfg["gen_running_max"] - 0.0
    cfg["gen_act_spread"] = np.where(sp == 0, 1.0, sp)
    cfg["gen_up_spread"] = cfg["gen_down_spread"] = 1.0
    cfg["gen_allow_abortion"] = hg.astype(np.int32)
    gp = gb.grid_power[sel].astype(np.float64)
    cfg["grid_max_import"] = cfg["grid_max_export"] = np.where(hr, gp, 0.0)
    cfg["grid_cost_per_unit_co2"] = np.where(hr, 0.1, 0.0)
    cfg["grid_act_low"] = -1 * cfg["grid_max_export"]
    sp = cfg["grid_max_import"] - cfg["grid_act_low"]
    cfg["grid_act_spread"] = np.where(sp == 0, 1.0, sp)
    cfg["loss_load_cost"], cfg["overgeneration_cost"] = 10.0, 1.0
    cfg["load_scale"], cfg["pv_scale"] = gb.load_scale[sel], gb.pv_scale[sel]
    cfg["series_scaled"] = 1
    load_tab = -np.abs(pr["load"])          # stored negative like the reference (base_timeseries_module.py:68-79)
    pv_tab = np.abs(pr["pv"])
    def bounds(tab_minmax, prof, scale):
        low0, high0 = tab_minmax[prof, 0] * scale, tab_minmax[prof, 1] * scale
        low = np.where(low0 > 0, 0.0, low0)                       # base_timeseries_module.py:81-88
        high = np.where((low0 <= 0) & (high0 < 0), 0.0, high0)
        spread = np.where(high - low == 0, 1.0, high - low)
        fill = np.minimum(np.maximum((high + low) / 2, low), high)
        return low, spread, (fill - low) / spread
    raw_l = np.stack([load_tab.min(axis=1), load_tab.max(axis=1)], axis=1)
    raw_p = np.stack([pv_tab.min(axis=1), pv_tab.max(axis=1)], axis=1)
    cfg["load_low"], cfg["load_spread"], cfg["load_fill_nrm"] = bounds(raw_l, gb.load_profile[sel], gb.load_scale[sel])
    cfg["pv_low"], cfg["pv_spread"], cfg["pv_fill_nrm"] = bounds(raw_p, gb.pv_profile[sel], gb.pv_scale[sel])
    cfg["load_series"], cfg["pv_series"] = gb.load_profile[sel], gb.pv_profile[sel]
    n_co2 = pr["co2"].shape[0]
    cfg["grid_series"] = np.where(hr, (gb.tariff[sel] - 1) * n_co2 + gb.co2_profile[sel], 0)
    cfg["initial_step"], cfg["final_step"] = 0, T
    status = gb.status[sel]
    cfg["grid_status_weak"] = (status.min(axis=1) < 1).astype(np.int32)
    grid_np = np.stack([grid_table(pr, t, c) for t in (1, 2) for c in range(n_co2)])
    # priority-list tables per architecture (running_min > 0 for every generated genset: nothing is redundant)
    tables, plist_rows, offsets = {}, [], {}
    for a in ((1, 0), (0, 1), (1, 1)):
        pls = priority_lists(bool(a[0]), bool(a[1]), 1.0)
        tables[a], offsets[a] = pls, len(plist_rows)
        for pl in pls:
            row = np.zeros(8, dtype=np.int8)
            for j in range(_cabi.MG_PLIST_WIDTH):
                row[j], row[3 + j] = (pl[j] if j < len(pl) else (_cabi.MG_MOD_NONE, 0))
            row[6] = len(pl)
            plist_rows.append(row)
    arch_key = hg.astype(np.int64) * 2 + hr.astype(np.int64)
    cfg["plist_offset"] = np.select([arch_key == 2, arch_key == 1, arch_key == 3], [offsets[(1, 0)], offsets[(0, 1)], offsets[(1, 1)]])
    cfg["plist_count"] = np.select([arch_key == 2, arch_key == 1, arch_key == 3], [len(tables[(1, 0)]), len(tables[(0, 1)]), len(tables[(1, 1)])])
    return cfg, tables, plist_rows, load_tab, pv_tab, grid_np
