"""Turn a LIVE reference `Microgrid` object into the neutral parameter record -- build container only.

Test infrastructure: lets the CPU tests construct the oracle / engine inputs from the reference's own
objects (independently of our scenario reader) so that loader bugs cannot hide behind matching outputs.
"""
from types import SimpleNamespace

import numpy as np


def params_from_reference(m):
    """m: pymgrid.Microgrid (live reference object).  Returns an attribute bag shaped like MicrogridParams."""
    mods = m.modules
    b = mods.battery[0]
    battery = SimpleNamespace(min_capacity=b.min_capacity, max_capacity=b.max_capacity, max_charge=b.max_charge,
                              max_discharge=b.max_discharge, efficiency=b.efficiency,
                              battery_cost_cycle=b.battery_cost_cycle, current_charge=b.current_charge)
    genset = grid = None
    if hasattr(mods, "genset"):
        g = mods.genset[0]
        genset = SimpleNamespace(running_min_production=g.running_min_production,
                                 running_max_production=g.running_max_production, genset_cost=g.genset_cost,
                                 co2_per_unit=g.co2_per_unit, cost_per_unit_co2=g.cost_per_unit_co2,
                                 start_up_time=g.start_up_time, wind_down_time=g.wind_down_time,
                                 allow_abortion=g.allow_abortion, current_status=int(g._current_status),
                                 goal_status=int(g._goal_status), steps_until_up=int(g._steps_until_up),
                                 steps_until_down=int(g._steps_until_down))
    if hasattr(mods, "grid"):
        g = mods.grid[0]
        grid = SimpleNamespace(max_import=g.max_import, max_export=g.max_export,
                               time_series=np.array(g.time_series, dtype=np.float64),
                               cost_per_unit_co2=g.cost_per_unit_co2)
    load, pv = mods.load[0], mods.pv[0] if hasattr(mods, "pv") else mods.renewable[0]
    unb = mods.unbalanced_energy[0] if hasattr(mods, "unbalanced_energy") else mods.balancing[0]
    shaper = getattr(m, "reward_shaping_func", None)
    shaper = {None: None, "PVCurtailmentShaper": "pv_curtailment",
              "BatteryDischargeShaper": "battery_discharge"}[None if shaper is None else type(shaper).__name__]
    return SimpleNamespace(battery=battery, genset=genset, grid=grid, reward_shaper=shaper,
                           load_ts=np.array(load.time_series[:, 0], dtype=np.float64),
                           pv_ts=np.array(pv.time_series[:, 0], dtype=np.float64),
                           loss_load_cost=unb.loss_load_cost, overgeneration_cost=unb.overgeneration_cost,
                           forecast_horizon=load.forecast_horizon, initial_step=m.initial_step,
                           final_step=m.final_step, current_step=m.current_step)
