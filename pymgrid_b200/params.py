"""Parameter record of ONE microgrid, in the reference's own vocabulary.

A `MicrogridParams` holds exactly what the reference's module constructors take for the modules on the
hot path plus the mutable state the reference serialises
(`base_module.py:852-868`, `genset_module.py:426-427`): battery charge, genset status tuple, step.
It is the neutral hand-over format between the scenario readers (`scenario.py`), the batched engine
(`engine.py`) and -- in the tests only -- the CPU oracle binding (`oracle/oracle.py`).
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

DEFAULT_HORIZON = 23  # reference: src/pymgrid/microgrid/__init__.py:1


@dataclass
class BatteryParams:
    """reference: modules/battery_module.py:66-91"""
    min_capacity: float
    max_capacity: float
    max_charge: float
    max_discharge: float
    efficiency: float
    battery_cost_cycle: float = 0.0
    current_charge: float = 0.0     # state: _current_charge
    # state: _soc.  The reference keeps the soc it was CONSTRUCTED with (init_soc, battery_module.py:96-106) until the
    # first update recomputes it as current_charge / max_capacity (:125-130); init_soc * max_capacity / max_capacity can
    # differ from init_soc in the last bit, so a record built from init_soc carries it.  None = current_charge / max_capacity.
    soc: Optional[float] = None

    @property
    def reported_soc(self):
        return self.current_charge / self.max_capacity if self.soc is None else self.soc

    @property
    def min_soc(self):
        return self.min_capacity / self.max_capacity

    @property
    def min_act(self):              # battery_module.py:332-334
        return -self.max_discharge / self.efficiency

    @property
    def max_act(self):              # battery_module.py:336-338
        return self.max_charge * self.efficiency


@dataclass
class GensetParams:
    """reference: modules/genset_module.py:61-92"""
    running_min_production: float
    running_max_production: float
    genset_cost: float
    co2_per_unit: float = 0.0
    cost_per_unit_co2: float = 0.0
    start_up_time: int = 0
    wind_down_time: int = 0
    allow_abortion: bool = True
    # state (genset_module.py:91-92, :429-433)
    current_status: int = 1
    goal_status: int = 1
    steps_until_up: int = 0
    steps_until_down: int = 0

    @classmethod
    def with_init(cls, init_start_up=True, **kw):
        g = cls(**kw)
        g.current_status = g.goal_status = int(init_start_up)
        if g.current_status:
            g.steps_until_up, g.steps_until_down = 0, g.wind_down_time
        else:
            g.steps_until_up, g.steps_until_down = g.start_up_time, 0
        return g


@dataclass
class GridParams:
    """reference: modules/grid_module.py:70-123; time_series columns import_price, export_price,
    co2_per_kwh, grid_status (a 3-column input gets status == 1, grid_module.py:112-118)."""
    max_import: float
    max_export: float
    time_series: np.ndarray         # [T, 4] float64
    cost_per_unit_co2: float = 0.0
    # optional per-grid status (0/1, [T]) overriding column 3 of a SHARED time_series table: weak grids of
    # MicrogridGenerator (MicrogridGenerator.py:321-340) differ only in this column
    status: Optional[np.ndarray] = None

    def effective_time_series(self):
        if self.status is None:
            return self.time_series
        ts = np.array(self.time_series, dtype=np.float64)
        ts[:, 3] = self.status
        return ts


@dataclass
class ForecasterParams:
    """Gaussian-noise forecaster of one time-series module: what the reference builds when a module's `forecaster`
    argument is a number (forecast/forecaster.py:10-89 get_forecaster, :220-250 GaussianNoiseForecaster)."""
    noise_std: float = 0.0                 # 0 = oracle forecaster (perfect forecast)
    increase_uncertainty: bool = False     # std_k = std * (1 + log(1 + k)) for forecast row k
    relative_noise: bool = False           # std *= |mean(time_series[initial_step:final_step])|


REWARD_SHAPERS = {None: 0, "pv_curtailment": 1, "battery_discharge": 2}     # MG_SHAPER_* / ORC_SHAPER_*
REWARD_SHAPER_ALIASES = {"PVCurtailmentShaper": "pv_curtailment", "!PVCurtailmentShaper": "pv_curtailment",
                         "BatteryDischargeShaper": "battery_discharge", "!BatteryDischargeShaper": "battery_discharge"}


@dataclass
class MicrogridParams:
    battery: BatteryParams
    load_ts: np.ndarray             # [T] float64, stored NEGATIVE like the reference (base_timeseries_module.py:68-79)
    pv_ts: np.ndarray               # [T] float64, >= 0
    genset: Optional[GensetParams] = None
    grid: Optional[GridParams] = None
    loss_load_cost: float = 10.0    # microgrid.py:103-104 defaults 10 / 2; pymgrid25 uses 10 / 1
    overgeneration_cost: float = 2.0
    forecast_horizon: int = DEFAULT_HORIZON   # 0 = no forecaster
    initial_step: int = 0
    final_step: int = -1            # <= 0 means len(series) (base_timeseries_module.py:317-330)
    current_step: int = 0
    name: str = ""
    meta: dict = field(default_factory=dict)
    # profile-times-scale series (MicrogridGenerator._scale_ts, MicrogridGenerator.py:137-148): the module's series is
    # load_ts * load_scale / pv_ts * pv_scale, with load_ts / pv_ts a shared profile.  1.0 for table-backed grids.
    load_scale: float = 1.0
    pv_scale: float = 1.0
    renewable_name: str = "pv"      # 'PV' for MicrogridGenerator grids: decides the gym-sorted observation order
    # name of the slack module in dicts, logs and `modules`: 'unbalanced_energy' in the pymgrid25 YAMLs; 'balancing' (the
    # class's module_type[0], module_container.py:366-374) when Microgrid(modules) appends it itself (microgrid.py:170-171)
    unbalanced_name: str = "unbalanced_energy"
    # Microgrid(reward_shaping_func=...): None, "pv_curtailment" (PVCurtailmentShaper) or "battery_discharge"
    # (BatteryDischargeShaper), microgrid/reward_shaping/*.py; the shaped value replaces the step reward (utils/step.py:38-46)
    reward_shaper: Optional[str] = None
    # per-module forecasters, keys among 'load', 'pv', 'grid'; a missing key is the oracle forecaster (pymgrid25's setting).
    # Noise only changes the forecast entries of the observation, never the physics (forecast/forecaster.py).
    forecasters: dict = field(default_factory=dict)

    def __post_init__(self):
        if not (self.load_scale > 0 and self.pv_scale > 0):
            raise ValueError("series scales must be positive")
        bad = set(self.forecasters) - {"load", "pv", "grid"}
        if bad:
            raise ValueError(f"forecasters: unknown module(s) {sorted(bad)}; time-series modules are load, pv, grid")
        self.forecasters = {k: (v if isinstance(v, ForecasterParams) else ForecasterParams(float(v)))
                            for k, v in self.forecasters.items()}
        self.reward_shaper = REWARD_SHAPER_ALIASES.get(self.reward_shaper, self.reward_shaper)
        if self.reward_shaper not in REWARD_SHAPERS:
            raise ValueError(f"unknown reward shaper {self.reward_shaper!r}; built in: {sorted(k for k in REWARD_SHAPERS if k)}")
        if self.reward_shaper == "pv_curtailment" and self.renewable_name != "pv":
            # sum_module_val(info, 'pv', ...) finds no such module and returns 0.0 (reward_shaping/base.py:11-16)
            raise ValueError("PVCurtailmentShaper reads the module named 'pv'; this grid names it " + repr(self.renewable_name))
        if self.reward_shaper == "battery_discharge" and self.unbalanced_name != "unbalanced_energy":
            # the shaper sums the loss load of the module NAMED 'unbalanced_energy' (battery_discharge_shaper.py:25); under
            # another name (e.g. the 'balancing' module Microgrid(modules) appends) the reference silently counts none
            raise ValueError("BatteryDischargeShaper reads the loss load of the module named 'unbalanced_energy'; this grid names "
                             "its slack module " + repr(self.unbalanced_name) + " (pass ('unbalanced_energy', UnbalancedEnergyModule(...)) "
                             "with add_unbalanced_module=False)")
        self.load_ts = -np.abs(np.ascontiguousarray(self.load_ts, dtype=np.float64).reshape(-1))
        self.pv_ts = np.abs(np.ascontiguousarray(self.pv_ts, dtype=np.float64).reshape(-1))
        if self.grid is not None:
            ts = np.asarray(self.grid.time_series, dtype=np.float64)
            if ts.ndim != 2 or ts.shape[1] not in (3, 4):
                raise ValueError('Time series must be two dimensional with three or four columns.')
            if ts.shape[1] == 3:
                ts = np.concatenate([ts, np.ones((ts.shape[0], 1))], axis=1)
            elif not ((ts[:, -1] == 0) | (ts[:, -1] == 1)).all():
                raise ValueError("Last column (grid status) must contain binary values.")
            if (ts < 0).any():
                raise ValueError('Time series must be non-negative.')
            self.grid.time_series = np.ascontiguousarray(ts)
            if len(ts) != len(self.load_ts):
                raise ValueError('all time series must have the same length')
        if len(self.pv_ts) != len(self.load_ts):
            raise ValueError('all time series must have the same length')
        if self.final_step <= 0:
            self.final_step = len(self.load_ts)
        if self.final_step <= self.initial_step:
            raise ValueError('final_step value must be greater than initial_step')

    @property
    def scaled(self):
        return self.load_scale != 1.0 or self.pv_scale != 1.0

    @property
    def effective_load_ts(self):
        return self.load_ts * self.load_scale if self.load_scale != 1.0 else self.load_ts

    @property
    def effective_pv_ts(self):
        return self.pv_ts * self.pv_scale if self.pv_scale != 1.0 else self.pv_ts

    # -- shape of the flat spaces ---------------------------------------------
    def __len__(self):
        return len(self.load_ts)

    @property
    def has_genset(self):
        return self.genset is not None

    @property
    def has_grid(self):
        return self.grid is not None

    @property
    def arch(self):
        return (int(self.has_genset), int(self.has_grid), int(self.forecast_horizon))

    @property
    def obs_dim(self):
        rows = 1 + self.forecast_horizon
        return rows * (2 + 4 * self.has_grid) + 2 + 4 * self.has_genset

    @property
    def n_act(self):
        return 1 + int(self.has_grid) + 2 * int(self.has_genset)
