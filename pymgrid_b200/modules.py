"""Module constructors with the reference's signatures (src/pymgrid/modules/*.py), so that code written as

    from pymgrid.modules import BatteryModule, GensetModule, GridModule, LoadModule, RenewableModule
    microgrid = Microgrid([battery, ("pv", pv), load, grid], loss_load_cost=..., ...)

runs against this package by changing the import.  In the reference a module is the unit that computes
(`BaseMicrogridModule.step`, base_module.py:95-159); here the physics of all modules of a microgrid is ONE fused CUDA
kernel, so these classes are parameter records: they take the reference's constructor arguments, apply the
constructor's own checks and defaults (cited per class) and are folded by `params_from_modules` into the
`MicrogridParams` record the engine is built from.  Live per-module attributes (`soc`, `max_production`, ...) are served
by the views of a built `Microgrid` (`microgrid.modules.battery[0]`, microgrid.py `ModuleView`).

What the fused kernel covers is the module set of every pymgrid25 / MicrogridGenerator grid -- exactly one load, one
renewable, one battery, at most one genset and one grid, one forecast horizon for all time-series modules.  Anything
else raises NotImplementedError at construction (nothing is approximated or silently dropped).
"""
from warnings import warn

import numpy as np

from .params import (DEFAULT_HORIZON, BatteryParams, ForecasterParams, GensetParams, GridParams, MicrogridParams)


class _Module:
    module_type = (None, None)
    _default_name = None

    def __init__(self, raise_errors=False, initial_step=0):
        self.raise_errors = raise_errors
        self.initial_step = initial_step
        self.name = (self._default_name, None)       # reference: self.name = ('battery', None) etc.

    def __repr__(self):
        keys = [k for k in vars(self) if not k.startswith("_") and k not in ("name", "time_series")]
        return f"{type(self).__name__}(" + ", ".join(f"{k}={getattr(self, k)!r}" for k in keys) + ")"

    # ---- the module on its own: the reference's operator API (modules/base/base_module.py:65-159) -----------------------
    # Inside a Microgrid a module is a parameter record (the kernels step all modules of a microgrid together).  Used
    # without one -- module.step(action), module.reset(), module.state, as the reference's module-level tests do -- it gets
    # a one-module device batch of its own (compose.StandaloneModule) on first use and is stepped by mgc_modules_step.
    _LIVE = ("state", "current_step", "current_charge", "soc", "current_status", "goal_status", "max_production",
             "min_production", "max_consumption", "current_load", "current_renewable", "min_obs", "max_obs", "min_act", "max_act",
             "production_marginal_cost", "absorption_marginal_cost", "marginal_cost", "import_price", "export_price",
             "co2_per_kwh", "grid_status", "action_space", "is_source", "is_sink", "next_status", "next_max_production",
             "next_min_production")

    def _standalone(self):
        runner = self.__dict__.get("_runner")
        if runner is None:
            from .compose import StandaloneModule
            runner = self.__dict__["_runner"] = StandaloneModule(self)
        return runner

    def step(self, action, normalized=True):
        """reference: BaseMicrogridModule.step (base_module.py:95-159) -> (normalised state after the step, reward, done, info)"""
        return self._standalone().step(action, normalized=normalized)

    def reset(self):
        return self._standalone().reset()

    def state_dict(self, normalized=False):
        return self._standalone().view.state_dict(normalized=normalized)

    def log_dict(self):
        """reference: BaseMicrogridModule.log_dict: {field: [value per step]} since the last reset"""
        rows = self._standalone().log_rows
        return {k: [r[k] for r in rows] for k in (rows[0] if rows else {})}

    def __getattr__(self, item):
        if item in type(self)._LIVE:
            return getattr(self._standalone().view, item)
        raise AttributeError(item)


class BatteryModule(_Module):
    """reference: modules/battery_module.py:66-106 (constructor, `_init_battery`)."""
    module_type = ("battery", "controllable")
    _default_name = "battery"

    def __init__(self, min_capacity, max_capacity, max_charge, max_discharge, efficiency, battery_cost_cycle=0.0,
                 battery_transition_model=None, init_charge=None, init_soc=None, initial_step=0, raise_errors=False):
        assert 0 < efficiency <= 1                                           # battery_module.py:78
        if battery_transition_model is not None:
            raise NotImplementedError("battery_transition_model: a Python callable cannot run inside the fused kernel; "
                                      "only the default transition model (battery_module.py:244-278) is built in")
        self.min_capacity, self.max_capacity = min_capacity, max_capacity
        self.max_charge, self.max_discharge = max_charge, max_discharge
        self.efficiency, self.battery_cost_cycle = efficiency, battery_cost_cycle
        self.min_soc, self.max_soc = min_capacity / max_capacity, 1
        if init_charge is not None:                                          # battery_module.py:96-106
            if init_soc is not None:
                warn("Passed both init_capacity and init_soc. Using init_charge and ignoring init_soc")
            init_soc = init_charge / max_capacity
        elif init_soc is not None:
            init_charge = init_soc * max_capacity
        else:
            raise ValueError("Must set one of init_charge and init_soc.")
        self.init_charge, self.init_soc = init_charge, init_soc
        super().__init__(raise_errors, initial_step)

    def _params(self):
        return BatteryParams(min_capacity=self.min_capacity, max_capacity=self.max_capacity, max_charge=self.max_charge,
                             max_discharge=self.max_discharge, efficiency=self.efficiency,
                             battery_cost_cycle=self.battery_cost_cycle, current_charge=self.init_charge,
                             soc=self.init_soc)


class GensetModule(_Module):
    """reference: modules/genset_module.py:61-98."""
    module_type = ("genset", "controllable")
    _default_name = "genset"

    def __init__(self, running_min_production, running_max_production, genset_cost, co2_per_unit=0.0, cost_per_unit_co2=0.0,
                 start_up_time=0, wind_down_time=0, allow_abortion=True, init_start_up=True, initial_step=0,
                 raise_errors=False, provided_energy_name="genset_production"):
        if running_min_production > running_max_production:
            raise ValueError("parameter min_production must not be greater than parameter max_production.")
        if not allow_abortion:
            warn("Gensets that do not allow abortions are not fully tested, setting allow_abortion=False "
                 "may lead to unexpected behavior.")
        if callable(genset_cost):
            raise NotImplementedError("genset_cost: a Python callable cannot run inside the fused kernel; pass the "
                                      "per-unit cost as a number (genset_module.py:183-198)")
        if provided_energy_name != "genset_production":
            raise NotImplementedError("provided_energy_name: the log column is fixed to 'genset_production'")
        self.running_min_production, self.running_max_production = running_min_production, running_max_production
        self.genset_cost, self.co2_per_unit, self.cost_per_unit_co2 = genset_cost, co2_per_unit, cost_per_unit_co2
        self.start_up_time, self.wind_down_time = start_up_time, wind_down_time
        self.allow_abortion, self.init_start_up = allow_abortion, init_start_up
        super().__init__(raise_errors, initial_step)

    def _params(self):
        return GensetParams.with_init(init_start_up=self.init_start_up, running_min_production=self.running_min_production,
                                      running_max_production=self.running_max_production, genset_cost=self.genset_cost,
                                      co2_per_unit=self.co2_per_unit, cost_per_unit_co2=self.cost_per_unit_co2,
                                      start_up_time=int(self.start_up_time), wind_down_time=int(self.wind_down_time),
                                      allow_abortion=bool(self.allow_abortion))


class _TimeSeriesModule(_Module):
    """reference: modules/base/timeseries/base_timeseries_module.py:22-88 (constructor, `_set_time_series`, `_sign_check`)."""
    _is_source, _is_sink = False, False

    def __init__(self, time_series, raise_errors, forecaster, forecast_horizon, forecaster_increase_uncertainty,
                 forecaster_relative_noise, initial_step, final_step):
        ts = np.array(time_series, dtype=np.float64)
        ts = ts.reshape((-1, ts.shape[1]) if ts.ndim > 1 else (-1, 1))
        assert len(ts) == len(time_series)
        if not (self._is_source and self._is_sink):
            if not ((np.sign(ts) <= 0).all() or (np.sign(ts) >= 0).all()):
                raise ValueError("time_series cannot contain both positive and negative values unless it is both "
                                 "a source and a sink.")
            ts = np.abs(ts) if self._is_source else -np.abs(ts)
        self.time_series = ts
        self.forecaster = forecaster
        self.forecast_horizon = forecast_horizon * (forecaster is not None)   # base_timeseries_module.py:42
        self.forecaster_increase_uncertainty = forecaster_increase_uncertainty
        self.forecaster_relative_noise = forecaster_relative_noise
        self.final_step = final_step
        super().__init__(raise_errors, initial_step)

    def __len__(self):
        return len(self.time_series)

    def _forecaster_params(self):
        """forecast/forecaster.py:10-89 (get_forecaster): None -> no forecast, 'oracle' -> perfect forecast, a number ->
        Gaussian noise of that standard deviation; callables are Python and cannot run in the kernel."""
        f = self.forecaster
        if f is None or (isinstance(f, str) and f == "oracle"):
            return None
        if isinstance(f, (int, float, np.integer, np.floating)) and not isinstance(f, bool):
            if f < 0:
                raise ValueError("noise_std must be non-negative")
            return ForecasterParams(noise_std=float(f), increase_uncertainty=bool(self.forecaster_increase_uncertainty),
                                    relative_noise=bool(self.forecaster_relative_noise))
        raise NotImplementedError(f"forecaster={f!r}: only None, 'oracle' and a noise standard deviation are built in "
                                  f"(user-defined forecasters are Python callables)")


class LoadModule(_TimeSeriesModule):
    """reference: modules/load_module.py:58-80."""
    module_type = ("load", "fixed")
    _default_name = "load"
    _is_sink = True

    def __init__(self, time_series, forecaster=None, forecast_horizon=DEFAULT_HORIZON, forecaster_increase_uncertainty=False,
                 forecaster_relative_noise=False, initial_step=0, final_step=-1, raise_errors=False):
        super().__init__(time_series, raise_errors, forecaster, forecast_horizon, forecaster_increase_uncertainty,
                         forecaster_relative_noise, initial_step, final_step)


class RenewableModule(_TimeSeriesModule):
    """reference: modules/renewable_module.py:60-84."""
    module_type = ("renewable", "flex")
    _default_name = "renewable"
    _is_source = True

    def __init__(self, time_series, raise_errors=False, forecaster=None, forecast_horizon=DEFAULT_HORIZON,
                 forecaster_increase_uncertainty=False, forecaster_relative_noise=False, initial_step=0, final_step=-1,
                 provided_energy_name="renewable_used"):
        if provided_energy_name != "renewable_used":
            raise NotImplementedError("provided_energy_name: the log column is fixed to 'renewable_used'")
        super().__init__(time_series, raise_errors, forecaster, forecast_horizon, forecaster_increase_uncertainty,
                         forecaster_relative_noise, initial_step, final_step)


class GridModule(_TimeSeriesModule):
    """reference: modules/grid_module.py:72-123 (constructor, `_check_params`)."""
    module_type = ("grid", "controllable")
    _default_name = "grid"
    _is_source = _is_sink = True

    def __init__(self, max_import, max_export, time_series, forecaster=None, forecast_horizon=DEFAULT_HORIZON,
                 forecaster_increase_uncertainty=False, forecaster_relative_noise=False, initial_step=0, final_step=-1,
                 cost_per_unit_co2=0.0, raise_errors=False):
        if max_import < 0:
            raise ValueError("parameter max_import must be non-negative.")
        if max_export < 0:
            raise ValueError("parameter max_export must be non-negative.")
        ts = np.asarray(time_series, dtype=np.float64)
        if ts.ndim != 2 or ts.shape[1] not in (3, 4):
            raise ValueError("Time series must be two dimensional with three or four columns.See docstring for details.")
        if ts.shape[1] == 4:
            if not ((ts[:, -1] == 0) | (ts[:, -1] == 1)).all():
                raise ValueError("Last column (grid status) must contain binary values.")
        else:
            ts = np.concatenate([ts, np.ones((ts.shape[0], 1))], axis=1)
        if (ts < 0).any():
            raise ValueError("Time series must be non-negative.")
        self.max_import, self.max_export, self.cost_per_unit_co2 = max_import, max_export, cost_per_unit_co2
        super().__init__(ts, raise_errors, forecaster, forecast_horizon, forecaster_increase_uncertainty,
                         forecaster_relative_noise, initial_step, final_step)

    def _params(self):
        return GridParams(max_import=self.max_import, max_export=self.max_export, time_series=self.time_series,
                          cost_per_unit_co2=self.cost_per_unit_co2)


class UnbalancedEnergyModule(_Module):
    """reference: modules/unbalanced_energy_module.py:13-26."""
    module_type = ("balancing", "flex")
    _default_name = "balancing"

    def __init__(self, raise_errors, initial_step=0, loss_load_cost=10, overgeneration_cost=2.0):
        self.loss_load_cost, self.overgeneration_cost = loss_load_cost, overgeneration_cost
        super().__init__(raise_errors, initial_step)


_KINDS = {"battery": BatteryModule, "genset": GensetModule, "grid": GridModule, "load": LoadModule,
          "renewable": RenewableModule, "balancing": UnbalancedEnergyModule}


def _named(modules):
    """[(name, module)] from a list of modules or (name, module) tuples (module_container.py:285-295, 355-403)."""
    out = []
    for m in modules:
        name = None
        if isinstance(m, tuple):
            name, m = m
        m = getattr(m, "_module_record", m)       # a module view of a built microgrid (`microgrid.modules.to_tuples()`)
        if not isinstance(m, _Module):
            raise TypeError(f"Module {m!r} is not one of pymgrid_b200.modules' classes")
        out.append((name if name is not None else m._default_name, m))
    return out


def params_from_modules(modules, add_unbalanced_module=True, loss_load_cost=10.0, overgeneration_cost=2.0):
    """Fold reference-style modules into the engine's parameter record.  Mirrors Microgrid.__init__ /
    _get_module_container (microgrid/microgrid.py:100-165): an UnbalancedEnergyModule with the given costs is appended
    unless add_unbalanced_module is False (then the list must contain one: the slack is part of the fused step)."""
    if isinstance(modules, (str, bytes)) or not hasattr(modules, "__iter__"):
        raise TypeError("modules must be list-like of modules.")
    named = _named(list(modules))
    if add_unbalanced_module:
        # appended un-named, so the container calls it 'balancing' (module_type[0]; microgrid.py:170-171)
        named.append(("balancing", UnbalancedEnergyModule(raise_errors=False, loss_load_cost=loss_load_cost,
                                                         overgeneration_cost=overgeneration_cost)))
    by_kind = {}
    for name, m in named:
        by_kind.setdefault(m.module_type[0], []).append((name, m))
    counts = {k: len(v) for k, v in by_kind.items()}
    need = {"load": (1, 1), "renewable": (1, 1), "battery": (1, 1), "balancing": (1, 1), "genset": (0, 1), "grid": (0, 1)}
    for kind, (lo, hi) in need.items():
        if not lo <= counts.get(kind, 0) <= hi:
            raise NotImplementedError(
                "the fused B200 step covers microgrids with exactly one load, one renewable, one battery and one "
                f"unbalanced-energy module and at most one genset and one grid; got {counts}")
    canonical = {"battery": "battery", "genset": "genset", "grid": "grid", "load": "load"}
    for kind, want in canonical.items():
        for name, _ in by_kind.get(kind, []):
            if name != want:
                raise NotImplementedError(f"module name {name!r}: only the renewable and the slack module can be renamed (the {kind} "
                                          f"module is addressed as {want!r} in controls, observations and logs)")
    # the container keeps insertion order inside a cell (module_container.py:355-413) and Microgrid.run dispatches and sums in
    # that order; the fused kernels' order is battery before grid, renewable before the slack module
    position = {m.module_type[0]: i for i, (_, m) in enumerate(named)}
    if ("grid" in position and position["grid"] < position["battery"]) or position["balancing"] < position["renewable"]:
        raise NotImplementedError("the fused B200 step dispatches battery before grid and the renewable before the slack module; "
                                  "module lists in another order run on the composed path (pymgrid_b200.Microgrid routes them)")
    if len({name for name, _ in named}) != len(named):      # module_container.py:391-396
        raise NameError("two modules share a name: " + repr(sorted(name for name, _ in named)))
    (ren_name, ren), (_, load), (_, bat), (unb_name, unb) = (by_kind[k][0] for k in ("renewable", "load", "battery", "balancing"))
    genset = by_kind["genset"][0][1] if "genset" in by_kind else None
    grid = by_kind["grid"][0][1] if "grid" in by_kind else None
    if "battery" <= ren_name <= "load":
        raise NotImplementedError(f"renewable module name {ren_name!r}: the flat observation follows gym's sorted key "
                                  "order; supported names sort before 'battery' (e.g. 'PV') or after 'load' (e.g. 'pv', 'renewable')")
    ts_modules = [m for m in (load, ren, grid) if m is not None]
    horizons = {m.forecast_horizon for m in ts_modules}
    if len(horizons) != 1:
        raise NotImplementedError(f"all time-series modules must share one forecast horizon (got {sorted(horizons)}); "
                                  "forecaster=None means horizon 0")
    if len({len(m) for m in ts_modules}) != 1:
        raise ValueError("all time series must have the same length")
    every = [m for _, m in named]
    initial = {m.initial_step for m in every}
    final = {m.final_step if m.final_step > 0 else len(m) for m in ts_modules}      # base_timeseries_module.py:317-330
    if len(initial) != 1 or len(final) != 1:
        raise ValueError("modules must agree on initial_step and final_step (microgrid.py:640-675 reads a unique value)")
    forecasters = {}
    for key, m in (("load", load), ("pv", ren), ("grid", grid)):
        f = m._forecaster_params() if m is not None else None
        if f is not None and f.noise_std != 0:
            forecasters[key] = f
    initial_step = initial.pop()
    return MicrogridParams(battery=bat._params(), genset=None if genset is None else genset._params(),
                           grid=None if grid is None else grid._params(), load_ts=load.time_series[:, 0],
                           pv_ts=ren.time_series[:, 0], loss_load_cost=unb.loss_load_cost,
                           overgeneration_cost=unb.overgeneration_cost, forecast_horizon=int(horizons.pop()),
                           initial_step=initial_step, current_step=initial_step, final_step=int(final.pop()),
                           renewable_name=ren_name, unbalanced_name=unb_name, forecasters=forecasters,
                           meta={"raise_errors": any(m.raise_errors for m in every),
                                 # per controllable module: only ITS clip raises (base_module.py:79-93, 213-221, 265-268)
                                 "raise_errors_by_module": {"genset": bool(genset.raise_errors) if genset is not None else False,
                                                            "battery": bool(bat.raise_errors),
                                                            "grid": bool(grid.raise_errors) if grid is not None else False}})
