"""Drop-in single-microgrid surface: `pymgrid.Microgrid`'s API on top of the batched engine (B = 1).

Mirrors the reference class (src/pymgrid/microgrid/microgrid.py): `Microgrid.from_scenario(n)`, `run(control,
normalized=True) -> (obs dict, reward float, done bool, info dict)`, `reset()`, `get_log()`, `sample_action()`,
`get_empty_action()`, `state_dict()`, `state_series()`, `current_step / initial_step / final_step`, `len()`,
`modules` / `fixed` / `flex` / `controllable` read-only module views.  Return values are the reference's exact Python
types; every number comes from the CUDA engine (one env, one kernel launch + a small device->host read per call -- this
wrapper exists for notebooks and control loops written against the reference; throughput lives in
`BatchedMicrogrid`).  Only what the hot path needs is mirrored: no YAML dump, no plotting, no MPC conversion.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import views
from ._cabi import FLAG_CLIP_MASK, FLAG_ERROR_MASK, FLAG_NAMES
from .engine import BatchedMicrogrid
from .params import MicrogridParams
from .scenario import load_pymgrid25


class ModuleView:
    """Read-only view of one module: constructor parameters + the live state the reference's callers read
    (max_production, max_consumption, soc, current_status, ...)."""

    def __init__(self, microgrid, kind, name=None):
        # kind: the engine's module key (load, pv, unbalanced_energy, genset, battery, grid); name: what the caller
        # called it -- the renewable (('pv', module) in pymgrid25, 'renewable' by default) and the slack module
        # ('unbalanced_energy' in pymgrid25, 'balancing' by default) carry names of their own
        self._m, self._kind, self.name = microgrid, kind, (name or kind, 0)

    def __repr__(self):
        return f"ModuleView({self.name[0]})"

    @property
    def _p(self):
        return self._m.params

    @property
    def current_step(self):
        return self._m.current_step

    # -- battery (battery_module.py:283-291) --
    def _require(self, kind):
        if self._kind != kind:
            raise AttributeError(f"{self._kind} module has no such attribute")

    @property
    def current_charge(self):
        self._require("battery")
        return self._m._state()["charge"]

    @property
    def soc(self):
        self._require("battery")
        return self._m._state()["soc"]

    @property
    def max_production(self):
        n, p, st = self._kind, self._p, self._m._state()
        if n == "battery":
            b = p.battery
            return min(b.max_discharge, st["charge"] - b.min_capacity) * b.efficiency
        if n == "genset":       # genset_module.py:466-482
            return st["genset"][0] * p.genset.running_max_production
        if n == "grid":         # grid_module.py:314-316
            return p.grid.max_import * p.grid.time_series[st["t"], 3]
        if n == "pv":
            return float(p.pv_ts[st["t"]])
        if n == "unbalanced_energy":
            return np.inf
        raise AttributeError("max_production")

    @property
    def max_consumption(self):
        n, p, st = self._kind, self._p, self._m._state()
        if n == "battery":
            b = p.battery
            return min(b.max_charge, b.max_capacity - st["charge"]) / b.efficiency
        if n == "grid":
            return p.grid.max_export * p.grid.time_series[st["t"], 3]
        if n == "load":
            return -1 * float(p.load_ts[st["t"]])
        if n == "unbalanced_energy":
            return np.inf
        raise AttributeError("max_consumption")

    @property
    def min_production(self):
        if self._kind == "genset":
            return self._m._state()["genset"][0] * self._p.genset.running_min_production
        return 0

    @property
    def current_status(self):
        if self._kind == "genset":
            return self._m._state()["genset"][0]
        if self._kind == "grid":
            return self._p.grid.time_series[self._m.current_step, 3]
        raise AttributeError("current_status")

    @property
    def goal_status(self):
        self._require("genset")
        return self._m._state()["genset"][1]

    # -- typing (module_container.py:355-413 sorts modules by these) --
    @property
    def module_type(self):
        """(class tag, dispatch type) like the reference's class attribute, e.g. ('load', 'fixed')"""
        tag = {"pv": "renewable", "unbalanced_energy": "balancing"}.get(self._kind, self._kind)
        return (tag, {"load": "fixed", "pv": "flex", "unbalanced_energy": "flex"}.get(self._kind, "controllable"))

    @property
    def is_source(self):
        return self._kind != "load"

    @property
    def is_sink(self):
        return self._kind in ("load", "battery", "grid", "unbalanced_energy")

    @property
    def action_space(self):
        """only `.shape` is read by the reference's callers (priority_list.py:27-33)"""
        from types import SimpleNamespace
        n = {"genset": 2, "battery": 1, "grid": 1, "pv": 1, "unbalanced_energy": 1}.get(self._kind, 0)
        return SimpleNamespace(shape=(n,))

    # -- costs (base_module.py:651-670 and the per-module overrides) --
    @property
    def production_marginal_cost(self):
        n, p = self._kind, self._p
        if n == "battery":
            return p.battery.battery_cost_cycle                           # battery_module.py:340-342
        if n == "genset":                                                 # genset_module.py:519-521: get_cost(1.0)
            return p.genset.genset_cost * 1.0 + p.genset.cost_per_unit_co2 * (p.genset.co2_per_unit * 1.0)
        if n == "grid":
            return self.import_price[0]                                   # grid_module.py:322-324
        if n == "unbalanced_energy":
            return p.loss_load_cost                                       # unbalanced_energy_module.py:111-113
        return 0.0

    @property
    def absorption_marginal_cost(self):
        n, p = self._kind, self._p
        if n == "battery":
            return p.battery.battery_cost_cycle
        if n == "grid":
            return self.export_price[0]
        if n == "unbalanced_energy":
            return p.overgeneration_cost
        return 0.0

    @property
    def marginal_cost(self):
        return self.production_marginal_cost

    # -- genset look-ahead (genset_module.py:360-424) --
    def next_status(self, goal_status):
        self._require("genset")
        cs, _, up, dn = self._m._state()["genset"]
        if goal_status:
            return 1 if (cs or up == 0) else 0
        return 0 if (not cs or dn == 0) else 1

    def next_max_production(self, goal_status):
        return self.next_status(goal_status) * self._p.genset.running_max_production

    def next_min_production(self, goal_status):
        return self.next_status(goal_status) * self._p.genset.running_min_production

    # -- state vectors (BaseMicrogridModule.state / state_dict, base_module.py:535-560) --
    def state_dict(self, normalized=False):
        st = self._m._state()
        d = views.state_dict(self._p, st["t"], st["charge"], st["genset"], st["soc"])[self._kind]
        if normalized:
            return OrderedDict(zip(d.keys(), self.to_normalized(np.array(list(d.values()), dtype=np.float64), obs=True)))
        return d

    @property
    def state(self):
        return np.array(list(self.state_dict().values()), dtype=np.float64)

    @property
    def min_obs(self):
        return self._bounds("obs")[0]

    @property
    def max_obs(self):
        return self._bounds("obs")[1]

    @property
    def min_act(self):
        return self._bounds("act")[0]

    @property
    def max_act(self):
        return self._bounds("act")[1]

    def _bounds(self, which):
        n, p, rows = self._kind, self._p, 1 + self._p.forecast_horizon
        if n == "battery":      # battery_module.py:323-338
            b = p.battery
            return ((np.array([b.min_soc, b.min_capacity]), np.array([1.0, b.max_capacity])) if which == "obs"
                    else (b.min_act, b.max_act))
        if n == "genset":       # genset_module.py:503-517
            g = p.genset
            return ((np.zeros(4), np.array([1.0, 1.0, g.start_up_time, g.wind_down_time])) if which == "obs"
                    else (np.array([0.0, 0.0]), np.array([1.0, g.running_max_production])))
        if n == "grid":         # grid_module.py:125-132
            ts = p.grid.effective_time_series()
            return ((np.tile(ts.min(axis=0), rows), np.tile(ts.max(axis=0), rows)) if which == "obs"
                    else (-1 * p.grid.max_export, p.grid.max_import))
        if n in ("load", "pv"):  # base_timeseries_module.py:81-88
            ts = (p.load_ts * p.load_scale) if n == "load" else (p.pv_ts * p.pv_scale)
            lo, hi = views.series_bounds(ts, True)
            return (np.full(rows, lo), np.full(rows, hi)) if which == "obs" else ((lo, hi) if n == "pv" else (np.array([]), np.array([])))
        return (np.array([]), np.array([])) if which == "obs" else (-np.inf, np.inf)

    def to_normalized(self, value, act=False, obs=False):
        """reference: BaseMicrogridModule.to_normalized -> ModuleSpace.normalize (utils/space.py:207-218)"""
        assert act + obs == 1, "One of act or obs must be True but not both."
        low, high = self._bounds("act" if act else "obs")
        spread = np.asarray(high, dtype=np.float64) - np.asarray(low, dtype=np.float64)
        spread = np.where(spread == 0, 1.0, spread)
        return (np.asarray(value, dtype=np.float64) - low) / spread

    def from_normalized(self, value, act=False, obs=False):
        """reference: ModuleSpace.denormalize (utils/space.py:220-231)"""
        assert act + obs == 1, "One of act or obs must be True but not both."
        low, high = self._bounds("act" if act else "obs")
        spread = np.asarray(high, dtype=np.float64) - np.asarray(low, dtype=np.float64)
        spread = np.where(spread == 0, 1.0, spread)
        return low + spread * np.asarray(value, dtype=np.float64)

    # -- grid columns of the current state (grid_module.py:248-299: state[k::4], current + forecast) --
    def _grid_column(self, k):
        self._require("grid")
        return self.state[k::4]

    @property
    def import_price(self):
        return self._grid_column(0)

    @property
    def export_price(self):
        return self._grid_column(1)

    @property
    def co2_per_kwh(self):
        return self._grid_column(2)

    @property
    def grid_status(self):
        return self._grid_column(3)

    @property
    def min_soc(self):
        self._require("battery")
        return self._p.battery.min_soc

    @property
    def max_soc(self):
        self._require("battery")
        return 1.0

    def __getattr__(self, item):   # constructor parameters, e.g. battery.max_capacity, genset.genset_cost, grid.max_import
        src = {"battery": self._p.battery, "genset": self._p.genset, "grid": self._p.grid}.get(self._kind)
        if src is not None and hasattr(src, item) and not item.startswith("_"):
            return getattr(src, item)
        if item == "time_series":
            return {"load": self._p.load_ts.reshape(-1, 1) * self._p.load_scale,
                    "pv": self._p.pv_ts.reshape(-1, 1) * self._p.pv_scale}[self._kind]
        if item == "forecast_horizon" and self._kind in ("load", "pv", "grid"):
            return self._p.forecast_horizon
        if item in ("initial_step", "final_step"):
            # a module's own window: what trajectory_func set at the last reset (microgrid.py:221-225, 652-684), else the microgrid's
            window = getattr(self._m, "_module_window", None)
            return window[item == "final_step"] if window is not None else getattr(self._m, item)
        if item in ("loss_load_cost", "overgeneration_cost") and self._kind == "unbalanced_energy":
            return getattr(self._p, item)
        raise AttributeError(item)


class ModuleContainerView(OrderedDict):
    """`microgrid.modules`: name -> [module]; attribute access and the reference's iteration helpers."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item)

    def iterdict(self):
        return self.items()

    def iterlist(self):
        return [m for lst in self.values() for m in lst]

    to_list = iterlist

    def _filtered(self, keep):
        return ModuleContainerView((n, lst) for n, lst in self.items() if keep(lst[0]))

    # the reference's second container level (module_container.py:405-411): sources / sinks / source_and_sinks
    @property
    def sources(self):
        return self._filtered(lambda m: m.is_source and not m.is_sink)

    @property
    def sinks(self):
        return self._filtered(lambda m: m.is_sink and not m.is_source)

    @property
    def source_and_sinks(self):
        return self._filtered(lambda m: m.is_source and m.is_sink)

    def to_dict(self):
        return dict(self)

    # the reference's first container level (module_container.py:405-411): fixed / flex / controllable
    @property
    def fixed(self):
        return self._filtered(lambda m: m.module_type[1] == "fixed")

    @property
    def flex(self):
        return self._filtered(lambda m: m.module_type[1] == "flex")

    @property
    def controllable(self):
        return self._filtered(lambda m: m.module_type[1] == "controllable")

    def names(self):
        return list(self.keys())

    def to_tuples(self):
        return [(name, m) for name, lst in self.items() for m in lst]

    def get_attrs(self, *attrs, unique=False, as_pandas=True):
        """reference: Container.get_attrs (module_container.py:97-195): the given attributes of every module that has them;
        unique=True returns the single value each attribute takes (ValueError when the modules disagree)"""
        return container_get_attrs(self, attrs, unique, as_pandas)


def container_get_attrs(container, attrs, unique, as_pandas):
    import pandas as pd
    rows = OrderedDict()
    for name, lst in container.items():
        for j, m in enumerate(lst):
            rows[(name, j)] = {a: getattr(m, a) for a in attrs if _has_attr(m, a)}
    missing = [a for a in attrs if not any(a in r for r in rows.values())]
    if missing:
        raise AttributeError(f'No values found for key(s) {missing}')
    if unique:
        out = {}
        for a in attrs:
            vals = [r[a] for r in rows.values() if a in r]
            if any(v != vals[0] for v in vals[1:]):
                raise ValueError(f"Attribute(s) {[a]} have non-unique values, cannot return single unique value.")
            out[a] = vals[0]
        return pd.Series(out) if as_pandas else out
    if as_pandas:
        return pd.DataFrame.from_dict({k: r for k, r in rows.items() if r}, orient="index")
    return {name: [rows[(name, j)] for j in range(len(lst))] for name, lst in container.items()}


def _has_attr(m, a):
    try:
        getattr(m, a)
        return True
    except (AttributeError, KeyError, IndexError):
        return False


def raise_for_flags(flags, params, raise_errors=False, step=None):
    """Turn the engine's per-env event flags of one step into the exception the reference raises at that point
    (include/pymgrid_b200.h, MG_FLAG_*): used by the single-microgrid surfaces (Microgrid.run, the env classes with one env)."""
    if flags & (1 << 5):        # load_module.py:111: the time series is indexed past its end
        where = "" if step is None else f"index {step} is out of bounds for axis 0 with size {len(params)}"
        raise IndexError(where or "step past the end of the time series")
    err = flags & FLAG_ERROR_MASK
    if err:
        names = [n for bit, n in FLAG_NAMES.items() if err & bit]
        if err & (1 << 1):      # a genset asked to absorb: as_sink compares with max_consumption, which a source-only
            # module does not implement (base_module.py:265, :604-619)
            raise TypeError("'>' not supported between instances of 'float' and 'NotImplementedType'")
        if err & (1 << 2):
            raise RuntimeError("Microgrid modules unable to balance energy production with consumption.\n")
        raise AssertionError(f"step rejected: {names}")
    mask = FLAG_CLIP_MASK
    by_module = params.meta.get("raise_errors_by_module")
    if by_module is not None:       # built from modules: only a module constructed with raise_errors=True raises for ITS clip
        mask = sum(bit for bit, key in ((1 << 8, "genset"), (1 << 9, "battery"), (1 << 10, "grid")) if by_module.get(key))
    if raise_errors and flags & mask:
        names = [n for bit, n in FLAG_NAMES.items() if flags & mask & bit]
        raise ValueError(f"requested value outside the module's limits: {names}")    # base_module.py:79-93


class Microgrid:
    def __new__(cls, modules=None, *args, **kw):
        """Module lists outside the fused kernels' scope (several loads / renewables / batteries / gensets / grids, no
        battery, per-module horizons, renamed modules ...) are served by the composed path: same surface, general
        dispatch kernel (compose.py, include/pymgrid_b200_compose.h)."""
        if cls is Microgrid and isinstance(modules, (list, tuple)):
            from .compose import ComposedMicrogrid, in_fused_scope
            add_unbalanced = args[0] if args else kw.get("add_unbalanced_module", True)
            if not in_fused_scope(modules, add_unbalanced):
                return ComposedMicrogrid(modules, *args, **kw)
        return super().__new__(cls)

    def __init__(self, modules, add_unbalanced_module=True, loss_load_cost=10., overgeneration_cost=2.,
                 reward_shaping_func=None, trajectory_func=None, device=None, obs_order="gym_sorted"):
        """The reference's constructor (microgrid/microgrid.py:100-128):

        `modules`: list of `pymgrid_b200.modules` objects or `(name, module)` tuples, as in the reference -- or a ready
        `MicrogridParams` record (scenario readers; then the three arguments after it are not used).
        `reward_shaping_func`: None, "pv_curtailment" / "battery_discharge", or an object of the reference's
        PVCurtailmentShaper / BatteryDischargeShaper classes (matched by class name); arbitrary Python callables cannot
        run inside the kernel and are rejected.
        `trajectory_func(initial_step, final_step) -> (initial, final)`: called on every `reset()` (microgrid.py:221-225),
        validated like `_check_trajectory_func` (:167-199).
        `device`, `obs_order`: engine options (not in the reference)."""
        if isinstance(modules, MicrogridParams):
            params = modules
        else:
            from .modules import params_from_modules
            params = params_from_modules(modules, add_unbalanced_module, loss_load_cost, overgeneration_cost)
        if reward_shaping_func is not None:
            import dataclasses
            name = reward_shaping_func if isinstance(reward_shaping_func, str) else type(reward_shaping_func).__name__
            params = dataclasses.replace(params, reward_shaper=name)
        self.params = params
        self._obs_order = obs_order
        self._engine = BatchedMicrogrid([params], np.zeros(1, dtype=np.int64), device=device, obs_order=obs_order,
                                        with_info=True, with_flags=True, action_order=views.CONTROL_ORDER)
        self._g = self._engine.groups[0]
        self._actions = torch.zeros((1, params.n_act), dtype=torch.float64, device=self._engine.device)
        self._log_rows = []
        self._soc0 = params.battery.soc     # the soc the battery was constructed with, reported until its first update
        self._initial_step, self._final_step = params.initial_step, params.final_step
        self.raise_errors = bool(params.meta.get("raise_errors", False))
        names = ["load", "pv", "unbalanced_energy"] + (["genset"] if params.has_genset else []) + ["battery"] + \
                (["grid"] if params.has_grid else [])
        self._modules = ModuleContainerView((self._caller_name(n), [ModuleView(self, n, self._caller_name(n))]) for n in names)
        self.trajectory_func = self._check_trajectory_func(trajectory_func)

    @property
    def _ren(self):
        """the caller's name of the renewable module ('pv' in pymgrid25, 'PV' in MicrogridGenerator grids, 'renewable' by default)"""
        return self.params.renewable_name

    def _check_trajectory_func(self, trajectory_func):
        """reference: Microgrid._check_trajectory_func (microgrid.py:167-199): see trajectory.validated"""
        from .trajectory import validated
        return validated(trajectory_func, self._initial_step, self._final_step)

    def _caller_name(self, kind):
        """engine key -> the caller's name of that module: only the renewable ('pv' in pymgrid25, 'renewable' by default,
        'PV' in MicrogridGenerator grids) and the slack module ('unbalanced_energy' in pymgrid25, 'balancing' when
        Microgrid(modules) appends it) carry names of their own"""
        return {"pv": self.params.renewable_name, "unbalanced_energy": self.params.unbalanced_name}.get(kind, kind)

    def _named(self, d):
        """engine keys -> caller names, for dicts keyed by module name or by (module name, number, field)"""
        if self._ren == "pv" and self.params.unbalanced_name == "unbalanced_energy":
            return d
        nm = self._caller_name
        return type(d)(((nm(k) if not isinstance(k, tuple) else (nm(k[0]),) + k[1:]), v) for k, v in d.items())

    # ---- construction ------------------------------------------------------------------------------------
    @classmethod
    def from_scenario(cls, microgrid_number=0, **kw):
        """reference: Microgrid.from_scenario (microgrid.py:958-980) -- pymgrid25 benchmark grid n."""
        return cls(load_pymgrid25(microgrid_number), **kw)

    # ---- state -------------------------------------------------------------------------------------------
    def _state(self):
        g = self._g
        gen = tuple(int(x) for x in self._engine.genset_status(0)[0].tolist()) if g.genset is not None else (0, 0, 0, 0)
        charge = float(g.charge[0].item())
        soc = charge / self.params.battery.max_capacity if self._soc0 is None else self._soc0     # battery_module.py:89, 130
        return dict(t=int(g.step[0].item()), charge=charge, genset=gen, soc=soc)

    @property
    def current_step(self):
        return int(self._g.step[0].item())

    @property
    def initial_step(self):
        return self._initial_step

    @initial_step.setter
    def initial_step(self, value):
        self._initial_step = int(value)
        self._engine.set_trajectories(np.array([self._initial_step]), np.array([self._final_step]))

    @property
    def final_step(self):
        return self._final_step

    @final_step.setter
    def final_step(self, value):
        self._final_step = int(value)
        self._engine.set_trajectories(np.array([self._initial_step]), np.array([self._final_step]))

    def __len__(self):
        return len(self.params)

    @property
    def modules(self):
        return self._modules

    def _typed(self, kinds):
        return ModuleContainerView((n, m) for n, m in self._modules.items() if n in kinds)

    @property
    def fixed(self):
        return self._typed(("load",))

    @property
    def flex(self):
        return self._typed((self._ren, self.params.unbalanced_name))

    @property
    def controllable(self):
        return ModuleContainerView((n, self._modules[n]) for n in views.CONTROL_ORDER if n in self._modules)

    # ---- the hot path ------------------------------------------------------------------------------------
    def run(self, control, normalized=True):
        """reference: Microgrid.run (microgrid.py:227-325).  Same arguments, return types, errors."""
        p = self.params
        row = views.control_dict_to_row(control, p, self._g.act_cols)
        pre = self._state()
        if pre["t"] >= len(p):
            raise IndexError(f"index {pre['t']} is out of bounds for axis 0 with size {len(p)}")   # load_module.py:111
        self._actions.copy_(torch.from_numpy(row).reshape(1, -1))
        obs, reward, done, info = self._engine.step(self._actions, normalized=normalized)
        self._soc0 = None
        flags = int(self._g.flags[0].item()) & 0xffffffff
        self._raise_for_flags(flags)
        obs_row, info_row = obs[0].cpu().numpy(), info[0].cpu().numpy()
        r = float(reward[0].item())
        post = self._state()
        row = self._named(views.log_row(p, views.state_dict(p, pre["t"], pre["charge"], pre["genset"], pre["soc"]),
                                        info_row, r, post["genset"]))
        stale = self.__dict__.pop("_stale_forecast", None)
        self._log_rows.append(row if stale is None else views.drop_stale_forecasts(row, stale))
        return (self._named(views.obs_row_to_dict(obs_row, p, self._obs_order)), r, bool(done[0].item()),
                self._named(views.info_row_to_dict(info_row, flags, p)))

    def run_priority_list(self, action_index, n_steps):
        """`n_steps` consecutive DiscreteMicrogridEnv-style steps with the same priority list (index into
        `engine.action_tables[0]`): what RuleBasedControl.run does every step (algos/rbc/rbc.py:87-91 ->
        priority_list.py:69-116 -> Microgrid.run(normalized=False)), expanded on the device.  The log rows of the whole
        episode are gathered on the device and brought back with ONE device->host copy, so a year costs seconds, not
        minutes.  Returns the number of steps taken (stops after the step that reports done)."""
        p, g = self.params, self._g
        t0 = self.current_step
        windows = (g.env_final_step is not None)
        final = int(g.env_final_step[0].item()) if windows else self._final_step
        n = max(0, min(int(n_steps), final - t0 if final - 1 >= t0 else 1, len(p) - t0))
        if n == 0:
            if n_steps > 0:
                raise IndexError(f"index {t0} is out of bounds for axis 0 with size {len(p)}")   # load_module.py:111
            return 0
        act = torch.full((1,), int(action_index), dtype=torch.int32, device=self._engine.device)
        soc0, self._soc0 = self._soc0, None
        pre_t, pre_charge, pre_gen, post_gen, infos, rewards, flags = [], [], [], [], [], [], []
        zero = torch.zeros(1, dtype=torch.int32, device=self._engine.device)
        gen = (lambda: g.genset.clone()) if g.genset is not None else (lambda: zero)
        for _ in range(n):
            pre_t.append(g.step.clone()); pre_charge.append(g.charge.clone()); pre_gen.append(gen())
            _, reward, _, info = self._engine.step_discrete(act, obs=False)
            infos.append(info.clone()); rewards.append(reward.clone()); post_gen.append(gen()); flags.append(g.flags.clone())
        host = lambda xs: torch.cat(xs).cpu().numpy()      # noqa: E731  (one device->host copy per column)
        pre_t, pre_charge, pre_gen, post_gen = host(pre_t), host(pre_charge), host(pre_gen), host(post_gen)
        infos, rewards, flags = torch.stack(infos).cpu().numpy()[:, 0], host(rewards), host(flags)
        unpack = lambda w: (int(w) & 0xff, (int(w) >> 8) & 0xff, (int(w) >> 16) & 0xff, (int(w) >> 24) & 0xff)   # noqa: E731
        for k in range(n):
            self._raise_for_flags(int(flags[k]) & 0xffffffff)
            state = views.state_dict(p, int(pre_t[k]), float(pre_charge[k]), unpack(pre_gen[k]), soc0 if k == 0 else None)
            self._log_rows.append(self._named(views.log_row(p, state, infos[k], float(rewards[k]), unpack(post_gen[k]))))
        return n

    def _raise_for_flags(self, flags):
        raise_for_flags(flags, self.params, self.raise_errors)

    def reset(self):
        """reference: Microgrid.reset (microgrid.py:205-225): step = initial_step, logs flushed, battery / genset kept."""
        if self.trajectory_func is not None:      # microgrid.py:221-225: the modules' window, not the microgrid's own bounds
            initial_step, final_step = self.trajectory_func(self._initial_step, self._final_step)
            self._engine.set_trajectories(np.array([initial_step]), np.array([final_step]))
            self._module_window = (int(initial_step), int(final_step))
        obs = self._engine.reset()
        flushed, self._log_rows = views.flushed_balance_log(self._log_rows), []
        by_name = self._named(views.obs_row_to_dict(obs[0].cpu().numpy(), self.params, self._obs_order))
        # reset() lists the modules in CONTAINER order (fixed, flex, controllable -- `modules.to_dict()`, microgrid.py:217-219),
        # run() in dispatch order
        out = type(by_name)((name, by_name[name]) for name in self._modules)
        out["balance"], out["other"] = flushed, {}
        return out

    # ---- actions -----------------------------------------------------------------------------------------
    def sample_action(self, strict_bound=False, sample_flex_modules=False):
        """reference: Microgrid.sample_action (microgrid.py:337-362): np.random.rand() per controllable module, genset
        first (goal, energy).  `strict_bound` is only defined for grids without a genset, like in the reference."""
        if strict_bound and self.params.has_genset:
            raise TypeError("Unable to normalize scalar value, expected array-like of shape 2")   # utils/space.py:146-147
        # microgrid.py:358-362: every module with an action space, in the container's listing order (flex modules first
        # when they are asked for), one np.random.rand() each (genset: goal, energy) -- base_module.py:326-356
        names = list(self._modules) if sample_flex_modules else [self._caller_name(n) for n in views.control_names(self.params)]
        out = {}
        for name in names:
            m = self._modules[name][0]
            if not m.action_space.shape[0]:
                continue
            if m._kind == "genset":
                out[name] = [np.array([np.random.rand(), np.random.rand()])]
                continue
            lo, hi = 0.0, 1.0
            if strict_bound:
                with np.errstate(invalid="ignore"):
                    act_lo, act_hi = (float(x) for x in m._bounds("act"))
                    spread = (act_hi - act_lo) or 1.0
                    if m.is_sink:
                        lo = (-1 * m.max_consumption - act_lo) / spread
                        lo = 0 if np.isnan(lo) else lo
                    if m.is_source:
                        hi = (m.max_production - act_lo) / spread
                        hi = 0 if np.isnan(hi) else hi
            out[name] = [np.random.rand() * (hi - lo) + lo]
        return out

    def export_params(self):
        """A copy of the parameter record carrying the LIVE state (step, battery charge, genset tuple): what the
        reference's deep copies of a running microgrid hold (e.g. BaseMicrogridEnv.from_microgrid, envs/base/base.py:270-290)."""
        import copy
        st = self._state()
        p = copy.deepcopy(self.params)
        p.current_step, p.initial_step, p.final_step = st["t"], self._initial_step, self._final_step
        p.battery.current_charge, p.battery.soc = st["charge"], self._soc0
        if p.genset is not None:
            g = p.genset
            g.current_status, g.goal_status, g.steps_until_up, g.steps_until_down = st["genset"]
        return p

    def __getattr__(self, item):
        """`microgrid.<module name>` (reference: Microgrid.__getattr__, microgrid.py:1023-1030)"""
        if item.startswith("_"):
            raise AttributeError(item)
        mods = self.__dict__.get("_modules")
        if mods is not None and item in mods:
            return mods[item]
        raise AttributeError(item)

    def get_cost_info(self):
        """reference: Microgrid.get_cost_info (microgrid.py:334-335)"""
        return {name: [dict(production_marginal_cost=m.production_marginal_cost, absorption_marginal_cost=m.absorption_marginal_cost)
                       for m in lst] for name, lst in self._modules.items()}

    def set_forecaster(self, forecaster, forecast_horizon=None, forecaster_increase_uncertainty=False,
                       forecaster_relative_noise=False):
        """reference: Microgrid.set_forecaster (microgrid.py:477-546): None (no forecast, horizon 0), "oracle", or a noise
        standard deviation, for every time-series module (BASELINE config 4 is `set_forecaster('oracle',
        forecast_horizon=24)`).  The engine is rebuilt around the live state; the log is kept."""
        import dataclasses
        from .params import DEFAULT_HORIZON, ForecasterParams
        if forecast_horizon is None:
            forecast_horizon = DEFAULT_HORIZON
        ts_names = {"load": "load", self._ren: "pv"}
        if self.params.has_grid:
            ts_names["grid"] = "grid"
        if isinstance(forecaster, dict):
            # the reference's dict branch (microgrid.py:520-533) calls set_forecaster on the module LIST of each name and
            # swallows the AttributeError that raises: names are checked, nothing else happens.  Mirrored.
            for name in forecaster:
                if name not in self._modules:
                    raise NameError(f'Unrecognized module {name}.')
            return
        settings = {key: forecaster for key in ts_names.values()}
        horizons = {forecast_horizon * (f is not None) for f in settings.values()}       # base_timeseries_module.py:237
        if len(horizons) != 1:
            raise NotImplementedError("one forecast horizon for all time-series modules on the fused path")
        noise = {}
        for key, f in settings.items():
            if f is None or (isinstance(f, str) and f == "oracle"):
                continue
            if isinstance(f, (int, float, np.integer, np.floating)) and not isinstance(f, bool):
                if f < 0:
                    raise ValueError("noise_std must be non-negative")
                if f != 0:
                    noise[key] = ForecasterParams(float(f), bool(forecaster_increase_uncertainty), bool(forecaster_relative_noise))
                continue
            raise NotImplementedError(f"forecaster={f!r}: only None, 'oracle' and a noise standard deviation are built in "
                                      f"(user-defined forecasters are Python callables)")
        old_horizon = self.params.forecast_horizon
        params = dataclasses.replace(self.export_params(), forecast_horizon=int(horizons.pop()), forecasters=noise)
        keep = (self._log_rows, self.trajectory_func, self.raise_errors, self._soc0, self._initial_step, self._final_step)
        Microgrid.__init__(self, params, device=self._engine.device, obs_order=self._obs_order)
        self._log_rows, self.trajectory_func, self.raise_errors, self._soc0, self._initial_step, self._final_step = keep
        # the next step still logs the forecast computed before the change (views.drop_stale_forecasts)
        self._stale_forecast = {(name, 0): old_horizon for name in ts_names}

    def set_module_attr(self, attr_name, value):
        """reference: Microgrid.set_module_attr (microgrid.py:584-612): set a constructor attribute on every module that
        has it -- 'forecast_horizon' on the time-series modules, a parameter of the battery / genset / grid records
        (`max_capacity`, `genset_cost`, `max_import` ...); AttributeError when no module has it.  The engine is rebuilt
        around the live state; the log is kept."""
        import dataclasses
        params = self.export_params()
        hit = False
        if attr_name == "forecast_horizon":
            params, hit = dataclasses.replace(params, forecast_horizon=int(value)), True
        for part in ("battery", "genset", "grid"):
            rec = getattr(params, part)
            if rec is not None and attr_name in {f.name for f in dataclasses.fields(rec)} and attr_name != "time_series":
                setattr(rec, attr_name, value)
                hit = True
        if attr_name in ("loss_load_cost", "overgeneration_cost"):
            params, hit = dataclasses.replace(params, **{attr_name: value}), True
        if not hit:
            raise AttributeError(f"No module has attribute '{attr_name}'.")
        keep = (self._log_rows, self.trajectory_func, self.raise_errors, self._soc0, self._initial_step, self._final_step)
        Microgrid.__init__(self, params, device=self._engine.device, obs_order=self._obs_order)
        self._log_rows, self.trajectory_func, self.raise_errors, self._soc0, self._initial_step, self._final_step = keep

    def get_forecast_horizon(self):
        """reference: Microgrid.get_forecast_horizon (microgrid.py:364-388)"""
        return self.params.forecast_horizon

    def to_normalized(self, data_dict, act=False, obs=False):
        """reference: Microgrid.to_normalized (microgrid.py:390-410): {name: [values]} through each module's space"""
        return {name: [self._modules[name][0].to_normalized(v, act=act, obs=obs) for v in vals] for name, vals in data_dict.items()}

    def from_normalized(self, data_dict, act=False, obs=False):
        return {name: [self._modules[name][0].from_normalized(v, act=act, obs=obs) for v in vals] for name, vals in data_dict.items()}

    def get_empty_action(self, sample_flex_modules=False):
        names = list(self._modules) if sample_flex_modules else [self._caller_name(n) for n in views.control_names(self.params)]
        return {name: [None] for name in names if self._modules[name][0].action_space.shape[0]}

    # ---- introspection -----------------------------------------------------------------------------------
    def state_dict(self, normalized=False):
        """reference: Microgrid.state_dict (microgrid.py:412-421): {module name: [module.state_dict(normalized)]}"""
        if normalized:      # every module through its own observation space (base_module.py:65-77, utils/space.py:207-218)
            return {name: [dict(views_list[0].state_dict(normalized=True))] for name, views_list in self._modules.items()}
        st = self._state()
        sd = views.state_dict(self.params, st["t"], st["charge"], st["genset"], st["soc"])
        return {name: [dict(d)] for name, d in self._named(sd).items()}

    def state_series(self, normalized=False):
        """reference: Microgrid.state_series (microgrid.py:423-432): the same values as one MultiIndex Series"""
        import pandas as pd
        if normalized:
            sd = {name: d[0] for name, d in self.state_dict(normalized=True).items()}
        else:
            st = self._state()
            sd = self._named(views.state_dict(self.params, st["t"], st["charge"], st["genset"], st["soc"]))
        data = OrderedDict(((name, 0, k), v) for name, d in sd.items() for k, v in d.items())
        return pd.Series(data)

    def get_log(self, as_frame=True, drop_singleton_key=False):
        """reference: Microgrid.get_log (microgrid.py:434-475): one row per step since the last reset."""
        df = views.log_frame(self._log_rows, self.current_step, drop_singleton_key)
        return df if as_frame else df.to_dict()

    @property
    def log(self):
        return self.get_log()

    def __repr__(self):
        return "Microgrid([" + ", ".join(f"{n} x 1" for n in self._modules) + "])"
