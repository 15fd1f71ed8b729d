"""Composed microgrids: `Microgrid.run` for ANY module list, on the GPU (C-ABI: include/pymgrid_b200_compose.h).

The fused engine (`engine.BatchedMicrogrid`) covers the module set of every pymgrid25 / MicrogridGenerator grid.  The
reference's `Microgrid` (src/pymgrid/microgrid/microgrid.py:100-325) takes any list of modules -- several loads and
renewables (its own balance tests, tests/microgrid/test_microgrid.py:188-455), several batteries / gensets / grids, no
battery, no slack module, one forecast horizon per time-series module.  This module is the host side of that general
path:

  Composition        the static structure of a module list: the container's listing and dispatch orders
                     (module_container.py:355-413), action columns, observation blocks, parameter packing
  ComposedBatch      B microgrids sharing one composition (own parameters, series, state): device tensors + mgc_* calls
  ComposedMicrogrid  the reference's single-microgrid surface (run / reset / get_log / state_dict / sample_action / modules,
                     set_forecaster, Python reward shapers / trajectory functions) on a batch of one;
                     `pymgrid_b200.Microgrid(modules)` returns it when the list is outside the fused scope
  ComposedDiscreteEnv / ComposedContinuousEnv / ComposedRuleBasedControl
                     the env and controller surfaces (priority lists expanded on the device: mgc_run_discrete)
  ComposedLogRecorder  reference-format log for a subset of a batch
  StandaloneModule   a module stepped on its own (mgc_modules_step): what `module.step(action)` of pymgrid_b200.modules runs on

No CPU fallback: without the CUDA extension or a CUDA device, construction raises (`_open_device_library`).
"""
import ctypes as C
import warnings
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi, views
from .modules import UnbalancedEnergyModule, _named

MGC_ABI_VERSION = 1
MGC_MAX_MODULES = 64
MGC_INFO_SLOTS = 5
MGC_BALANCE_SLOTS = 6
MGC_CFG_HEADER = 2
KIND = {"load": 0, "renewable": 1, "battery": 2, "genset": 3, "grid": 4, "balancing": 5}
PARAM_COUNT = {"load": 3, "renewable": 3, "battery": 6, "genset": 8, "grid": 12, "balancing": 2}
FLAG_GENSET_GOAL_RANGE, FLAG_NOT_A_SINK, FLAG_BALANCE, FLAG_BATTERY_MIN_CAP = 1 << 0, 1 << 1, 1 << 2, 1 << 3
FLAG_NEGATIVE_ABSORB, FLAG_STEP_PAST_END, FLAG_CLIP, FLAG_CLIP_RAISES, FLAG_EXCESS = 1 << 4, 1 << 5, 1 << 8, 1 << 11, 1 << 14

_i32, _vp = C.c_int32, C.c_void_p


class MgcModule(C.Structure):
    _fields_ = [(n, _i32) for n in ("kind", "horizon", "act_col", "obs_off", "param_off", "fstate_off", "istate_off",
                                    "listing", "raise_errors")]


class MgcLayout(C.Structure):
    _fields_ = [("abi_version", _i32), ("n_modules", _i32), ("modules", MgcModule * MGC_MAX_MODULES),
                ("n_act", _i32), ("obs_dim", _i32), ("n_fstate", _i32), ("n_istate", _i32), ("cfg_stride", _i32),
                ("n_cfg", _i32), ("series_len", _i32), ("n_series", _i32), ("n_envs", C.c_int64),
                ("cfg", _vp), ("series", _vp), ("series_off", _vp), ("series_nrm", _vp), ("step", _vp), ("fstate", _vp), ("istate", _vp),
                ("cfg_index", _vp), ("plist", _vp), ("n_plist", _i32), ("plist_width", _i32),
                ("env_initial_step", _vp), ("env_final_step", _vp), ("obs_select", _vp)]


class MgcIO(C.Structure):
    _fields_ = [("actions", _vp), ("obs", _vp), ("reward", _vp), ("done", _vp), ("info", _vp), ("flags", _vp), ("mask", _vp),
                ("dactions", _vp), ("dactions_const", C.c_int64)]


EXPORTED_SYMBOLS = ("mgc_abi_version", "mgc_sizeof", "mgc_param_count", "mgc_create", "mgc_destroy", "mgc_run",
                    "mgc_run_discrete", "mgc_modules_step", "mgc_reset", "mgc_observe", "mgc_forecast_noise", "mgc_launch_count")
MAX_PRIORITY_ELEMENTS = 8       # (elements)! permutations are enumerated like the reference does (priority_list.py:15-38)
FLAG_BAD_ACTION = 1 << 6


def bind(L):
    """argtypes + ABI checks on a loaded library (libpymgrid_b200.so; the test suite passes its host build)"""
    if getattr(L, "_mgc_bound", False):
        return L
    L.mgc_abi_version.restype = C.c_int
    L.mgc_sizeof.restype, L.mgc_sizeof.argtypes = C.c_int64, [C.c_int]
    L.mgc_param_count.restype, L.mgc_param_count.argtypes = _i32, [C.c_int]
    L.mgc_create.argtypes = [C.POINTER(MgcLayout), C.POINTER(_vp)]
    L.mgc_destroy.argtypes = [_vp]
    L.mgc_run.argtypes = [_vp, C.POINTER(MgcIO), _i32, _i32, C.c_int, _vp]
    L.mgc_run_discrete.argtypes = [_vp, C.POINTER(MgcIO), _i32, _i32, _vp]
    L.mgc_modules_step.argtypes = [_vp, C.POINTER(MgcIO), C.c_int, _vp]
    L.mgc_reset.argtypes = [_vp, C.POINTER(MgcIO), _vp]
    L.mgc_observe.argtypes = [_vp, C.POINTER(MgcIO), _vp]
    L.mgc_forecast_noise.argtypes = [_vp, _vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, _vp]
    L.mgc_launch_count.restype, L.mgc_launch_count.argtypes = C.c_int64, [_vp]
    L.mg_last_error.restype = C.c_char_p
    if L.mgc_abi_version() != MGC_ABI_VERSION:
        raise _cabi.EngineError(f"composed-step ABI mismatch: library {L.mgc_abi_version()} vs binding {MGC_ABI_VERSION}")
    for which, struct in enumerate((MgcModule, MgcLayout, MgcIO)):
        if L.mgc_sizeof(which) != C.sizeof(struct):
            raise _cabi.EngineError(f"struct {struct.__name__}: library sizeof {L.mgc_sizeof(which)} != binding {C.sizeof(struct)}")
    for name, kind in KIND.items():
        if L.mgc_param_count(kind) != PARAM_COUNT[name]:
            raise _cabi.EngineError(f"parameter block of {name}: library {L.mgc_param_count(kind)} != binding {PARAM_COUNT[name]}")
    L._mgc_bound = True
    return L


# ---- the static structure of a module list ------------------------------------------------------------------------------
class _Slot:
    """one module of a composition"""
    __slots__ = ("name", "index", "kind", "dispatch", "horizon", "raise_errors", "listing", "act_col", "n_act", "obs_off",
                 "obs_len", "param_off", "fstate_off", "istate_off", "is_source", "is_sink")


def _obs_len(kind, horizon):
    return {"load": 1 + horizon, "renewable": 1 + horizon, "grid": 4 * (1 + horizon), "battery": 2, "genset": 4}.get(kind, 0)


class Composition:
    def __init__(self, modules, add_unbalanced_module=True, loss_load_cost=10.0, overgeneration_cost=2.0, obs_order="gym_sorted"):
        """`modules`: list of pymgrid_b200.modules objects or (name, module) tuples, as Microgrid takes them
        (microgrid.py:100-165).  Returns the structure AND keeps the (name, module) records in listing order."""
        if isinstance(modules, (str, bytes)) or not hasattr(modules, "__iter__"):
            raise TypeError("modules must be list-like of modules.")
        import copy
        named = [(name, copy.copy(m)) for name, m in _named(list(modules))]      # microgrid.py:165 works on copies too
        for _, m in named:
            m.__dict__.pop("_runner", None)       # a standalone runner (module.step()) belongs to the caller's object only
        if add_unbalanced_module:       # appended un-named -> 'balancing' (microgrid.py:170-171)
            named.append(("balancing", UnbalancedEnergyModule(raise_errors=False, loss_load_cost=loss_load_cost,
                                                             overgeneration_cost=overgeneration_cost)))
        if len(named) > MGC_MAX_MODULES:
            raise NotImplementedError(f"at most {MGC_MAX_MODULES} modules per microgrid")
        # module_container.py:355-413: cells (fixed, flex, controllable) x (sources, sinks, source_and_sinks); names in
        # insertion order inside a cell; the modules of one name in insertion order
        cells = OrderedDict(((d, s), OrderedDict()) for d in ("fixed", "flex", "controllable")
                            for s in ("sources", "sinks", "source_and_sinks"))
        types = {}
        for name, m in named:
            kind, dispatch = m.module_type
            src, snk = kind != "load", kind in ("load", "battery", "grid", "balancing")
            cell = (dispatch, "source_and_sinks" if src and snk else "sources" if src else "sinks")
            if types.setdefault(name, cell) != cell:
                raise NameError(f"Attempted to add module {name} of type {cell}, but there is an identically named "
                                f"module of type {types[name]}.")
            cells[cell].setdefault(name, []).append(m)
        self.by_name = OrderedDict()            # name -> [module records], LISTING order (Container.to_dict)
        for cell in cells.values():
            self.by_name.update(cell)
        self.slots = []
        for name, lst in self.by_name.items():
            for j, m in enumerate(lst):
                s = _Slot()
                s.name, s.index, (s.kind, s.dispatch) = name, j, m.module_type
                s.horizon = int(getattr(m, "forecast_horizon", 0))
                s.raise_errors, s.listing = bool(m.raise_errors), len(self.slots)
                s.is_source, s.is_sink = s.kind != "load", s.kind in ("load", "battery", "grid", "balancing")
                s.n_act = {"genset": 2, "battery": 1, "grid": 1}.get(s.kind, 0)
                s.obs_len = _obs_len(s.kind, s.horizon)
                self.slots.append(s)
        self.records = [m for lst in self.by_name.values() for m in lst]       # listing order, parallel to self.slots
        # dispatch order of Microgrid.run: fixed, controllable, flex (microgrid.py:255-314)
        self.dispatch = [s for d in ("fixed", "controllable", "flex") for s in self.slots if s.dispatch == d]
        col = 0
        for s in self.dispatch:
            s.act_col = col if s.n_act else -1
            col += s.n_act
        self.n_act = col
        self.obs_order = obs_order
        if obs_order == "gym_sorted":       # gym.spaces.Dict sorts its keys (envs/base/base.py:128-163, 211-223)
            order = [s for name in sorted(self.by_name) for s in self.slots if s.name == name]
        elif obs_order == "container":
            order = list(self.slots)
        else:
            raise ValueError("obs_order must be 'gym_sorted' or 'container'")
        off = 0
        for s in order:
            s.obs_off = off
            off += s.obs_len
        self.obs_dim = off
        p, f, i = MGC_CFG_HEADER, 0, 0
        for s in self.slots:
            s.param_off = p
            p += PARAM_COUNT[s.kind]
            s.fstate_off, s.istate_off = (f if s.kind == "battery" else -1), (i if s.kind == "genset" else -1)
            f += 2 * (s.kind == "battery")
            i += 4 * (s.kind == "genset")
        self.cfg_stride, self.n_fstate, self.n_istate = p, f, i
        ts = [m for m in self.records if hasattr(m, "time_series")]
        if len({len(m) for m in ts}) > 1:
            raise ValueError("all time series must have the same length")
        self.series_len = len(ts[0]) if ts else 0
        initial = {m.initial_step for m in self.records}
        final = {(m.final_step if m.final_step > 0 else len(m)) for m in ts}      # base_timeseries_module.py:317-330
        if len(initial) > 1 or len(final) > 1:      # microgrid.py:640-675 reads a unique value
            raise ValueError("Attribute(s) ['initial_step' / 'final_step'] have non-unique values, cannot return single unique value.")
        self.initial_step = initial.pop() if initial else 0
        self.final_step = final.pop() if final else 0
        for m in ts:
            m._forecaster_params()          # None / 'oracle' / a noise standard deviation; Python callables are refused here

    def records_named(self):
        """[(name, record)] in listing order: what rebuilds the same microgrid (add_unbalanced_module=False)"""
        return [(s.name, r) for s, r in zip(self.slots, self.records)]

    @property
    def signature(self):
        """what B microgrids must share to be stepped as one batch"""
        return tuple((s.name, s.kind, s.horizon, s.raise_errors) for s in self.slots) + (self.obs_order, self.series_len)

    def controllable(self):
        """[(name, [slots])] in Microgrid.controllable's iteration order"""
        out = OrderedDict()
        for s in self.dispatch:
            if s.dispatch == "controllable":
                out.setdefault(s.name, []).append(s)
        return list(out.items())

    # ---- priority lists (algos/priority_list/priority_list.py:15-67) ----
    def priority_elements(self):
        """one element per action-space dimension of every controllable source (a genset: goal 0, goal 1), then one per
        controllable source-and-sink (battery, grid), in the container's order: [(slot, action)]"""
        ctl = [s for s in self.slots if s.dispatch == "controllable"]
        ordered = [s for s in ctl if not s.is_sink] + [s for s in ctl if s.is_sink]
        return [(s, a) for s in ordered for a in range(2 if s.kind == "genset" else 1)]

    def priority_lists(self, remove_redundant_gensets=False):
        """All deployment orders: every permutation of the elements, later elements of an already listed module dropped,
        duplicates removed in first-seen order; optionally without the lists that switch off a genset whose
        running_min_production is 0 (:53-67).  Tuples of (slot, action)."""
        from itertools import permutations
        elements = self.priority_elements()
        if len(elements) > MAX_PRIORITY_ELEMENTS:
            raise NotImplementedError(f"{len(elements)} priority-list elements: the action table is enumerated from "
                                      f"{len(elements)}! permutations like the reference's; at most {MAX_PRIORITY_ELEMENTS} are supported")
        seen, out = set(), []
        for perm in permutations(range(len(elements))):
            listed, pl = set(), []
            for i in perm:
                s = elements[i][0]
                if s.listing not in listed:
                    listed.add(s.listing)
                    pl.append(i)
            pl = tuple(pl)
            if pl not in seen:
                seen.add(pl)
                out.append(pl)
        if remove_redundant_gensets:
            redundant = {i for i, (s, a) in enumerate(elements)
                         if s.kind == "genset" and a == 0 and self.records[s.listing].running_min_production == 0}
            out = [pl for pl in out if not redundant.intersection(pl)]
        return [tuple(elements[i] for i in pl) for pl in out]

    def priority_table(self, lists):
        """int16 [n, width, 2] = (index in the DISPATCH-ordered module table, action); -1 pads"""
        width = max((len(pl) for pl in lists), default=1) or 1
        tab = np.full((max(len(lists), 1), width, 2), -1, dtype=np.int16)
        index = {id(s): k for k, s in enumerate(self.dispatch)}
        for r, pl in enumerate(lists):
            for c, (s, a) in enumerate(pl):
                tab[r, c] = index[id(s)], a
        return tab

    # ---- packing ----
    def config_record(self, series_index):
        """the f64 parameter record of this microgrid (include/pymgrid_b200_compose.h: enum MGC_* lists the blocks).
        `series_index(array, pull_zero) -> int` registers a series in the pool."""
        rec = np.zeros(self.cfg_stride)
        rec[0], rec[1] = self.initial_step, self.final_step
        for s, m in zip(self.slots, self.records):
            o = s.param_off
            if s.kind in ("load", "renewable"):     # bounds: base_timeseries_module.py:81-88
                lo, hi = views.series_bounds(m.time_series[:, 0], True)
                rec[o:o + 3] = series_index(m.time_series, True), lo, hi
            elif s.kind == "grid":                  # per-column bounds: grid_module.py:125-132
                ts = m.time_series
                rec[o:o + 4] = series_index(ts, False), m.max_import, m.max_export, m.cost_per_unit_co2
                rec[o + 4:o + 8], rec[o + 8:o + 12] = ts.min(axis=0), ts.max(axis=0)
            elif s.kind == "battery":
                rec[o:o + 6] = (m.min_capacity, m.max_capacity, m.max_charge, m.max_discharge, m.efficiency, m.battery_cost_cycle)
            elif s.kind == "genset":
                rec[o:o + 8] = (m.running_min_production, m.running_max_production, m.genset_cost, m.co2_per_unit,
                                m.cost_per_unit_co2, int(m.start_up_time), int(m.wind_down_time), int(bool(m.allow_abortion)))
            else:
                rec[o:o + 2] = m.loss_load_cost, m.overgeneration_cost
        return rec

    def initial_state(self):
        """(fstate row, istate row) of a freshly constructed microgrid: battery_module.py:89-106, genset_module.py:91-92"""
        f, i = np.zeros(self.n_fstate), np.zeros(self.n_istate, dtype=np.int32)
        for s, m in zip(self.slots, self.records):
            if s.kind == "battery":
                f[s.fstate_off:s.fstate_off + 2] = m.init_charge, m.init_soc
            elif s.kind == "genset":
                on = int(bool(m.init_start_up))
                i[s.istate_off:s.istate_off + 4] = (on, on, 0, int(m.wind_down_time)) if on else (0, 0, int(m.start_up_time), 0)
        return f, i

    _FIELDS = {"load": ["load"], "renewable": ["renewable"], "grid": ["import_price", "export_price", "co2_per_kwh", "grid_status"]}

    def field_names(self, s):
        """names of the elements of slot s's observation block = the keys of the module's state_dict()"""
        if s.kind in self._FIELDS:
            comps = self._FIELDS[s.kind]
            return [f"{c}_current" for c in comps] + [f"{c}_forecast_{j}" for j in range(s.horizon) for c in comps]
        return {"battery": ["soc", "current_charge"],
                "genset": ["current_status", "goal_status", "steps_until_up", "steps_until_down"]}.get(s.kind, [])

    def select_observation(self, keys):
        """BaseMicrogridEnv(observation_keys=...) (envs/base/base.py:109-123, 211-218): the observation is
        `state_series(normalized=True).loc[:, :, keys]` -- for every key in the order given, the modules that have such a
        field in listing order.  Returns [(slot, element)]; NameError for keys no module has."""
        if isinstance(keys, str):
            keys = [keys]
        names = {id(s): self.field_names(s) for s in self.slots}
        bad = [k for k in keys if not any(k in n for n in names.values())]
        if bad:
            raise NameError(f'Keys {bad} not found in state.')
        return [(s, names[id(s)].index(k)) for k in keys for s in self.slots if k in names[id(s)]]

    def noise_rows(self, elements):
        """GaussianNoiseForecaster parameters per element of the observation row (mgc_forecast_noise): the normalised standard
        deviation -- the module's noise_std, times |mean(time_series[initial_step:final_step])| under relative_noise
        (forecast/forecaster.py:237-250, base_timeseries_module.py:233-240), divided by the column's observation spread, 0 for
        constant columns (the forecaster's clip pins them), for the current value and for modules without noise -- and the
        increase_uncertainty flags.  `elements`: [(slot, element of its block)] in row order.  Returns None without noise."""
        per_slot = {}
        for s, r in zip(self.slots, self.records):
            f = r._forecaster_params() if hasattr(r, "time_series") else None
            if f is None or f.noise_std == 0:
                continue
            ts = r.time_series
            std = float(f.noise_std)
            if f.relative_noise:
                std *= float(np.abs(ts[self.initial_step:self.final_step].mean()))
            cols = []
            for c in range(ts.shape[1]):
                low, high = float(ts[:, c].min()), float(ts[:, c].max())
                if s.kind != "grid":
                    low, high = min(low, 0.0), max(high, 0.0)
                cols.append(std / (high - low) if high > low else 0.0)
            per_slot[id(s)] = (cols, float(bool(f.increase_uncertainty)))
        if not per_slot:
            return None
        sigma, inc = np.zeros(len(elements)), np.zeros(len(elements))
        for j, (s, k) in enumerate(elements):
            if id(s) in per_slot:
                cols, flag = per_slot[id(s)]
                C_ = len(cols)
                if k // C_ > 0:
                    sigma[j], inc[j] = cols[k % C_], flag
        return np.concatenate([sigma, inc])

    def row_elements(self):
        """[(slot, element)] of the full observation row, in row order"""
        order = sorted((s for s in self.slots if s.obs_len), key=lambda s: s.obs_off)
        return [(s, k) for s in order for k in range(s.obs_len)]

    def module_table(self):
        arr = (MgcModule * MGC_MAX_MODULES)()
        for k, s in enumerate(self.dispatch):
            arr[k] = MgcModule(KIND[s.kind], s.horizon, s.act_col, s.obs_off if s.obs_len else 0, s.param_off, s.fstate_off,
                               s.istate_off, s.listing, int(s.raise_errors))
        return arr


# ---- the batch ----------------------------------------------------------------------------------------------------------
def _open_device_library(device=None):
    """The CUDA extension bound for the composed entry points, and the CUDA device a batch lives on.  The only way in:
    there is no CPU execution path in this package -- a missing extension or a missing GPU raises EngineError."""
    L = bind(_cabi.lib())       # raises EngineError when the CUDA extension is missing
    if not torch.cuda.is_available():
        raise _cabi.EngineError("ComposedBatch needs a CUDA device: there is no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise _cabi.EngineError("ComposedBatch runs on CUDA devices only")
    return L, dev


class ComposedBatch:
    def __init__(self, microgrids, env_config=None, device=None, obs_order="gym_sorted", with_info=False,
                 microgrid_kwargs=None, prenormalised=True, observation_keys=()):
        """`microgrids`: list of module lists (one per parameter set; all with the same composition) or ready
        `Composition`s; `env_config[e]`: which one env e is (default: one env per entry)."""
        kw = dict(microgrid_kwargs or {})
        self.compositions = [m if isinstance(m, Composition) else Composition(m, obs_order=obs_order, **kw) for m in microgrids]
        if not self.compositions:
            raise ValueError("at least one microgrid is needed")
        self.comp = comp = self.compositions[0]
        for c in self.compositions[1:]:
            if c.signature != comp.signature:
                raise ValueError("the microgrids of one ComposedBatch must share one composition (module names, kinds, "
                                 "forecast horizons, series length); build one batch per composition")
        self._L, self.device = _open_device_library(device)
        env_config = np.arange(len(self.compositions)) if env_config is None else np.asarray(env_config, dtype=np.int64)
        if env_config.ndim != 1 or len(env_config) < 1 or env_config.min() < 0 or env_config.max() >= len(self.compositions):
            raise ValueError("env_config must index the microgrid list")
        self.env_config = env_config
        self.n_envs = n = len(env_config)
        # series pool, deduplicated by content
        pool, pool_nrm, offsets, seen, total = [], [], [], {}, 0

        def series_index(ts, pull_zero):
            """registers a series (deduplicated by content and bounds rule) and its normalised twin: (ts - low) / spread per
            column with the owning module kind's bounds, spread 0 -> 1 (utils/space.py:204-218) -- the same f64 operations
            the reference applies per step, done once"""
            nonlocal total
            arr = np.ascontiguousarray(ts, dtype=np.float64)
            key = (arr.shape, bool(pull_zero), arr.tobytes())
            if key not in seen:
                seen[key] = len(offsets)
                offsets.append(total)
                pool.append(arr.reshape(-1))
                if pull_zero:
                    lo, hi = views.series_bounds(arr[:, 0], True)
                    lo, hi = np.array([lo]), np.array([hi])
                else:
                    lo, hi = arr.min(axis=0), arr.max(axis=0)
                spread = hi - lo
                spread = np.where(spread == 0, 1.0, spread)
                pool_nrm.append(((arr - lo) / spread).reshape(-1))
                total += arr.size
            return seen[key]
        cfg = np.stack([c.config_record(series_index) for c in self.compositions])
        states = [c.initial_state() for c in self.compositions]
        dev = self.device
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dtype=dt).to(dev)      # noqa: E731
        self.cfg = t(cfg, torch.float64)
        self.series = t(np.concatenate(pool) if pool else np.zeros(1), torch.float64)
        self.series_nrm = t(np.concatenate(pool_nrm) if pool_nrm else np.zeros(1), torch.float64)
        self.series_off = t(np.array(offsets if offsets else [0], dtype=np.int64), torch.int64)
        self.cfg_index = t(env_config, torch.int32)
        self.step_counter = t(np.array([self.compositions[c].initial_step for c in env_config]), torch.int32)
        # per-env episode windows (trajectory functions rewrite them; the kernel reads them at every step / reset)
        self.env_initial_step = self.step_counter.clone()
        self.env_final_step = t(np.array([self.compositions[c].final_step for c in env_config]), torch.int32)
        self.fstate = t(np.stack([states[c][0] for c in env_config]).reshape(n, comp.n_fstate), torch.float64)
        self.istate = t(np.stack([states[c][1] for c in env_config]).reshape(n, comp.n_istate), torch.int32)
        # observation_keys: the kernel writes only the selected elements, in the order the reference's envs return them
        self.selection = comp.select_observation(observation_keys) if observation_keys else None
        self.obs_dim = len(self.selection) if self.selection is not None else comp.obs_dim
        self.obs = torch.zeros((n, self.obs_dim), dtype=torch.float64, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float64, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.flags = torch.zeros(n, dtype=torch.int32, device=dev)
        self.n_info = len(comp.slots) * MGC_INFO_SLOTS + MGC_BALANCE_SLOTS
        self.info = torch.zeros((n, self.n_info), dtype=torch.float64, device=dev) if with_info else None
        L = MgcLayout()
        L.abi_version, L.n_modules, L.modules = MGC_ABI_VERSION, len(comp.slots), comp.module_table()
        L.n_act, L.obs_dim, L.n_fstate, L.n_istate = comp.n_act, self.obs_dim, comp.n_fstate, comp.n_istate
        if self.selection is not None:
            index = {id(s): k for k, s in enumerate(comp.dispatch)}
            select = (C.c_int32 * max(self.obs_dim, 1))(*[(index[id(s)] << 16) | k for s, k in self.selection])
            L.obs_select = C.cast(select, _vp)
        L.cfg_stride, L.n_cfg, L.series_len, L.n_series, L.n_envs = comp.cfg_stride, len(cfg), comp.series_len, len(offsets), n
        L.cfg, L.series, L.series_off = self.cfg.data_ptr(), self.series.data_ptr(), self.series_off.data_ptr()
        L.series_nrm = self.series_nrm.data_ptr() if prenormalised else None      # None: the kernel normalises every element itself
        L.step, L.cfg_index = self.step_counter.data_ptr(), self.cfg_index.data_ptr()
        L.fstate = self.fstate.data_ptr() if comp.n_fstate else None
        L.istate = self.istate.data_ptr() if comp.n_istate else None
        # the discrete action table: every priority list of the composition (redundant genset lists included; callers
        # that drop them map their own indices, see ComposedDiscreteEnv)
        self.action_lists = None
        if comp.n_act and len(comp.priority_elements()) <= MAX_PRIORITY_ELEMENTS:
            self.action_lists = comp.priority_lists(False)
            self.plist = t(comp.priority_table(self.action_lists), torch.int16)
            L.plist, L.n_plist, L.plist_width = self.plist.data_ptr(), len(self.action_lists), self.plist.shape[1]
        L.env_initial_step, L.env_final_step = self.env_initial_step.data_ptr(), self.env_final_step.data_ptr()
        # Gaussian-noise forecasters of the modules (forecast/forecaster.py:220-262): on by default when any module asks for one
        elements = self.selection if self.selection is not None else comp.row_elements()
        rows = [c.noise_rows([(c.slots[s.listing], k) for s, k in elements]) for c in self.compositions]
        self._noise = None
        self._noise_calls = 0
        if any(r is not None for r in rows):
            rows = [r if r is not None else np.zeros(2 * self.obs_dim) for r in rows]
            self._noise_rows = t(np.stack(rows), torch.float64)
            self.set_forecast_noise(seed=0)
        self._handle = _vp()
        with self._on_device():         # mgc_create uploads its element table to the CURRENT device
            self._check(self._L.mgc_create(C.byref(L), C.byref(self._handle)), "mgc_create")

    def set_trajectories(self, initial_step, final_step):
        """per-env episode windows [B] (microgrid/trajectory/*.py through `trajectory_func`, microgrid.py:221-225): `reset`
        puts an env at its initial step, `done` is reported at its final_step - 1.  Must stay inside the configured window."""
        ini = torch.as_tensor(np.asarray(initial_step), dtype=torch.int32, device=self.device).reshape(self.n_envs)
        fin = torch.as_tensor(np.asarray(final_step), dtype=torch.int32, device=self.device).reshape(self.n_envs)
        lo = torch.as_tensor(np.array([self.compositions[c].initial_step for c in self.env_config]), dtype=torch.int32, device=self.device)
        hi = torch.as_tensor(np.array([self.compositions[c].final_step for c in self.env_config]), dtype=torch.int32, device=self.device)
        if bool((ini < lo).any()) or bool((fin > hi).any()) or bool((ini >= fin).any()):
            raise ValueError("trajectory windows must lie inside [initial_step, final_step] and be non-empty")     # microgrid.py:184-197
        self.env_initial_step.copy_(ini)
        self.env_final_step.copy_(fin)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                self._L.mgc_destroy(h)
            except Exception:       # interpreter shutdown
                pass

    def _check(self, code, what):
        if code != 0:
            raise _cabi.EngineError(f"{what} failed ({code}): {self._L.mg_last_error().decode()}")

    def set_forecast_noise(self, seed=0, env_offset=0):
        """(Re)seed the Gaussian-noise forecasts: after every step / reset / observe the forecast entries of the freshly written
        rows get N(0, std_k) added and are clipped to their bounds (mgc_forecast_noise).  The draw is a pure function of
        (seed, call number, env_offset + env, the env's step, element); shards of one batch pass their own env_offset.
        `rollout` keeps the oracle forecast inside its launch."""
        if getattr(self, "_noise_rows", None) is None:
            raise ValueError("no module of this batch has a Gaussian-noise forecaster")
        self._noise = (int(seed) & (2 ** 64 - 1), int(env_offset))

    def clear_forecast_noise(self):
        self._noise = None

    def _apply_noise(self, obs):
        if self._noise is None or obs is None:
            return
        seed, base = self._noise
        self._noise_calls += 1
        with self._on_device():
            self._check(self._L.mgc_forecast_noise(self._handle, self._noise_rows.data_ptr(), obs.data_ptr(), base, seed,
                                                   self._noise_calls, self._stream()), "mgc_forecast_noise")

    def _on_device(self):
        import contextlib
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    @property
    def launch_count(self):
        return int(self._L.mgc_launch_count(self._handle))

    def _actions(self, actions, lead):
        comp = self.comp
        if comp.n_act == 0:
            return None
        a = torch.as_tensor(actions, dtype=torch.float64, device=self.device).contiguous()
        if tuple(a.shape) != lead + (self.n_envs, comp.n_act):
            raise ValueError(f"actions must have shape {lead + (self.n_envs, comp.n_act)}, got {tuple(a.shape)}")
        return a

    def step(self, actions=None, normalized=True, obs=True):
        """Microgrid.run for every env (microgrid.py:227-325): `actions` [B, n_act] in the composition's control order
        (`Composition.controllable()`; genset: goal, energy).  Returns (obs [B, D] | None, reward [B], done [B], info | None);
        the tensors are the batch's own buffers, overwritten by the next call."""
        a = self._actions(actions, ())
        io = MgcIO(a.data_ptr() if a is not None else None, self.obs.data_ptr() if obs else None, self.reward.data_ptr(),
                   self.done.data_ptr(), self.info.data_ptr() if self.info is not None else None, self.flags.data_ptr(), None)
        with self._on_device():
            self._check(self._L.mgc_run(self._handle, C.byref(io), 1, 1, int(bool(normalized)), self._stream()), "mgc_run")
        self._apply_noise(self.obs if obs else None)
        return (self.obs if obs else None), self.reward, self.done, self.info

    def modules_step(self, actions=None, normalized=True, obs=True):
        """BaseMicrogridModule.step for every module on its own (mgc_modules_step): `actions` [B, W], one column per module in
        dispatch order (load: none; renewable, battery, grid, slack: one; genset: goal, energy).  No energy balance."""
        W = sum(2 if s.kind == "genset" else 0 if s.kind == "load" else 1 for s in self.comp.slots)
        a = None
        if W:
            a = torch.as_tensor(actions, dtype=torch.float64, device=self.device).contiguous()
            if tuple(a.shape) != (self.n_envs, W):
                raise ValueError(f"actions must have shape {(self.n_envs, W)}, got {tuple(a.shape)}")
        io = MgcIO(a.data_ptr() if a is not None else None, self.obs.data_ptr() if obs else None, self.reward.data_ptr(),
                   self.done.data_ptr(), self.info.data_ptr() if self.info is not None else None, self.flags.data_ptr(), None)
        with self._on_device():
            self._check(self._L.mgc_modules_step(self._handle, C.byref(io), int(bool(normalized)), self._stream()), "mgc_modules_step")
        self._apply_noise(self.obs if obs else None)
        return (self.obs if obs else None), self.reward, self.done, self.info

    def state_dict(self):
        """the reference's serialisable state (base_module.py:852-868, genset_module.py:426-427) of every env: step, battery
        charge + soc, genset integers, episode windows -- a checkpoint is these five tensors"""
        return {a: getattr(self, a).clone() for a in ("step_counter", "fstate", "istate", "env_initial_step", "env_final_step")}

    def load_state_dict(self, state):
        for a, v in state.items():
            getattr(self, a).copy_(v)

    def recorder(self, env_ids):
        """opt-in reference-format log for a subset of the envs (see ComposedLogRecorder)"""
        return ComposedLogRecorder(self, env_ids)

    def _dactions(self, actions, lead):
        if self.action_lists is None:
            raise NotImplementedError("this composition has no discrete action table (no controllable module, or more than "
                                      f"{MAX_PRIORITY_ELEMENTS} priority-list elements)")
        a = torch.as_tensor(actions, dtype=torch.int32, device=self.device).contiguous()
        if tuple(a.shape) != lead + (self.n_envs,):
            raise ValueError(f"discrete actions must have shape {lead + (self.n_envs,)}, got {tuple(a.shape)}")
        return a

    def step_discrete(self, actions, obs=True):
        """DiscreteMicrogridEnv.step for every env (envs/discrete/discrete.py:109-143): `actions` int32 [B], indices into
        `action_lists`; the expansion into controls (priority_list.py:69-116) runs on the device in front of the step."""
        a = self._dactions(actions, ())
        io = MgcIO(None, self.obs.data_ptr() if obs else None, self.reward.data_ptr(), self.done.data_ptr(),
                   self.info.data_ptr() if self.info is not None else None, self.flags.data_ptr(), None, a.data_ptr(), 0)
        with self._on_device():
            self._check(self._L.mgc_run_discrete(self._handle, C.byref(io), 1, 1, self._stream()), "mgc_run_discrete")
        self._apply_noise(self.obs if obs else None)
        return (self.obs if obs else None), self.reward, self.done, self.info

    def rollout_discrete(self, actions, n_steps=None, ring=1, obs=True):
        """`actions` int32 [T, B]: T discrete steps in one launch; or int32 [B] with `n_steps`: the same priority list
        every step -- rule-based control (algos/rbc/rbc.py:64-93)."""
        const = n_steps is not None
        T = int(n_steps) if const else int(torch.as_tensor(actions).shape[0])
        a = self._dactions(actions, () if const else (T,))
        reward = torch.empty((T, self.n_envs), dtype=torch.float64, device=self.device)
        done = torch.empty((T, self.n_envs), dtype=torch.uint8, device=self.device)
        ring_buf = torch.zeros((ring, self.n_envs, self.obs_dim), dtype=torch.float64, device=self.device) if obs else None
        io = MgcIO(None, ring_buf.data_ptr() if obs else None, reward.data_ptr(), done.data_ptr(),
                   self.info.data_ptr() if self.info is not None else None, self.flags.data_ptr(), None, a.data_ptr(), int(const))
        with self._on_device():
            self._check(self._L.mgc_run_discrete(self._handle, C.byref(io), T, int(ring), self._stream()), "mgc_run_discrete")
        return dict(reward=reward, done=done, obs_ring=ring_buf, flags=self.flags)

    def rollout(self, actions=None, n_steps=None, normalized=True, ring=1, obs=True, out=None):
        """`n_steps` consecutive steps in ONE launch; `actions` [T, B, n_act].  Returns dict(reward [T, B], done [T, B],
        obs_ring [ring, B, D] | None, flags [B] OR-ed over the steps); `out`: a dict returned by an earlier call of the same
        shape, to write into instead of allocating."""
        comp = self.comp
        if comp.n_act:
            T = int(torch.as_tensor(actions).shape[0]) if n_steps is None else int(n_steps)
            a = self._actions(actions, (T,))
        else:
            if n_steps is None:
                raise ValueError("n_steps is required when the composition has no controllable module")
            T, a = int(n_steps), None
        if out is not None and tuple(out["reward"].shape) == (T, self.n_envs) and \
                (not obs or (out["obs_ring"] is not None and out["obs_ring"].shape[0] == ring)):
            reward, done, ring_buf = out["reward"], out["done"], (out["obs_ring"] if obs else None)
        else:
            reward = torch.empty((T, self.n_envs), dtype=torch.float64, device=self.device)
            done = torch.empty((T, self.n_envs), dtype=torch.uint8, device=self.device)
            ring_buf = torch.zeros((ring, self.n_envs, self.obs_dim), dtype=torch.float64, device=self.device) if obs else None
        io = MgcIO(a.data_ptr() if a is not None else None, ring_buf.data_ptr() if obs else None, reward.data_ptr(),
                   done.data_ptr(), self.info.data_ptr() if self.info is not None else None, self.flags.data_ptr(), None)
        with self._on_device():
            self._check(self._L.mgc_run(self._handle, C.byref(io), T, int(ring), int(bool(normalized)), self._stream()), "mgc_run")
        return dict(reward=reward, done=done, obs_ring=ring_buf, flags=self.flags)

    def host_rollout(self, actions, reward, done, chunk=32, normalized=True, ring=1, obs=True):
        """`rollout` with HOST buffers: `actions` [T, B, n_act] f64, `reward` [T, B] f64 and `done` [T, B] uint8 are pinned
        host tensors.  The steps run in launches of `chunk` steps; the host -> device copy of the next chunk's actions, the
        kernel of the current chunk and the device -> host copy of the previous chunk's rewards overlap on three streams
        (two device buffers each way).  Returns dict(obs_ring, flags): the ring as the LAST launch left it (slot s % ring of
        that launch) and the event bits OR-ed over all steps.  The caller's stream is synchronised with the copies on
        return; the host buffers are complete after `torch.cuda.current_stream().synchronize()`."""
        comp = self.comp
        T = int(reward.shape[0])
        if tuple(reward.shape) != (T, self.n_envs) or tuple(done.shape) != (T, self.n_envs):
            raise ValueError(f"reward / done must have shape ({T}, {self.n_envs})")
        if comp.n_act and tuple(actions.shape) != (T, self.n_envs, comp.n_act):
            raise ValueError(f"actions must have shape ({T}, {self.n_envs}, {comp.n_act})")
        if reward.dtype != torch.float64 or done.dtype != torch.uint8 or (comp.n_act and actions.dtype != torch.float64):
            raise ValueError("host_rollout: actions / reward are float64, done is uint8")
        chunk = max(1, min(int(chunk), T))
        key = (chunk, int(ring), bool(obs))
        st = getattr(self, "_host_rollout_state", None)
        if st is None or st["key"] != key:
            with self._on_device():
                dev = self.device
                st = dict(key=key, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev),
                          act=[torch.empty((chunk, self.n_envs, comp.n_act), dtype=torch.float64, device=dev) for _ in range(2)],
                          rew=[torch.empty((chunk, self.n_envs), dtype=torch.float64, device=dev) for _ in range(2)],
                          done=[torch.empty((chunk, self.n_envs), dtype=torch.uint8, device=dev) for _ in range(2)],
                          ring=torch.zeros((ring, self.n_envs, self.obs_dim), dtype=torch.float64, device=dev) if obs else None,
                          flags=torch.zeros_like(self.flags),
                          in_ready=[torch.cuda.Event() for _ in range(2)], in_free=[torch.cuda.Event() for _ in range(2)],
                          out_ready=[torch.cuda.Event() for _ in range(2)], out_free=[torch.cuda.Event() for _ in range(2)])
            self._host_rollout_state = st
        with self._on_device():
            cur = torch.cuda.current_stream(self.device)
            st["s_in"].wait_stream(cur)
            st["s_out"].wait_stream(cur)
            st["flags"].zero_()
            for c, k0 in enumerate(range(0, T, chunk)):
                n, b = min(chunk, T - k0), c & 1
                if comp.n_act:
                    with torch.cuda.stream(st["s_in"]):
                        if c >= 2:
                            st["s_in"].wait_event(st["in_free"][b])         # the launch that read this buffer has finished
                        st["act"][b][:n].copy_(actions[k0:k0 + n], non_blocking=True)
                        st["in_ready"][b].record(st["s_in"])
                    cur.wait_event(st["in_ready"][b])
                if c >= 2:
                    cur.wait_event(st["out_free"][b])                       # its rewards have left the device
                out = dict(reward=st["rew"][b][:n], done=st["done"][b][:n], obs_ring=st["ring"])
                self.rollout(st["act"][b][:n] if comp.n_act else None, n_steps=n, normalized=normalized, ring=ring, obs=obs, out=out)
                st["flags"] |= self.flags
                st["in_free"][b].record(cur)
                st["out_ready"][b].record(cur)
                with torch.cuda.stream(st["s_out"]):
                    st["s_out"].wait_event(st["out_ready"][b])
                    reward[k0:k0 + n].copy_(st["rew"][b][:n], non_blocking=True)
                    done[k0:k0 + n].copy_(st["done"][b][:n], non_blocking=True)
                    st["out_free"][b].record(st["s_out"])
            cur.wait_stream(st["s_out"])
        return dict(obs_ring=st["ring"], flags=st["flags"])

    def reset(self, mask=None):
        """Microgrid.reset (microgrid.py:205-225) for the masked envs (default all); returns every env's observation"""
        m = None if mask is None else torch.as_tensor(mask, dtype=torch.uint8, device=self.device).contiguous()
        io = MgcIO(None, self.obs.data_ptr(), None, None, None, None, m.data_ptr() if m is not None else None)
        with self._on_device():
            self._check(self._L.mgc_reset(self._handle, C.byref(io), self._stream()), "mgc_reset")
        self._apply_noise(self.obs)
        return self.obs

    def observe(self):
        io = MgcIO(None, self.obs.data_ptr(), None, None, None, None, None)
        with self._on_device():
            self._check(self._L.mgc_observe(self._handle, C.byref(io), self._stream()), "mgc_observe")
        self._apply_noise(self.obs)
        return self.obs


def _raise_for(flags):
    """the reference's exceptions for the event bits of one env (raised after the step has been applied)"""
    if flags & FLAG_NOT_A_SINK:     # a source-only module asked to absorb: as_sink compares with max_consumption, which such a
        # module does not implement (base_module.py:265, :604-619)
        raise TypeError("'>' not supported between instances of 'float' and 'NotImplementedType'")
    if flags & (FLAG_GENSET_GOAL_RANGE | FLAG_BATTERY_MIN_CAP | FLAG_NEGATIVE_ABSORB):
        raise AssertionError(f"step rejected (flags {flags:#x})")
    if flags & FLAG_CLIP_RAISES:
        raise ValueError("requested value outside the module's limits")                            # base_module.py:79-93


# ---- B = 1: the reference's Microgrid surface ---------------------------------------------------------------------------
class ModuleList(list):
    """microgrid.modules.<name>: list of module views (module_container.py ModuleList)"""

    def item(self):
        if len(self) != 1:
            raise ValueError("Can only convert a ModuleList of length one to a scalar")
        return self[0]

    def to_list(self):
        return self


_STATE_NAMES = {"load": ["load"], "renewable": ["renewable"], "grid": ["import_price", "export_price", "co2_per_kwh", "grid_status"]}
_ENERGY_NAMES = {"load": (None, "load_met"), "renewable": ("renewable_used", None), "battery": ("discharge_amount", "charge_amount"),
                 "genset": ("genset_production", None), "grid": ("grid_import", "grid_export"), "balancing": ("loss_load", "overgeneration")}


def slot_state_dict(s, r, t, f, i):
    """unnormalised state of module slot `s` (record `r`) at step t: `f` = its two battery doubles, `i` = its four genset
    integers (None for other kinds) -- the modules' _state_dict of the reference"""
    d = OrderedDict()
    if s.kind in _STATE_NAMES:
        ts = r.time_series
        if s.kind == "grid":
            lo, hi = ts.min(axis=0), ts.max(axis=0)
        else:
            lo, hi = views.series_bounds(ts[:, 0], True)
            lo, hi = np.array([lo]), np.array([hi])
        vals = views.series_state(ts, t, s.horizon, lo, hi)
        comps = _STATE_NAMES[s.kind]
        keys = [f"{c}_current" for c in comps] + [f"{c}_forecast_{j}" for j in range(s.horizon) for c in comps]
        d.update(zip(keys, (float(v) for v in vals)))
    elif s.kind == "battery":
        d.update(soc=float(f[1]), current_charge=float(f[0]))
    elif s.kind == "genset":
        d.update(current_status=int(i[0]), goal_status=int(i[1]), steps_until_up=int(i[2]), steps_until_down=int(i[3]))
    return d


def log_row_from(comp, records, t, fstate, istate_pre, istate_post, info, reward, shaped=None):
    """One row of Microgrid.get_log() (base_module.py:276-290 per module, microgrid.py:259-260, 281, 317-319 for the
    balance) from an env's raw state rows BEFORE the step, its genset integers AFTER it (the genset logs its state after the
    status update, genset_module.py:148-149) and its info row."""
    row = OrderedDict()
    for s, r in zip(comp.slots, records):
        f = fstate[s.fstate_off:s.fstate_off + 2] if s.kind == "battery" else None
        i = istate_post[s.istate_off:s.istate_off + 4] if s.kind == "genset" else None
        x = info[s.listing * MGC_INFO_SLOTS:(s.listing + 1) * MGC_INFO_SLOTS]
        key = (s.name, s.index)
        row[key + ("reward",)] = float(x[3])
        if s.kind in ("genset", "grid"):
            row[key + ("co2_production",)] = float(x[2])
        elif s.kind == "renewable":
            row[key + ("curtailment",)] = float(x[2])
        p_name, a_name = _ENERGY_NAMES[s.kind]
        if p_name is not None:
            row[key + (p_name,)] = float(x[0])
        if a_name is not None:
            row[key + (a_name,)] = float(x[1])
        for k, val in slot_state_dict(s, r, t, f, i).items():
            row[key + (k,)] = val
    bal = info[len(comp.slots) * MGC_INFO_SLOTS:]
    for k, val in (("reward", reward), ("shaped_reward", reward if shaped is None else shaped), ("overall_provided_to_microgrid", bal[4]),
                   ("overall_absorbed_from_microgrid", bal[5]), ("controllable_provided_to_microgrid", bal[2]),
                   ("controllable_absorbed_from_microgrid", bal[3]), ("fixed_provided_to_microgrid", bal[0]),
                   ("fixed_absorbed_from_microgrid", bal[1])):
        row[("balance", 0, k)] = float(val)
    return row


class ComposedLogRecorder:
    """Opt-in reference-format log for a SUBSET of a composed batch (the full log of 65 536 envs would be hundreds of GB per
    year): step the batch through the recorder and read `get_log(env)` -- the DataFrame Microgrid.get_log() returns
    (microgrid.py:434-475) for each recorded env.  Host-side, from the info block and small state reads of the recorded envs."""

    def __init__(self, batch, env_ids):
        if batch.info is None:
            raise ValueError("the recorder needs ComposedBatch(with_info=True)")
        self.batch, self.env_ids = batch, [int(e) for e in env_ids]
        self._idx = torch.as_tensor(self.env_ids, dtype=torch.int64, device=batch.device)
        self.rows = {e: [] for e in self.env_ids}

    def _snap(self):
        b = self.batch
        return (b.step_counter.index_select(0, self._idx).cpu().numpy(), b.fstate.index_select(0, self._idx).cpu().numpy(),
                b.istate.index_select(0, self._idx).cpu().numpy())

    def _record(self, pre):
        b = self.batch
        t, f, _ = pre
        _, _, i_post = self._snap()
        info = b.info.index_select(0, self._idx).cpu().numpy()
        reward = b.reward.index_select(0, self._idx).cpu().numpy()
        flags = b.flags.index_select(0, self._idx).cpu().numpy()
        for k, e in enumerate(self.env_ids):
            if int(flags[k]) & (FLAG_STEP_PAST_END | FLAG_BAD_ACTION):
                continue            # the reference raises there and logs nothing
            comp = b.compositions[int(b.env_config[e])]
            self.rows[e].append(log_row_from(comp, comp.records, int(t[k]), f[k], None, i_post[k], info[k], float(reward[k])))

    def step(self, actions=None, normalized=True, obs=True):
        pre = self._snap()
        out = self.batch.step(actions, normalized=normalized, obs=obs)
        self._record(pre)
        return out

    def step_discrete(self, actions, obs=True):
        pre = self._snap()
        out = self.batch.step_discrete(actions, obs=obs)
        self._record(pre)
        return out

    def reset(self, mask=None):
        out = self.batch.reset(mask)
        m = None if mask is None else np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask, dtype=bool)
        for e in self.env_ids:
            if m is None or m[e]:
                self.rows[e] = []
        return out

    def get_log(self, env, drop_singleton_key=False):
        stop = int(self.batch.step_counter[int(env)].item())
        return views.log_frame(self.rows[int(env)], stop, drop_singleton_key)


class ComposedModuleView:
    """Read-only view of one module of a ComposedMicrogrid: constructor parameters + the live values the reference's
    callers read."""
    _STATE_NAMES, _ENERGY_NAMES = _STATE_NAMES, _ENERGY_NAMES

    def __init__(self, microgrid, slot, record):
        self._m, self._s, self._r = microgrid, slot, record
        self.name = (slot.name, slot.index)

    def __repr__(self):
        return f"{type(self._r).__name__}View{self.name}"

    def __getattr__(self, item):        # constructor parameters: battery.max_capacity, genset.genset_cost, load.time_series ...
        if item.startswith("_"):
            raise AttributeError(item)
        return getattr(self._r, item)

    _module_record = property(lambda self: self._r)      # what `Microgrid(microgrid.modules.to_tuples())` rebuilds from
    module_type = property(lambda self: self._r.module_type)
    is_source = property(lambda self: self._s.is_source)
    is_sink = property(lambda self: self._s.is_sink)
    current_step = property(lambda self: self._m.current_step)
    # a module's own window: what trajectory_func set at the last reset (microgrid.py:221-225), else the microgrid's
    initial_step = property(lambda self: int(self._m._batch.env_initial_step[0].item()))
    final_step = property(lambda self: int(self._m._batch.env_final_step[0].item()))
    provided_energy_name = property(lambda self: self._ENERGY_NAMES[self._s.kind][0])
    absorbed_energy_name = property(lambda self: self._ENERGY_NAMES[self._s.kind][1])

    @property
    def action_space(self):
        from types import SimpleNamespace
        n = self._s.n_act if self._s.dispatch == "controllable" else (1 if self._s.dispatch == "flex" else 0)
        return SimpleNamespace(shape=(n,))

    def _state(self):
        t, f, i = self._m._state()
        s = self._s
        return t, (f[s.fstate_off:s.fstate_off + 2] if s.kind == "battery" else None), \
            (i[s.istate_off:s.istate_off + 4] if s.kind == "genset" else None)

    def _row(self):
        return self._r.time_series[self._m.current_step]

    # live values
    current_load = property(lambda self: -1 * self._row().item())                  # load_module.py:100-111
    current_renewable = property(lambda self: self._row().item())                  # renewable_module.py:99-110
    current_charge = property(lambda self: float(self._state()[1][0]))
    soc = property(lambda self: float(self._state()[1][1]))
    current_status = property(lambda self: int(self._state()[2][0]) if self._s.kind == "genset" else self._row()[3])
    goal_status = property(lambda self: int(self._state()[2][1]))

    @property
    def max_production(self):
        k, r = self._s.kind, self._r
        if k == "battery":
            return min(r.max_discharge, self.current_charge - r.min_capacity) * r.efficiency
        if k == "genset":
            return self.current_status * r.running_max_production
        if k == "grid":
            return r.max_import * self._row()[3]
        if k == "renewable":
            return self.current_renewable
        if k == "balancing":
            return np.inf
        return 0.0

    @property
    def min_production(self):
        return self.current_status * self._r.running_min_production if self._s.kind == "genset" else 0

    @property
    def max_consumption(self):
        k, r = self._s.kind, self._r
        if k == "battery":
            return min(r.max_charge, r.max_capacity - self.current_charge) / r.efficiency
        if k == "grid":
            return r.max_export * self._row()[3]
        if k == "load":
            return self.current_load
        if k == "balancing":
            return np.inf
        return 0.0

    # genset look-ahead (genset_module.py:360-424): the status / limits one step after asking for `goal_status`
    def next_status(self, goal_status):
        cs, _, up, dn = (int(x) for x in self._state()[2])
        if goal_status:
            return 1 if (cs or up == 0) else 0
        return 0 if (not cs or dn == 0) else 1

    def next_max_production(self, goal_status):
        return self.next_status(goal_status) * self._r.running_max_production

    def next_min_production(self, goal_status):
        return self.next_status(goal_status) * self._r.running_min_production

    @property
    def production_marginal_cost(self):
        k, r = self._s.kind, self._r
        return {"battery": lambda: r.battery_cost_cycle,
                "genset": lambda: r.genset_cost * 1.0 + r.cost_per_unit_co2 * (r.co2_per_unit * 1.0),
                "grid": lambda: float(self.state[0]), "balancing": lambda: r.loss_load_cost}.get(k, lambda: 0.0)()   # grid_module.py:322-324

    @property
    def absorption_marginal_cost(self):
        k, r = self._s.kind, self._r
        return {"battery": lambda: r.battery_cost_cycle, "grid": lambda: float(self.state[1]),
                "balancing": lambda: r.overgeneration_cost}.get(k, lambda: 0.0)()

    marginal_cost = production_marginal_cost

    def _series_bounds(self):
        ts = self._r.time_series
        if self._s.kind == "grid":
            return ts.min(axis=0), ts.max(axis=0)
        lo, hi = views.series_bounds(ts[:, 0], True)
        return np.array([lo]), np.array([hi])

    def state_dict(self, normalized=False):
        """base_module.py:473-490 / the modules' _state_dict"""
        t, f, i = self._state()
        d = slot_state_dict(self._s, self._r, t, f, i)
        if normalized and d:
            lo, hi = self.min_obs, self.max_obs
            spread = np.where(hi - lo == 0, 1.0, hi - lo)
            d = OrderedDict(zip(d.keys(), (np.array(list(d.values()), dtype=np.float64) - lo) / spread))
        return d

    @property
    def state(self):
        return np.array(list(self.state_dict().values()), dtype=np.float64)

    def _obs_bounds(self):
        s, r = self._s, self._r
        if s.kind in self._STATE_NAMES:
            lo, hi = self._series_bounds()
            return np.tile(lo, 1 + s.horizon), np.tile(hi, 1 + s.horizon)
        if s.kind == "battery":
            return np.array([r.min_capacity / r.max_capacity, r.min_capacity]), np.array([1.0, r.max_capacity])
        if s.kind == "genset":
            return np.zeros(4), np.array([1.0, 1.0, r.start_up_time, r.wind_down_time], dtype=np.float64)
        return np.array([]), np.array([])

    min_obs = property(lambda self: self._obs_bounds()[0])
    max_obs = property(lambda self: self._obs_bounds()[1])

    def _act_bounds(self):
        k, r = self._s.kind, self._r
        if k == "battery":
            return -r.max_discharge / r.efficiency, r.max_charge * r.efficiency
        if k == "genset":
            return np.array([0.0, 0.0]), np.array([1.0, r.running_max_production])
        if k == "grid":
            return -1 * r.max_export, r.max_import
        if k == "renewable":
            lo, hi = self._series_bounds()
            return lo[0], hi[0]
        if k == "balancing":
            return -np.inf, np.inf
        return np.array([]), np.array([])

    min_act = property(lambda self: self._act_bounds()[0])
    max_act = property(lambda self: self._act_bounds()[1])

    def _affine(self, act, obs):
        assert act + obs == 1, "One of act or obs must be True but not both."
        low, high = (self._act_bounds() if act else self._obs_bounds())
        low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
        spread = high - low
        return low, np.where(spread == 0, 1.0, spread)

    def to_normalized(self, value, act=False, obs=False):
        """reference: BaseMicrogridModule.to_normalized -> ModuleSpace.normalize (utils/space.py:207-218)"""
        low, spread = self._affine(act, obs)
        return (np.asarray(value, dtype=np.float64) - low) / spread

    def from_normalized(self, value, act=False, obs=False):
        """reference: ModuleSpace.denormalize (utils/space.py:220-231)"""
        low, spread = self._affine(act, obs)
        return low + spread * np.asarray(value, dtype=np.float64)

    # grid columns of the current state, current value first then the forecast (grid_module.py:248-299: state[k::4])
    import_price = property(lambda self: self.state[0::4])
    export_price = property(lambda self: self.state[1::4])
    co2_per_kwh = property(lambda self: self.state[2::4])
    grid_status = property(lambda self: self.state[3::4])

    def sample_action(self, strict_bound=False):
        """base_module.py:326-356 (genset: genset_module.py:348-349)"""
        if self._s.kind == "genset":
            return np.array([np.random.rand(), np.random.rand()])
        lo_b, hi_b = 0, 1
        if strict_bound:
            lo, hi = self._act_bounds()
            spread = (hi - lo) or 1.0
            if self.is_sink:
                lo_b = (-1 * self.max_consumption - lo) / spread
                lo_b = 0 if np.isnan(lo_b) else lo_b
            if self.is_source:
                hi_b = (self.max_production - lo) / spread
                hi_b = 0 if np.isnan(hi_b) else hi_b
        return np.random.rand() * (hi_b - lo_b) + lo_b


class ComposedContainer(OrderedDict):
    """`microgrid.modules` and its sub-containers: name -> ModuleList, attribute access, the reference's helpers"""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item)

    def iterdict(self):
        return self.items()

    def to_dict(self):
        return OrderedDict(self)

    def iterlist(self):
        return [m for lst in self.values() for m in lst]

    to_list = iterlist

    def names(self):
        return list(self.keys())

    def to_tuples(self):
        return [(name, m) for name, lst in self.items() for m in lst]

    def get_attrs(self, *attrs, unique=False, as_pandas=True):
        """reference: Container.get_attrs (module_container.py:97-195)"""
        from .microgrid import container_get_attrs
        return container_get_attrs(self, attrs, unique, as_pandas)

    def __len__(self):      # counts modules, like the reference's container
        return sum(len(v) for v in self.values())

    def _filtered(self, keep):
        return ComposedContainer((n, lst) for n, lst in self.items() if keep(lst[0]))

    fixed = property(lambda self: self._filtered(lambda m: m.module_type[1] == "fixed"))
    flex = property(lambda self: self._filtered(lambda m: m.module_type[1] == "flex"))
    controllable = property(lambda self: self._filtered(lambda m: m.module_type[1] == "controllable"))
    sources = property(lambda self: self._filtered(lambda m: m.is_source and not m.is_sink))
    sinks = property(lambda self: self._filtered(lambda m: m.is_sink and not m.is_source))
    source_and_sinks = property(lambda self: self._filtered(lambda m: m.is_source and m.is_sink))


class ComposedMicrogrid:
    """pymgrid.Microgrid's surface for any module list (see the module docstring).  What the reference raises mid-step
    (AssertionError, ValueError under raise_errors=True) is raised here AFTER the step has been applied with the
    reference's clip semantics -- the engine reports events, it does not unwind."""

    def __init__(self, modules, add_unbalanced_module=True, loss_load_cost=10., overgeneration_cost=2.,
                 reward_shaping_func=None, trajectory_func=None, device=None, obs_order="gym_sorted"):
        if reward_shaping_func is not None and not callable(reward_shaping_func):
            raise TypeError("reward_shaping_func must be callable: f(energy_info, cost_info) -> float (microgrid/utils/step.py:41-46)")
        # B = 1: the shaper is the caller's Python function of the step's info dict, evaluated on the host exactly where
        # the reference evaluates it (every MicrogridStep.balance() and the output, microgrid.py:259, 277, 316, 325)
        self.reward_shaping_func = reward_shaping_func
        comp = Composition(modules, add_unbalanced_module, loss_load_cost, overgeneration_cost, obs_order=obs_order)
        self.composition = comp
        self._batch = ComposedBatch([comp], device=device, obs_order=obs_order, with_info=True)
        self._modules = ComposedContainer()
        for s, r in zip(comp.slots, comp.records):
            self._modules.setdefault(s.name, ModuleList()).append(ComposedModuleView(self, s, r))
        self._views = [v for lst in self._modules.values() for v in lst]        # listing order
        self._log_rows = []
        self._actions = np.zeros((1, comp.n_act))
        self._initial_step, self._final_step = comp.initial_step, comp.final_step
        self.trajectory_func = self._check_trajectory_func(trajectory_func)

    def _check_trajectory_func(self, trajectory_func):
        """reference: Microgrid._check_trajectory_func (microgrid.py:167-199): see trajectory.validated"""
        from .trajectory import validated
        return validated(trajectory_func, self._initial_step, self._final_step)

    def _set_window(self, initial_step, final_step):
        """the modules' episode window (microgrid.py:221-225, 652-684): the env's entry of the batch's per-env window arrays"""
        self._batch.env_initial_step[0], self._batch.env_final_step[0] = int(initial_step), int(final_step)

    # ---- state ----
    def _state(self):
        b = self._batch
        # copies: on a CPU tensor .numpy() is a view that the next step would change under the caller
        return int(b.step_counter[0].item()), b.fstate[0].cpu().numpy().copy(), b.istate[0].cpu().numpy().copy()

    current_step = property(lambda self: int(self._batch.step_counter[0].item()))
    modules = property(lambda self: self._modules)
    fixed = property(lambda self: self._modules.fixed)
    flex = property(lambda self: self._modules.flex)
    controllable = property(lambda self: self._modules.controllable)

    def __len__(self):
        return self.composition.series_len

    def __repr__(self):
        return "Microgrid([" + ", ".join(f"{n} x {len(lst)}" for n, lst in self._modules.items()) + "])"

    @property
    def initial_step(self):
        return self._initial_step

    @initial_step.setter
    def initial_step(self, value):
        self._initial_step = int(value)
        self._set_window(self._initial_step, self._final_step)

    @property
    def final_step(self):
        return self._final_step

    @final_step.setter
    def final_step(self, value):
        self._final_step = int(value)
        self._set_window(self._initial_step, self._final_step)

    # ---- conversions ----
    def _obs_dict(self, row, order):
        out = OrderedDict()
        for s in order:
            out.setdefault(s.name, []).append(views.module_obs(row[s.obs_off:s.obs_off + s.obs_len]))
        return out

    def _info_dict(self, info):
        out = OrderedDict()
        for s in self.composition.dispatch:
            r = info[s.listing * MGC_INFO_SLOTS:(s.listing + 1) * MGC_INFO_SLOTS]
            d = OrderedDict()
            d["absorbed_energy" if r[4] else "provided_energy"] = float(r[1] if r[4] else r[0])
            if s.kind in ("genset", "grid"):
                d["co2_production"] = float(r[2])
            elif s.kind == "renewable":
                d["curtailment"] = float(r[2])
            out.setdefault(s.name, []).append(d)
        return out

    def _control_row(self, control):
        """microgrid.py:262-284: missing modules raise ValueError, extra keys warn, a bare scalar stands for [scalar]"""
        row = self._actions
        row[:] = 0.0
        control = dict(control)
        for name, slots in self.composition.controllable():
            try:
                vals = control.pop(name)
            except KeyError:
                raise ValueError(f'Control for module "{name}" not found. Available controls:\n\t{control.keys()}')
            try:
                pairs = list(zip(slots, vals))
            except TypeError:
                pairs = list(zip(slots, [vals]))
            for s, v in pairs:
                arr = np.asarray(v, dtype=np.float64).reshape(-1)
                if arr.size != s.n_act:
                    raise ValueError(f"Bad action {v}")
                row[0, s.act_col:s.act_col + s.n_act] = arr
        if control:
            warnings.warn(f'\nIgnoring the following keys in passed control:\n {list(control.keys())}')
        return row

    # ---- the hot path ----
    def run(self, control, normalized=True):
        """reference: Microgrid.run (microgrid.py:227-325).  Same arguments, return types and errors."""
        comp, b = self.composition, self._batch
        row = self._control_row(control)
        pre = self._pre_step()
        b.step(row if comp.n_act else None, normalized=normalized)
        return self._finish_step(pre)

    def _pre_step(self):
        """what the reference snapshots before stepping: every module's state for the log (base_module.py:152) and the cost
        info the reward shaper gets (microgrid.py:253)"""
        self._cost_info = self.get_cost_info() if self.reward_shaping_func is not None else None
        return self._state()

    def _finish_step(self, pre):
        """log row, the reference's exceptions from the event flags, and the reference's return types"""
        comp, b = self.composition, self._batch
        flags = int(b.flags[0].item()) & 0xffffffff
        if flags & FLAG_BAD_ACTION:
            raise ValueError("Action not in action space")                                          # envs/discrete/discrete.py:84
        if flags & FLAG_STEP_PAST_END:
            t = self.current_step
            raise IndexError(f"index {t} is out of bounds for axis 0 with size {len(self)}")     # e.g. load_module.py:111
        info = b.info[0].cpu().numpy()
        reward = shaped = float(b.reward[0].item())
        if self.reward_shaping_func is not None:
            shaped = self._shaped_reward(info)
        row = self._log_row(pre, info, reward, shaped)
        stale = self.__dict__.pop("_stale_forecast", None)
        self._log_rows.append(row if stale is None else views.drop_stale_forecasts(row, stale))
        _raise_for(flags)
        if flags & FLAG_BALANCE:
            raise RuntimeError("Microgrid modules unable to balance energy production with consumption.\n")
        return (self._obs_dict(b.obs[0].cpu().numpy(), comp.dispatch), shaped, bool(b.done[0].item()), self._info_dict(info))

    def _shaped_reward(self, info):
        """MicrogridStep.shaped_reward (microgrid/utils/step.py:41-46): the caller's function of (energy info, cost info),
        called where the reference calls it -- after the fixed modules, after the controllable ones, after the flex ones
        (the three balance() calls of Microgrid.run) and once more for the output; the last value is the step's reward."""
        cost_info, full = self._cost_info, self._info_dict(info)
        value = None
        for classes in (("fixed",), ("fixed", "controllable"), ("fixed", "controllable", "flex"), ("fixed", "controllable", "flex")):
            names = {s.name for s in self.composition.dispatch if s.dispatch in classes}
            value = self.reward_shaping_func(OrderedDict((k, v) for k, v in full.items() if k in names), cost_info)
        return value

    def run_priority_list(self, priority_list, n_steps=1):
        """`n_steps` DiscreteMicrogridEnv-style steps with one priority list -- an index into `action_lists`, or a list of
        PriorityListElement-like objects (`.module`, `.action`): what RuleBasedControl.run does every step
        (algos/rbc/rbc.py:87-91 -> priority_list.py:69-116 -> Microgrid.run(normalized=False)), expanded on the device.
        Stops after the step that reports done; returns the last step's (obs, reward, done, info)."""
        index = priority_list if isinstance(priority_list, (int, np.integer)) else self.priority_list_index(priority_list)
        out = None
        act = np.array([int(index)], dtype=np.int32)
        for _ in range(int(n_steps)):
            pre = self._pre_step()
            self._batch.step_discrete(act)
            out = self._finish_step(pre)
            if out[2]:
                break
        return out

    @property
    def action_lists(self):
        """every priority list of the microgrid as PriorityListElement tuples, in the reference's order
        (PriorityListAlgo.get_priority_lists(remove_redundant_gensets=False))"""
        return [self._elements(pl) for pl in (self._batch.action_lists or [])]

    def _elements(self, pl):
        from .algos import PriorityListElement
        return tuple(PriorityListElement(module=(s.name, s.index), module_actions=2 if s.kind == "genset" else 1, action=a,
                                         marginal_cost=self._views[s.listing].marginal_cost) for s, a in pl)

    def priority_list_index(self, priority_list):
        want = tuple((tuple(el.module), int(el.action)) for el in priority_list)
        for k, pl in enumerate(self._batch.action_lists or []):
            if tuple(((s.name, s.index), a) for s, a in pl) == want:
                return k
        raise ValueError('Invalid priority list. Use RuleBasedControl.get_priority_lists to view all valid priority lists.')

    def copy(self):
        """a second microgrid over the same module records carrying this one's live state (the deep copies the reference
        takes in RuleBasedControl.__init__ / BaseMicrogridEnv.from_microgrid)"""
        comp = self.composition
        named = [(s.name, r) for s, r in zip(comp.slots, comp.records)]
        other = ComposedMicrogrid(named, add_unbalanced_module=False, device=self._batch.device if self._batch.device.type == "cuda" else None,
                                  obs_order=comp.obs_order)
        for a in ("step_counter", "fstate", "istate", "env_initial_step", "env_final_step"):
            getattr(other._batch, a).copy_(getattr(self._batch, a))
        other.reward_shaping_func, other.trajectory_func = self.reward_shaping_func, self.trajectory_func
        other._initial_step, other._final_step = self._initial_step, self._final_step
        return other

    def _log_row(self, pre, info, reward, shaped=None):
        """one row of get_log() (log_row_from): `pre` = (t, fstate row, istate row) before the step"""
        t, f, _ = pre
        return log_row_from(self.composition, self.composition.records, t, f, None, self._batch.istate[0].cpu().numpy(), info,
                            reward, shaped)

    def reset(self):
        """reference: Microgrid.reset (microgrid.py:205-225): modules in LISTING order, then 'balance' and 'other'"""
        if self.trajectory_func is not None:      # microgrid.py:221-225: a new episode window per reset
            self._set_window(*self.trajectory_func(self._initial_step, self._final_step))
        obs = self._batch.reset()[0].cpu().numpy()
        flushed, self._log_rows = views.flushed_balance_log(self._log_rows), []
        out = self._obs_dict(obs, self.composition.slots)
        out["balance"], out["other"] = flushed, {}
        return out

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
        """reference: Microgrid.set_forecaster (microgrid.py:477-546): one setting for every time-series module.  None = no
        forecast (horizon 0), "oracle" = perfect forecast.  The batch is rebuilt around the live state; the log is kept."""
        from .params import DEFAULT_HORIZON
        if forecast_horizon is None:
            forecast_horizon = DEFAULT_HORIZON
        comp = self.composition
        if isinstance(forecaster, dict):
            # the reference's dict branch (microgrid.py:520-533) calls set_forecaster on the module LIST of each name and
            # swallows the AttributeError that raises: names are checked, nothing else happens.  Mirrored.
            for name in forecaster:
                if name not in self._modules:
                    raise NameError(f'Unrecognized module {name}.')
            return
        chosen = {s.name: forecaster for s in comp.slots}
        stale = {(s.name, s.index): s.horizon for s in comp.slots if s.obs_len and s.kind in ("load", "renewable", "grid")}
        for s, r in zip(comp.slots, comp.records):
            if s.name in chosen and hasattr(r, "time_series"):
                f = chosen[s.name]
                r.forecaster, r.forecast_horizon = f, forecast_horizon * (f is not None)       # base_timeseries_module.py:237
                r.forecaster_increase_uncertainty, r.forecaster_relative_noise = forecaster_increase_uncertainty, forecaster_relative_noise
        self._rebuild()
        self._stale_forecast = stale       # the next step still logs the forecast computed before the change

    def get_forecast_horizon(self):
        """reference: Microgrid.get_forecast_horizon (microgrid.py:548-582): the horizon the time-series modules share;
        ValueError when they differ, the default (with a warning) when there is none"""
        from .params import DEFAULT_HORIZON
        horizons = [s.horizon for s in self.composition.slots if s.kind in ("load", "renewable", "grid")]
        if not horizons:
            warnings.warn(f"No forecast horizon found in microgrid.modules. Using default horizon {DEFAULT_HORIZON}")
            return DEFAULT_HORIZON
        if min(horizons) != max(horizons):
            raise ValueError(f"Modules have inconsistent forecast horizons: {sorted(set(horizons))}")
        return horizons[0]

    def to_normalized(self, data_dict, act=False, obs=False):
        """reference: Microgrid.to_normalized (microgrid.py:390-409): {name: [value per module]} through each module's space"""
        assert act + obs == 1, 'One of act or obs must be True but not both.'
        return {name: [m.to_normalized(v, act=act, obs=obs) for m, v in zip(lst, data_dict[name])]
                for name, lst in self._modules.items() if name in data_dict}

    def from_normalized(self, data_dict, act=False, obs=False):
        assert act + obs == 1, 'One of act or obs must be True but not both.'
        return {name: [m.from_normalized(v, act=act, obs=obs) for m, v in zip(lst, data_dict[name])]
                for name, lst in self._modules.items() if name in data_dict}

    def set_module_attr(self, attr_name, value):
        """reference: Microgrid.set_module_attr (microgrid.py:584-612): set a constructor attribute on every module that
        has it (e.g. 'forecast_horizon'); AttributeError when none does"""
        hit = False
        for r in self.composition.records:
            if hasattr(r, attr_name):
                setattr(r, attr_name, value)
                hit = True
        if not hit:
            raise AttributeError(f"No module has attribute '{attr_name}'.")
        self._rebuild()

    def _rebuild(self):
        """a new composition / batch from the (modified) module records, carrying over state, windows and the log"""
        old = self._batch
        comp = self.composition
        named = [(s.name, r) for s, r in zip(comp.slots, comp.records)]
        state = {a: getattr(old, a).clone() for a in ("step_counter", "fstate", "istate", "env_initial_step", "env_final_step")}
        keep = (self._log_rows, self.reward_shaping_func, self.trajectory_func, self._initial_step, self._final_step)
        ComposedMicrogrid.__init__(self, named, add_unbalanced_module=False, device=old.device if old.device.type == "cuda" else None,
                                   obs_order=comp.obs_order)
        self._log_rows, self.reward_shaping_func, self.trajectory_func, self._initial_step, self._final_step = keep
        for a, v in state.items():
            getattr(self._batch, a).copy_(v)

    # ---- actions ----
    def sample_action(self, strict_bound=False, sample_flex_modules=False):
        """reference: Microgrid.sample_action (microgrid.py:337-362)"""
        it = self._modules if sample_flex_modules else self._modules.controllable
        return {name: [m.sample_action(strict_bound=strict_bound) for m in lst] for name, lst in it.items()
                if lst[0].action_space.shape[0]}

    def get_empty_action(self, sample_flex_modules=False):
        it = self._modules if sample_flex_modules else self._modules.controllable
        return {name: [None] * len(lst) for name, lst in it.items() if lst[0].action_space.shape[0]}

    # ---- introspection ----
    def state_dict(self, normalized=False):
        return {name: [m.state_dict(normalized=normalized) for m in lst] for name, lst in self._modules.items()}

    def state_series(self, normalized=False):
        import pandas as pd
        data = OrderedDict(((name, j, k), v) for name, lst in self._modules.items() for j, m in enumerate(lst)
                           for k, v in m.state_dict(normalized=normalized).items())
        return pd.Series(data, dtype=np.float64)

    def get_log(self, as_frame=True, drop_singleton_key=False):
        """reference: Microgrid.get_log (microgrid.py:434-475): one row per step since the last reset"""
        df = views.log_frame(self._log_rows, self.current_step, drop_singleton_key)
        return df if as_frame else df.to_dict()

    log = property(lambda self: self.get_log())


def in_fused_scope(modules, add_unbalanced_module=True):
    """True when the module list is one the fused kernels cover: exactly one load, renewable, battery and slack module, at
    most one genset and one grid, canonical names, one forecast horizon (modules.params_from_modules' conditions)."""
    try:
        named = _named(list(modules))
    except TypeError:
        return True        # let the fused constructor raise its own error
    counts, names, horizons = {}, {}, set()
    for name, m in named:
        k = m.module_type[0]
        counts[k] = counts.get(k, 0) + 1
        names[k] = name
        if hasattr(m, "forecast_horizon"):
            horizons.add(m.forecast_horizon)
    if add_unbalanced_module:
        counts["balancing"] = counts.get("balancing", 0) + 1
    need = {"load": (1, 1), "renewable": (1, 1), "battery": (1, 1), "balancing": (1, 1), "genset": (0, 1), "grid": (0, 1)}
    if any(not lo <= counts.get(k, 0) <= hi for k, (lo, hi) in need.items()):
        return False
    if any(names.get(k, k) != k for k in ("battery", "genset", "grid", "load")):
        return False
    if "battery" <= names["renewable"] <= "load" or len(horizons) != 1:
        return False
    # the container keeps INSERTION order inside a cell (module_container.py:355-413), and that order is the dispatch order
    # and the summation order of Microgrid.run: the fused kernels step battery before grid and the renewable before the slack
    # module, so lists given the other way round take the general path
    position = {m.module_type[0]: i for i, (_, m) in enumerate(named)}
    if "grid" in position and position["grid"] < position["battery"]:
        return False
    if "balancing" in position and position["balancing"] < position["renewable"]:
        return False
    return True


# ---- Gym-style envs and rule-based control on composed microgrids -------------------------------------------------------
class _ComposedEnv:
    """The reference's BaseMicrogridEnv surface (envs/base/base.py:84-223) for any module list: `batch=None` is one
    microgrid with the reference's host types, `batch=B` steps B replicas with device tensors."""

    def __init__(self, modules, add_unbalanced_module=True, loss_load_cost=10., overgeneration_cost=2., reward_shaping_func=None,
                 trajectory_func=None, flat_spaces=True, observation_keys=(), batch=None, device=None, obs_order="gym_sorted",
                 ):
        from .envs import Box
        if not flat_spaces:
            raise NotImplementedError("flat_spaces=False (nested gym spaces) is not part of the batched surface")
        self.single = batch is None
        if not self.single and reward_shaping_func is not None:
            raise NotImplementedError("batched composed envs: a reward_shaping_func is a Python callable and runs for single "
                                      "microgrids only (the fused module set has on-device shapers)")
        self._mg = ComposedMicrogrid(modules, add_unbalanced_module, loss_load_cost, overgeneration_cost, reward_shaping_func,
                                     trajectory_func, device=device, obs_order=obs_order) \
            if not isinstance(modules, ComposedMicrogrid) else modules
        self.trajectory_func = self._mg.trajectory_func
        comp = self.composition = self._mg.composition
        if self.single:
            self.batch = self._mg._batch
        else:
            self.batch = ComposedBatch([comp], np.zeros(int(batch), dtype=np.int64), device=device, obs_order=comp.obs_order,
                                       observation_keys=observation_keys)
            for a in ("step_counter", "fstate", "istate"):      # replicas start from the microgrid's live state
                getattr(self.batch, a).copy_(getattr(self._mg._batch, a).expand_as(getattr(self.batch, a)))
        self.n_envs = self.batch.n_envs
        # observation_keys (base.py:109-163, 211-218): a batch writes only the selected elements (ComposedBatch); a single
        # microgrid keeps its full row for Microgrid.run's dicts and the env picks the selected elements out of it
        observation_keys = observation_keys or ()      # (None is the reference DiscreteMicrogridEnv's own default)
        self.observation_keys = [observation_keys] if isinstance(observation_keys, str) else list(observation_keys)
        self._take = None
        if self.observation_keys:
            self._take = np.array([s.obs_off + k for s, k in comp.select_observation(self.observation_keys)], dtype=np.int64)
        self.observation_space = Box(0.0, 1.0, (len(self._take) if self._take is not None else comp.obs_dim,))     # base.py:161-163

    @classmethod
    def from_microgrid(cls, microgrid, **kw):
        """reference: BaseMicrogridEnv.from_microgrid (envs/base/base.py:270-290): an env over a copy of a (possibly
        running) microgrid, state included"""
        if not isinstance(microgrid, ComposedMicrogrid):
            raise TypeError("from_microgrid needs a composed pymgrid_b200.Microgrid")
        return cls(microgrid.copy(), **kw)

    modules = property(lambda self: self._mg.modules)
    fixed = property(lambda self: self._mg.fixed)
    flex = property(lambda self: self._mg.flex)
    controllable = property(lambda self: self._mg.controllable)
    initial_step = property(lambda self: self._mg.initial_step)
    final_step = property(lambda self: self._mg.final_step)
    log = property(lambda self: self._mg.log)

    @property
    def current_step(self):
        return self._mg.current_step if self.single else self.batch.step_counter

    def __len__(self):
        return len(self._mg)

    def get_log(self, *a, **kw):
        return self._mg.get_log(*a, **kw)

    def reset(self, mask=None):
        """reference: BaseMicrogridEnv.reset (base.py:165-167): the flat observation after Microgrid.reset"""
        if self.single:
            self._mg.reset()
            return self._single_obs()
        if self.trajectory_func is not None:      # microgrid.py:221-225: a new episode window per reset, per env
            self._draw_windows(mask)
        return self.batch.reset(mask)

    def _draw_windows(self, mask):
        """one (initial_step, final_step) pair per env being reset; the vectorised classes of pymgrid_b200.trajectory draw all
        of them in one call (`n=`), a plain reference-style callable is called once per env"""
        from .trajectory import draw
        lo, hi, n = self._mg.initial_step, self._mg.final_step, self.n_envs
        initial, final = draw(self.trajectory_func, lo, hi, n)
        if mask is not None:                      # envs that keep running keep their window
            keep = ~np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask, dtype=bool).reshape(n)
            initial[keep] = self.batch.env_initial_step.cpu().numpy()[keep]
            final[keep] = self.batch.env_final_step.cpu().numpy()[keep]
        self.batch.set_trajectories(initial, final)

    def _single_obs(self):
        row = self.batch.obs[0].cpu().numpy()
        return row[self._take] if self._take is not None else row.copy()

    def _single_result(self, out):
        obs, reward, done, info = out
        return self._single_obs(), reward, done, info


class ComposedDiscreteEnv(_ComposedEnv):
    """Action = index of a priority list (reference: envs/discrete/discrete.py:60-143)"""

    def __init__(self, modules, *args, remove_redundant_gensets=True, **kw):
        from .envs import Discrete
        super().__init__(modules, *args, **kw)
        full = self.batch.action_lists
        if full is None:
            raise NotImplementedError("no discrete action table for this composition")
        kept = self.composition.priority_lists(remove_redundant_gensets)
        key = lambda pl: tuple((s.listing, a) for s, a in pl)      # noqa: E731
        where = {key(pl): k for k, pl in enumerate(full)}
        self._index = np.array([where[key(pl)] for pl in kept], dtype=np.int32)     # env action -> row of the device table
        self.actions_list = [self._mg._elements(pl) for pl in kept]
        self.action_space = Discrete(len(kept))
        self._index_dev = torch.from_numpy(self._index).to(self.batch.device)

    def remove_action(self, action_number):
        """reference: DiscreteMicrogridEnv.remove_action (envs/discrete/discrete.py:90-105)"""
        from .envs import Discrete
        if action_number not in self.action_space:
            raise ValueError('Cannot remove action that is not in the action space!')
        self.actions_list.pop(action_number)
        self._index = np.delete(self._index, action_number)
        self.action_space = Discrete(self.action_space.n - 1)
        self._index_dev = torch.from_numpy(self._index).to(self.batch.device)

    def step(self, action):
        if self.single:
            if action not in self.action_space:
                raise ValueError(f" Action {action} not in action space {self.action_space}")       # discrete.py:84
            out = self._mg.run_priority_list(int(self._index[int(action)]), 1)
            self._mg._log_rows[-1][("action", 0, "")] = int(action)      # discrete.py:141 logs the action it was given
            return self._single_result(out)
        a = torch.as_tensor(action, device=self.batch.device).to(torch.int64)
        bad = (a < 0) | (a >= len(self._index))
        mapped = torch.where(bad, torch.full_like(a, -1), self._index_dev.to(torch.int64)[a.clamp(0, len(self._index) - 1)])
        obs, reward, done, _ = self.batch.step_discrete(mapped.to(torch.int32))
        return obs, reward, done, {"flags": self.batch.flags}

    def sample_action(self, strict_bound=False, sample_flex_modules=False):
        if self.single:
            return self.action_space.sample()
        return torch.randint(0, self.action_space.n, (self.n_envs,), dtype=torch.int32, device=self.batch.device)


class ComposedContinuousEnv(_ComposedEnv):
    """Action = flat vector in [0,1]^n_act: the controllable modules' normalised actions in Microgrid.controllable's order
    (`action_layout`: (name, index) -> first column); the intended semantics of the reference's ContinuousMicrogridEnv
    """

    def __init__(self, modules, *args, **kw):
        from .envs import Box
        super().__init__(modules, *args, **kw)
        self.action_space = Box(0.0, 1.0, (self.composition.n_act,))
        self.action_layout = {(s.name, s.index): s.act_col for s in self.composition.dispatch if s.n_act}

    def step(self, action, normalized=True):
        if self.single:
            row = np.asarray(action, dtype=np.float64).reshape(-1)
            control = {name: [row[s.act_col:s.act_col + s.n_act] if s.n_act > 1 else row[s.act_col] for s in slots]
                       for name, slots in self.composition.controllable()}
            return self._single_result(self._mg.run(control, normalized=normalized))
        obs, reward, done, _ = self.batch.step(action, normalized=normalized)
        return obs, reward, done, {"flags": self.batch.flags}

    def sample_action(self, strict_bound=False, sample_flex_modules=False):
        if self.single:
            return self.action_space.sample()
        return torch.rand((self.n_envs, self.composition.n_act), dtype=torch.float64, device=self.batch.device)


class ComposedRuleBasedControl:
    """pymgrid.algos.RuleBasedControl (algos/rbc/rbc.py:7-140) on a composed microgrid: a fixed priority list -- by default
    the first one sorted by marginal cost, ties towards the higher action number (rbc.py:31-44,
    priority_list_element.py:73-80) -- deployed every step on the device."""

    def __init__(self, microgrid, priority_list=None, remove_redundant_gensets=True):
        if not isinstance(microgrid, ComposedMicrogrid):
            raise TypeError("ComposedRuleBasedControl needs a composed pymgrid_b200.Microgrid")
        self._microgrid = microgrid.copy()          # rbc.py:28-30: the controller works on a copy
        self._remove_redundant_gensets = remove_redundant_gensets
        lists = self.get_priority_lists(remove_redundant_gensets)
        if priority_list is None:
            priority_list = sorted(lists[0])
        elif tuple(priority_list) not in [tuple(pl) for pl in lists]:
            raise ValueError('Invalid priority list. Use RuleBasedControl.get_priority_lists to view all '
                             'valid priority lists.')
        self._priority_list = list(priority_list)
        self._index = self._microgrid.priority_list_index(self._priority_list)

    def get_priority_lists(self, remove_redundant_gensets=None):
        if remove_redundant_gensets is None:
            remove_redundant_gensets = self._remove_redundant_gensets
        return [self._microgrid._elements(pl) for pl in self._microgrid.composition.priority_lists(remove_redundant_gensets)]

    def reset(self):
        return self._microgrid.reset()

    def run(self, max_steps=None, verbose=False):
        """reference: RuleBasedControl.run (rbc.py:64-93)"""
        if max_steps is None:
            max_steps = len(self._microgrid)
        self.reset()
        self._microgrid.run_priority_list(self._index, max_steps)
        return self._microgrid.get_log(as_frame=True)

    def get_empty_action(self):
        return self._microgrid.get_empty_action()

    microgrid = property(lambda self: self._microgrid)
    fixed = property(lambda self: self._microgrid.fixed)
    flex = property(lambda self: self._microgrid.flex)
    modules = property(lambda self: self._microgrid.modules)
    priority_list = property(lambda self: self._priority_list)


# ---- a module on its own: the reference's operator API (BaseMicrogridModule.step / reset / state) ------------------------
class StandaloneModule:
    """`module.step(action, normalized)` without a Microgrid around it (modules/base/base_module.py:95-159; the reference's
    module-level tests use its modules this way): a batch of one holding just this module, stepped by mgc_modules_step.
    Built lazily by pymgrid_b200.modules' classes on the first `step()` / `reset()` / live-attribute access."""

    def __init__(self, record, device=None):
        name = record.module_type[0]
        self.mg = ComposedMicrogrid([(name, record)], add_unbalanced_module=False, obs_order="container", device=device)
        self.view = self.mg.modules[name][0]
        self.kind = record.module_type[0]
        self.width = 2 if self.kind == "genset" else 0 if self.kind == "load" else 1
        self.log_rows = []

    def step(self, action, normalized=True):
        b = self.mg._batch
        row = None
        if self.width:
            try:
                arr = np.asarray(action, dtype=np.float64).reshape(-1)
            except (TypeError, ValueError):
                raise ValueError(f'Bad action {action}')
            if arr.size != self.width:
                raise ValueError(f'Bad action {action}')
            row = arr.reshape(1, -1)
        pre = self.mg._state()
        b.modules_step(row, normalized=normalized)
        flags = int(b.flags[0].item()) & 0xffffffff
        if flags & FLAG_NOT_A_SINK:         # raised before the module reads its series, so also ahead of the IndexError
            _raise_for(flags)
        if flags & FLAG_STEP_PAST_END:
            raise IndexError(f"index {self.mg.current_step} is out of bounds for axis 0 with size {len(self.mg)}")
        _raise_for(flags)
        info = b.info[0].cpu().numpy()
        reward = float(b.reward[0].item())
        full = self.mg._log_row(pre, info, reward)
        self.log_rows.append(OrderedDict((k[2], v) for k, v in full.items() if k[0] != "balance"))
        return (b.obs[0].cpu().numpy().copy(), reward, bool(b.done[0].item()), self.mg._info_dict(info)[self.view.name[0]][0])

    def reset(self):
        """BaseMicrogridModule.reset (base_module.py:65-77): step = initial_step, log flushed, normalised state returned"""
        obs = self.mg._batch.reset()[0].cpu().numpy().copy()
        self.log_rows = []
        return obs
