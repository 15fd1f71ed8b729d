"""ctypes binding of the C oracle (oracle/mg_oracle.c) -- TEST INFRASTRUCTURE, not product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
`OracleGrid` wraps one `OrcGrid`; `build()` compiles liboracle.so with gcc (-ffp-contract=off).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
N_INFO = 16
INFO_NAMES = ("load_met", "pv_used", "curtailment", "loss_load", "overgeneration", "genset_production",
              "genset_co2", "battery_discharge", "battery_charge", "grid_import", "grid_export", "grid_co2",
              "reward_genset", "reward_battery", "reward_grid", "reward_unbalanced")
ORDER_GYM_SORTED, ORDER_CONTAINER = 0, 1
MOD_GENSET, MOD_BATTERY, MOD_GRID = 0, 1, 2


class OrcGrid(C.Structure):
    _fields_ = [
        ("has_genset", C.c_int32), ("has_grid", C.c_int32), ("horizon", C.c_int32), ("T", C.c_int32),
        ("initial_step", C.c_int32), ("final_step", C.c_int32),
        ("min_capacity", C.c_double), ("max_capacity", C.c_double), ("max_charge", C.c_double),
        ("max_discharge", C.c_double), ("efficiency", C.c_double), ("battery_cost_cycle", C.c_double),
        ("running_min_production", C.c_double), ("running_max_production", C.c_double), ("genset_cost", C.c_double),
        ("co2_per_unit", C.c_double), ("gen_cost_per_unit_co2", C.c_double),
        ("start_up_time", C.c_int32), ("wind_down_time", C.c_int32), ("allow_abortion", C.c_int32), ("reward_shaper", C.c_int32),
        ("max_import", C.c_double), ("max_export", C.c_double), ("grid_cost_per_unit_co2", C.c_double),
        ("loss_load_cost", C.c_double), ("overgeneration_cost", C.c_double),
        ("load_ts", C.POINTER(C.c_double)), ("pv_ts", C.POINTER(C.c_double)), ("grid_ts", C.POINTER(C.c_double)),
        ("t", C.c_int32), ("cs", C.c_int32), ("gs", C.c_int32), ("up", C.c_int32), ("dn", C.c_int32), ("_pad1", C.c_int32),
        ("charge", C.c_double), ("soc", C.c_double),
        ("prepared", C.c_int32), ("_pad2", C.c_int32),
        ("load_low", C.c_double), ("load_high", C.c_double), ("pv_low", C.c_double), ("pv_high", C.c_double),
        ("grid_low", C.c_double * 4), ("grid_high", C.c_double * 4),
    ]


def build(force=False):
    src = os.path.join(_HERE, "mg_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(src[:-2] + ".h"))):
        return _LIB_PATH
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-o", _LIB_PATH, src,
                           "-lm", "-lpthread"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        P = C.POINTER
        L.orc_prepare.argtypes = [P(OrcGrid)]
        L.orc_obs_dim.argtypes = [P(OrcGrid)]
        L.orc_n_act.argtypes = [P(OrcGrid)]
        L.orc_run.argtypes = [P(OrcGrid), P(C.c_double), C.c_int, C.c_int, P(C.c_double), P(C.c_double),
                              P(C.c_int32), P(C.c_double), P(C.c_uint32)]
        L.orc_observe.argtypes = [P(OrcGrid), C.c_int, P(C.c_double)]
        L.orc_reset.argtypes = [P(OrcGrid)]
        L.orc_genset_update_status.argtypes = [P(OrcGrid), C.c_double]
        L.orc_genset_next_status.argtypes = [P(OrcGrid), C.c_int]
        L.orc_priority_control.argtypes = [P(OrcGrid), P(C.c_int8), P(C.c_int8), C.c_int, P(C.c_double)]
        L.orc_rollout.argtypes = [P(OrcGrid), C.c_int64, P(C.c_double), C.c_int32, C.c_int32, C.c_int, C.c_int,
                                  P(C.c_double), P(C.c_uint8), P(C.c_double), C.c_int32, C.c_int32]
        L.orc_rollout_discrete.argtypes = [P(OrcGrid), C.c_int64, P(C.c_int32), C.c_int32, P(C.c_int8), P(C.c_int8),
                                           P(C.c_int32), C.c_int32, C.c_int, P(C.c_double), P(C.c_uint8),
                                           P(C.c_double), C.c_int32, C.c_int32]
        L.orc_priority_control.restype = C.c_uint32
        for f in (L.orc_prepare, L.orc_run, L.orc_observe, L.orc_reset, L.orc_genset_update_status,
                  L.orc_rollout, L.orc_rollout_discrete):
            f.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fill_struct(g, p, keep):
    """Fill an OrcGrid from any object with the attribute layout of pymgrid_b200.params.MicrogridParams.
    `keep` collects the numpy arrays whose memory the struct points into."""
    # the oracle always works on explicit series: profile * scale is materialised here (MicrogridGenerator grids)
    load = np.ascontiguousarray(getattr(p, "effective_load_ts", p.load_ts), dtype=np.float64)
    pv = np.ascontiguousarray(getattr(p, "effective_pv_ts", p.pv_ts), dtype=np.float64)
    keep += [load, pv]
    g.has_genset, g.has_grid = int(p.genset is not None), int(p.grid is not None)
    g.horizon, g.T = int(p.forecast_horizon), len(load)
    g.initial_step, g.final_step = int(p.initial_step), int(p.final_step)
    b = p.battery
    g.min_capacity, g.max_capacity, g.max_charge = b.min_capacity, b.max_capacity, b.max_charge
    g.max_discharge, g.efficiency, g.battery_cost_cycle = b.max_discharge, b.efficiency, b.battery_cost_cycle
    g.charge = b.current_charge
    soc = getattr(b, "soc", None)      # _soc as constructed (init_soc); derived when the record does not carry one
    g.soc = b.current_charge / b.max_capacity if soc is None else soc
    if p.genset is not None:
        s = p.genset
        g.running_min_production, g.running_max_production = s.running_min_production, s.running_max_production
        g.genset_cost, g.co2_per_unit, g.gen_cost_per_unit_co2 = s.genset_cost, s.co2_per_unit, s.cost_per_unit_co2
        g.start_up_time, g.wind_down_time, g.allow_abortion = s.start_up_time, s.wind_down_time, int(s.allow_abortion)
        g.cs, g.gs, g.up, g.dn = s.current_status, s.goal_status, s.steps_until_up, s.steps_until_down
    if p.grid is not None:
        ts = p.grid.effective_time_series() if hasattr(p.grid, "effective_time_series") else p.grid.time_series
        ts = np.ascontiguousarray(ts, dtype=np.float64)
        keep.append(ts)
        g.max_import, g.max_export, g.grid_cost_per_unit_co2 = p.grid.max_import, p.grid.max_export, p.grid.cost_per_unit_co2
        g.grid_ts = _dp(ts)
    g.loss_load_cost, g.overgeneration_cost = p.loss_load_cost, p.overgeneration_cost
    g.reward_shaper = {None: 0, "pv_curtailment": 1, "battery_discharge": 2}[getattr(p, "reward_shaper", None)]
    g.load_ts, g.pv_ts = _dp(load), _dp(pv)
    g.t = int(p.current_step)
    g.prepared = 0


class OracleGrid:
    """One microgrid stepped by the C oracle; mirrors Microgrid.run/reset for the tests."""

    def __init__(self, params, order=ORDER_GYM_SORTED):
        self._keep = []
        self.g = OrcGrid()
        fill_struct(self.g, params, self._keep)
        self.order = order
        lib().orc_prepare(C.byref(self.g))
        self.obs_dim = lib().orc_obs_dim(C.byref(self.g))
        self.n_act = lib().orc_n_act(C.byref(self.g))

    def run(self, control, normalized=True):
        ctrl = np.ascontiguousarray(control, dtype=np.float64)
        assert ctrl.shape == (self.n_act,)
        obs = np.empty(self.obs_dim)
        info = np.empty(N_INFO)
        r, d, e = C.c_double(), C.c_int32(), C.c_uint32()
        lib().orc_run(C.byref(self.g), _dp(ctrl), int(normalized), self.order, _dp(obs), C.byref(r), C.byref(d),
                      _dp(info), C.byref(e))
        return obs, r.value, bool(d.value), info, e.value

    def observe(self):
        obs = np.empty(self.obs_dim)
        lib().orc_observe(C.byref(self.g), self.order, _dp(obs))
        return obs

    def reset(self):
        lib().orc_reset(C.byref(self.g))
        return self.observe()

    def priority_control(self, plist):
        """plist: sequence of (module, action) with module in {MOD_GENSET, MOD_BATTERY, MOD_GRID}."""
        mods = np.array([m for m, _ in plist], dtype=np.int8)
        acts = np.array([a for _, a in plist], dtype=np.int8)
        out = np.empty(self.n_act)
        self.list_flags = lib().orc_priority_control(C.byref(self.g), mods.ctypes.data_as(C.POINTER(C.c_int8)),
                                                     acts.ctypes.data_as(C.POINTER(C.c_int8)), len(plist), _dp(out))
        return out

    @property
    def state(self):
        g = self.g
        return dict(t=g.t, charge=g.charge, genset=(g.cs, g.gs, g.up, g.dn))


class OracleBatch:
    """Array of OrcGrid for the batched drivers (CPU baseline timing, bulk parity)."""

    def __init__(self, params_list, order=ORDER_GYM_SORTED):
        self._keep = []
        self.n = len(params_list)
        self.arr = (OrcGrid * self.n)()
        cache = {}
        for i, p in enumerate(params_list):
            fill_struct(self.arr[i], p, self._keep if id(p) not in cache else [])
            if id(p) in cache:   # replicas share the series memory of the first instance
                j = cache[id(p)]
                self.arr[i].load_ts, self.arr[i].pv_ts, self.arr[i].grid_ts = \
                    self.arr[j].load_ts, self.arr[j].pv_ts, self.arr[j].grid_ts
            else:
                cache[id(p)] = i
            lib().orc_prepare(C.byref(self.arr[i]))
        self.order = order
        self.obs_dims = np.array([lib().orc_obs_dim(C.byref(self.arr[i])) for i in range(self.n)])
        self.n_acts = np.array([lib().orc_n_act(C.byref(self.arr[i])) for i in range(self.n)])

    def rollout(self, actions, normalized=True, n_threads=1, want_obs=True):
        """actions [n_steps, n, max_act] float64 (container order, padded).  Returns rewards, dones, last obs."""
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        n_steps, n, max_act = actions.shape
        assert n == self.n
        rewards = np.empty((n_steps, n))
        dones = np.empty((n_steps, n), dtype=np.uint8)
        stride = int(self.obs_dims.max())
        obs = np.zeros((n, stride)) if want_obs else None
        lib().orc_rollout(self.arr, n, _dp(actions), max_act, n_steps, int(normalized), self.order, _dp(rewards),
                          dones.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(obs) if want_obs else None, stride,
                          n_threads)
        return rewards, dones, obs

    def rollout_discrete(self, actions, lut_module, lut_action, lut_offset, n_threads=1):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        n_steps, n = actions.shape
        lut_module = np.ascontiguousarray(lut_module, dtype=np.int8)
        lut_action = np.ascontiguousarray(lut_action, dtype=np.int8)
        lut_offset = np.ascontiguousarray(lut_offset, dtype=np.int32)
        rewards = np.empty((n_steps, n))
        dones = np.empty((n_steps, n), dtype=np.uint8)
        stride = int(self.obs_dims.max())
        obs = np.zeros((n, stride))
        I8 = C.POINTER(C.c_int8)
        lib().orc_rollout_discrete(self.arr, n, actions.ctypes.data_as(C.POINTER(C.c_int32)), n_steps,
                                   lut_module.ctypes.data_as(I8), lut_action.ctypes.data_as(I8),
                                   lut_offset.ctypes.data_as(C.POINTER(C.c_int32)), lut_module.shape[1], self.order,
                                   _dp(rewards), dones.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(obs), stride,
                                   n_threads)
        return rewards, dones, obs

    def state(self):
        t = np.array([g.t for g in self.arr], dtype=np.int32)
        charge = np.array([g.charge for g in self.arr])
        gen = np.array([(g.cs, g.gs, g.up, g.dn) for g in self.arr], dtype=np.int32)
        return t, charge, gen
