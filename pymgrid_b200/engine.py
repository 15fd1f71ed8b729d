"""BatchedMicrogrid: the Python host of the B200 engine.

Holds per-module parameters and the load / PV / grid time series as torch DEVICE tensors, groups the envs by
architecture (has_genset, has_grid, forecast_horizon), and calls the hand-written sm_100a kernels through the
C-ABI of include/pymgrid_b200.h.  It is the batched counterpart of the reference's `Microgrid`
(src/pymgrid/microgrid/microgrid.py): `step` = `Microgrid.run` / `BaseMicrogridEnv.step` for B envs at once,
`step_discrete` = `DiscreteMicrogridEnv.step`, `reset` = `Microgrid.reset`.

torch is plumbing here (device memory, streams); every number is produced by the CUDA extension.  There is no
CPU path: constructing a BatchedMicrogrid without a CUDA device or without the built extension raises.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import (MgForecastNoise, MG_ABI_VERSION, MG_MAX_GROUPS, MG_N_INFO, MG_N_LOG, MG_OBS_CONTAINER, MG_OBS_GYM_SORTED, EngineError,
                    MgConfig, MgHostRolloutIO, MgLayout, MgPriorityList, MgRolloutIO, MgStepIO)
from .params import REWARD_SHAPERS, MicrogridParams
from .priority_list import priority_lists

OBS_ORDERS = {"gym_sorted": MG_OBS_GYM_SORTED, "container": MG_OBS_CONTAINER,
              "gym_sorted_pv_first": _cabi.MG_OBS_GYM_SORTED_PV_FIRST}


def _spread(low, high):
    s = high - low                      # utils/space.py:204-205
    return 1.0 if s == 0 else s


def scaled_bounds(ts_min, ts_max, scale):
    """Observation bounds of the series `profile * scale` from the profile's extremes (multiplication by a positive
    scale is monotonic in IEEE arithmetic, so min / max commute with it exactly), then the reference's rules:
    pull towards zero (base_timeseries_module.py:81-88), spread 0 -> 1 (utils/space.py:204-205), forecaster fill
    (high + low) / 2 clipped to the bounds (forecast/forecaster.py:95, 139-149) and normalised."""
    low, high = ts_min * scale, ts_max * scale
    if low > 0:
        low = 0.0
    elif high < 0:
        high = 0.0
    spread = _spread(low, high)
    fill = min(max((high + low) / 2, low), high)
    return low, high, spread, (fill - low) / spread


def config_record(p: MicrogridParams, load_series, pv_series, grid_series, plist_offset, plist_count):
    """MgConfig of one parameter set; derived constants computed in f64 exactly as the reference's ModuleSpace does."""
    c = MgConfig()
    b = p.battery
    c.bat_min_capacity, c.bat_max_capacity = b.min_capacity, b.max_capacity
    c.bat_max_charge, c.bat_max_discharge = b.max_charge, b.max_discharge
    c.bat_efficiency, c.bat_cost_cycle = b.efficiency, b.battery_cost_cycle
    lo, hi = -b.max_discharge / b.efficiency, b.max_charge * b.efficiency
    c.bat_act_low, c.bat_act_spread = lo, _spread(lo, hi)
    min_soc = b.min_capacity / b.max_capacity
    c.bat_soc_low, c.bat_soc_spread = min_soc, _spread(min_soc, 1.0)
    c.bat_charge_spread = _spread(float(b.min_capacity), float(b.max_capacity))
    if p.genset is not None:
        g = p.genset
        if not (0 <= g.start_up_time <= 255 and 0 <= g.wind_down_time <= 255):
            raise ValueError("genset start_up_time / wind_down_time must fit in 8 bits")
        c.gen_running_min, c.gen_running_max, c.gen_cost = g.running_min_production, g.running_max_production, g.genset_cost
        c.gen_co2_per_unit, c.gen_cost_per_unit_co2 = g.co2_per_unit, g.cost_per_unit_co2
        c.gen_act_spread = _spread(0.0, float(g.running_max_production))
        c.gen_up_spread, c.gen_down_spread = _spread(0.0, float(g.start_up_time)), _spread(0.0, float(g.wind_down_time))
        c.gen_start_up_time, c.gen_wind_down_time, c.gen_allow_abortion = g.start_up_time, g.wind_down_time, int(g.allow_abortion)
    else:
        c.gen_act_spread = c.gen_up_spread = c.gen_down_spread = 1.0
    if p.grid is not None:
        g = p.grid
        c.grid_max_import, c.grid_max_export, c.grid_cost_per_unit_co2 = g.max_import, g.max_export, g.cost_per_unit_co2
        glo = -1 * g.max_export
        c.grid_act_low, c.grid_act_spread = glo, _spread(glo, float(g.max_import))
    else:
        c.grid_act_spread = 1.0
    c.loss_load_cost, c.overgeneration_cost = p.loss_load_cost, p.overgeneration_cost
    c.reward_shaper = REWARD_SHAPERS[p.reward_shaper]
    c.load_scale, c.pv_scale = p.load_scale, p.pv_scale
    if p.scaled:
        c.series_scaled = 1
        for name, ts, scale in (("load", p.load_ts, p.load_scale), ("pv", p.pv_ts, p.pv_scale)):
            low, high, spread, fill_nrm = scaled_bounds(float(ts.min()), float(ts.max()), scale)
            setattr(c, f"{name}_low", low)
            setattr(c, f"{name}_spread", spread)
            setattr(c, f"{name}_fill_nrm", fill_nrm)
    if p.grid is not None and p.grid.status is not None:
        c.grid_status_weak = int(np.min(p.grid.status) < 1)
    c.load_series, c.pv_series, c.grid_series = load_series, pv_series, grid_series
    c.initial_step, c.final_step = p.initial_step, p.final_step
    c.plist_offset, c.plist_count = plist_offset, plist_count
    return c


def forecast_noise_record(p: MicrogridParams):
    """MgForecastNoise of one config: the reference's per-module noise standard deviation (GaussianNoiseForecaster.
    _get_noise_std, forecast/forecaster.py:237-250; series window of set_forecaster, base_timeseries_module.py:233-240)
    divided by the column's observation spread, 0 for constant columns (the forecaster's clip pins them)."""
    rec = MgForecastNoise()
    stop = p.final_step if p.final_step > 0 else len(p)

    def sigmas(name, ts, pull_zero):
        f = p.forecasters.get(name)
        ts = ts.reshape(len(ts), -1)
        if f is None or f.noise_std == 0:
            return [0.0] * ts.shape[1], 0
        std = float(f.noise_std)
        if f.relative_noise:
            std = std * float(np.abs(ts[p.initial_step:stop].mean()))
        out = []
        for c in range(ts.shape[1]):
            low, high = float(ts[:, c].min()), float(ts[:, c].max())
            if pull_zero:
                low, high = min(low, 0.0), max(high, 0.0)
            out.append(std / (high - low) if high > low else 0.0)
        return out, int(bool(f.increase_uncertainty))

    (rec.load_sigma,), rec.load_increase = sigmas("load", p.load_ts * p.load_scale, True)
    (rec.pv_sigma,), rec.pv_increase = sigmas("pv", p.pv_ts * p.pv_scale, True)
    if p.grid is not None:
        g, rec.grid_increase = sigmas("grid", p.grid.effective_time_series(), False)
        for c in range(4):
            rec.grid_sigma[c] = g[c]
    elif "grid" in p.forecasters:
        raise ValueError("forecasters: this microgrid has no grid module")
    return rec


@dataclass
class Group:
    """One architecture group: envs sharing the observation / action row layout."""
    arch: tuple                    # (has_genset, has_grid, horizon)
    env_ids: np.ndarray            # global env ids, in slot order
    n_act: int
    obs_dim: int
    act_cols: dict                 # module name -> first column of the action row
    step: torch.Tensor = None      # int32 [n]
    charge: torch.Tensor = None    # f64 [n]
    genset: torch.Tensor = None    # int32 [n] (packed cs | gs<<8 | up<<16 | dn<<24) or None
    soc_reported: Optional[torch.Tensor] = None   # f64 [n]: the soc the batteries were constructed with (mg_set_reported_soc)
    cfg_index: torch.Tensor = None
    env_initial_step: Optional[torch.Tensor] = None
    env_final_step: Optional[torch.Tensor] = None
    status_bits: Optional[torch.Tensor] = None   # int32 [n, status_words] per-env grid status bitmask (weak grids)
    status_words: int = 0
    obs: torch.Tensor = None       # f64 [n, obs_dim] default output buffer
    reward: torch.Tensor = None
    done: torch.Tensor = None
    info: Optional[torch.Tensor] = None
    flags: Optional[torch.Tensor] = None
    n_actions: int = 0             # discrete action-space size (max over the group's configs)

    @property
    def n_envs(self):
        return len(self.env_ids)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class HostIO:
    """Host-side view of one step for callers whose actions live in host memory (the reference's callers all do).

    `actions` (list of pinned [n, n_act] host tensors of `bm.action_dtype`, one per group; int32 [n] when discrete) is filled by
    the caller; `step()` uploads them with ONE host->device copy, launches the fused kernel and brings reward + done
    back with ONE device->host copy; `reward` / `done` are pinned host views valid after `sync()`.  Observations stay
    on the device (`bm.groups[g].obs`) unless `fetch_obs()` is called.
    """

    def __init__(self, bm, normalized=True, discrete=False, obs=True, use_graph=False):
        self.bm = bm
        self._use_graph, self._graph = use_graph, None
        dt = torch.int32 if discrete else bm.action_dtype
        item = 4 if discrete else dt.itemsize
        sizes = [g.n_envs * (1 if discrete else g.n_act) for g in bm.groups]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n * item + 15) // 16 * 16 // item      # keep every group's block 16-byte aligned
        self._h_act = torch.empty(total, dtype=dt).pin_memory()
        self._d_act = torch.empty(total, dtype=dt, device=bm.device)
        shape = (lambda g: (g.n_envs,)) if discrete else (lambda g: (g.n_envs, g.n_act))
        self.actions = [self._h_act[o:o + n].view(shape(g)) for o, n, g in zip(offs, sizes, bm.groups)]
        d_views = [self._d_act[o:o + n].view(shape(g)) for o, n, g in zip(offs, sizes, bm.groups)]
        self._h_out = torch.empty(bm.n_envs * 9, dtype=torch.uint8).pin_memory()
        self.reward = self._h_out[:bm.n_envs * 8].view(torch.float64)
        self.done = self._h_out[bm.n_envs * 8:]
        self._launch = bm.prepare_step(d_views if len(d_views) > 1 else d_views[0], normalized=normalized, obs=obs,
                                       discrete=discrete)
        self.h2d_bytes = self._h_act.numel() * item
        self.d2h_bytes = self._h_out.numel()

    def _body(self):
        self._d_act.copy_(self._h_act, non_blocking=True)
        self._launch()
        self._h_out.copy_(self.bm._out, non_blocking=True)

    def step(self):
        """One H2D copy, the fused kernel, one D2H copy.  With use_graph=True the three are replayed from a CUDA graph
        (captured on first use; capture executes nothing) -- measured no faster on B200 (the PCIe copies dominate), so
        the default is the plain sequence."""
        if not self._use_graph:
            return self._body()
        if self._graph is None:
            torch.cuda.current_stream(self.bm.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self._graph = g
        self._graph.replay()

    def sync(self):
        torch.cuda.current_stream(self.bm.device).synchronize()

    def fetch_obs(self):
        return [g.obs.cpu() for g in self.bm.groups]


class HostRollout:
    """A rollout whose actions and results live in HOST memory (BASELINE configs[2]: a year of pre-generated actions),
    pipelined over three CUDA streams so that PCIe, not the sum of copy + kernel + copy, sets the pace:

        copy-in stream :  H2D chunk c+1 ........ H2D chunk c+2 ........
        current stream :  mg_rollout(chunk c) .. mg_rollout(chunk c+1)      (persistent kernel, `chunk` steps per launch)
        copy-out stream:  D2H reward/done c-1 .. D2H reward/done c ....

    `actions[g]`: pinned [n_steps, n_g, n_act] of `bm.action_dtype` (int32 [n_steps, n_g] when discrete), filled by the caller;
    `reward[g]` / `done[g]`: pinned [n_steps, n_g] float64 / uint8, valid after `sync()`.  Per step the same bytes cross
    the bus as in `HostIO.step()` (all actions in, reward + done out); observations go to a device ring of `ring`
    buffers (`obs_ring[g]`; every chunk restarts at slot 0: step s of a chunk writes slot s % ring), where a policy or a
    logger would read them.  Device staging is double-buffered (2 x chunk steps of actions and results), whatever
    n_steps is.

    pipeline="native" (default): ONE C-ABI call, `mg_rollout_host`, which takes the host pointers and owns streams,
    events and staging (include/pymgrid_b200.h) -- the entry point a binding without torch would use.
    pipeline="torch": the same schedule built here from torch streams / events around `mg_rollout` (kept as the
    cross-check of the native one; `run(n)` then needs n to end on a bound chunk length).
    In both, every cross-stream wait refers to an event recorded earlier by the same host thread: no deadlock possible.
    """

    def __init__(self, bm, n_steps, chunk=64, normalized=True, discrete=False, ring=4, keep_obs=True, pipeline="native"):
        if n_steps < 1 or chunk < 1 or ring < 1:
            raise ValueError("n_steps, chunk and ring must be positive")
        if pipeline not in ("native", "torch"):
            raise ValueError("pipeline must be 'native' or 'torch'")
        self.bm, self.n_steps, self.chunk, self.ring = bm, int(n_steps), int(min(chunk, n_steps)), int(ring)
        self.discrete, self.normalized, self.pipeline = bool(discrete), bool(normalized), pipeline
        dev, C = bm.device, self.chunk
        adt = torch.int32 if discrete else bm.action_dtype
        ashape = (lambda g: (g.n_envs,)) if discrete else (lambda g: (g.n_envs, g.n_act))
        self.actions = [torch.empty((self.n_steps,) + ashape(g), dtype=adt, pin_memory=True) for g in bm.groups]
        self.reward = [torch.empty((self.n_steps, g.n_envs), dtype=torch.float64, pin_memory=True) for g in bm.groups]
        self.done = [torch.empty((self.n_steps, g.n_envs), dtype=torch.uint8, pin_memory=True) for g in bm.groups]
        self.obs_ring = [torch.empty((ring, g.n_envs, g.obs_dim), dtype=bm.obs_dtype, device=dev) if keep_obs else None
                         for g in bm.groups]
        item = 4 if discrete else adt.itemsize
        self.h2d_bytes_per_step = sum(a[0].numel() * item for a in self.actions)
        self.d2h_bytes_per_step = 9 * bm.n_envs
        self._launch0 = bm.launch_count
        if pipeline == "native":
            self._io = (MgHostRolloutIO * len(bm.groups))()
            for gi, g in enumerate(bm.groups):
                io = self._io[gi]
                if discrete:
                    io.dactions = self.actions[gi].data_ptr()
                else:
                    io.actions = self.actions[gi].data_ptr()
                io.reward, io.done = self.reward[gi].data_ptr(), self.done[gi].data_ptr()
                io.obs_ring, io.flags = _ptr(self.obs_ring[gi]), _ptr(g.flags)
            return
        self._d_act = [[torch.empty((C,) + ashape(g), dtype=adt, device=dev) for g in bm.groups] for _ in range(2)]
        self._d_rew = [[torch.empty((C, g.n_envs), dtype=torch.float64, device=dev) for g in bm.groups] for _ in range(2)]
        self._d_done = [[torch.empty((C, g.n_envs), dtype=torch.uint8, device=dev) for g in bm.groups] for _ in range(2)]
        self._s_in, self._s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._launch = {}
        lengths = {C} | ({self.n_steps % C} - {0})
        for slot in range(2):
            for n in lengths:       # bind the argument blocks once: a chunk costs one C call
                out = [dict(reward=self._d_rew[slot][gi][:n], done=self._d_done[slot][gi][:n], obs_ring=self.obs_ring[gi],
                            reward_sum=None) for gi in range(len(bm.groups))]
                acts = [a[:n] for a in self._d_act[slot]]
                self._launch[slot, n] = bm.rollout(acts if len(acts) > 1 else acts[0], normalized=normalized,
                                                   discrete=discrete, ring=ring, keep_obs=keep_obs, out=out, bind_only=True)

    @property
    def launches(self):
        """kernel launches enqueued since this object was built (one per chunk)"""
        return self.bm.launch_count - self._launch0

    def run(self, n_steps=None):
        """Enqueue the first n_steps (default: all) of the rollout, asynchronously w.r.t. the host; results are complete
        after `sync()` or after any later work on the current stream, which is made to wait for the last copy-out."""
        T = self.n_steps if n_steps is None else int(n_steps)
        if not 1 <= T <= self.n_steps:
            raise ValueError(f"n_steps must be in [1, {self.n_steps}]")
        if self.pipeline == "native":
            bm = self.bm
            bm._mark_stepped()
            rc = bm._lib.mg_rollout_host(bm._handle, self._io, T, self.chunk, self.ring, int(self.discrete),
                                         int(self.normalized), bm._stream())
            _cabi.check(rc, "mg_rollout_host")
            return
        C, G = self.chunk, range(len(self.bm.groups))
        if (0, T % C or C) not in self._launch:
            raise ValueError(f"run({T}): a last chunk of {T % C} steps was not bound; use a multiple of chunk={C} or n_steps={self.n_steps}")
        cur = torch.cuda.current_stream(self.bm.device)
        s_in, s_out = self._s_in, self._s_out
        s_in.wait_stream(cur)              # the caller's earlier work (state loads, a previous run) comes first
        s_out.wait_stream(cur)
        computed, drained = [None, None], [None, None]
        for c, s0 in enumerate(range(0, T, C)):
            slot, n = c & 1, min(C, T - s0)
            with torch.cuda.stream(s_in):
                if computed[slot] is not None:
                    s_in.wait_event(computed[slot])        # the kernel that read this slot's actions is done
                for gi in G:
                    self._d_act[slot][gi][:n].copy_(self.actions[gi][s0:s0 + n], non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(s_in)
            cur.wait_event(uploaded)
            if drained[slot] is not None:
                cur.wait_event(drained[slot])              # this slot's previous results have left the device
            self._launch[slot, n]()
            computed[slot] = torch.cuda.Event()
            computed[slot].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(computed[slot])
                for gi in G:
                    self.reward[gi][s0:s0 + n].copy_(self._d_rew[slot][gi][:n], non_blocking=True)
                    self.done[gi][s0:s0 + n].copy_(self._d_done[slot][gi][:n], non_blocking=True)
                drained[slot] = torch.cuda.Event()
                drained[slot].record(s_out)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)

    def sync(self):
        torch.cuda.current_stream(self.bm.device).synchronize()

    def __del__(self):
        # the native pipeline copies from / into the pinned buffers behind torch's back: drain before they are released
        try:
            torch.cuda.synchronize(self.bm.device)
        except Exception:
            pass


class LogRecorder:
    """Opt-in log for a SUBSET of a batch (reference: Microgrid.get_log, microgrid.py:434-475; the full log is 169-176 f64
    columns per env-step, ~800 GB per year at 65 536 envs, so it cannot be always-on -- SURVEY.md section 5).

        rec = bm.recorder([3, 17, 4242])
        for ...:  rec.step(actions)          # instead of bm.step(actions); same return value
        df = rec.get_log(17)                 # the reference's DataFrame for that env

    Per step it gathers the selected envs' pre-step state and their info / reward rows on the device and appends them to
    host lists (one small device->host copy per group per step); the DataFrame is built with views.log_row."""

    def __init__(self, bm, env_ids):
        if any(g.info is None for g in bm.groups):
            raise ValueError("LogRecorder needs an engine built with with_info=True")
        if bm.configs is None:
            raise ValueError("LogRecorder needs per-config MicrogridParams (not available for array-form batches)")
        self.bm = bm
        self.env_ids = [int(e) for e in env_ids]
        self._sel = []
        for gi, g in enumerate(bm.groups):
            mine = [(e, int(bm.env_slot[e])) for e in self.env_ids if bm.env_group[e] == gi]
            slots = torch.tensor([s for _, s in mine], dtype=torch.long, device=bm.device)
            self._sel.append(([e for e, _ in mine], slots))
        self.rows = {e: [] for e in self.env_ids}
        self.first_step = {}

    def _snapshot(self):
        out = []
        for (envs, slots), g in zip(self._sel, self.bm.groups):
            gen = g.genset[slots].cpu().numpy() if g.genset is not None else None
            out.append((g.step[slots].cpu().numpy(), g.charge[slots].cpu().numpy(), gen))
        return out

    def step(self, actions, normalized=True, discrete=False, obs=True):
        from . import views
        pre = self._snapshot()
        pristine = self.bm._soc_pristine       # first update of every battery: the logged soc is the constructed-with one
        res = self.bm.step_discrete(actions, obs=obs) if discrete else self.bm.step(actions, normalized=normalized, obs=obs)
        post = self._snapshot()
        for gi, ((envs, slots), g) in enumerate(zip(self._sel, self.bm.groups)):
            if not envs:
                continue
            info, reward = g.info[slots].cpu().numpy(), g.reward[slots].cpu().numpy()
            for k, e in enumerate(envs):
                p = self.bm.configs[self.bm.env_config[e]]
                unpack = lambda w: (int(w) & 0xff, (int(w) >> 8) & 0xff, (int(w) >> 16) & 0xff, (int(w) >> 24) & 0xff)  # noqa: E731
                gen_pre = unpack(pre[gi][2][k]) if pre[gi][2] is not None else (0, 0, 0, 0)
                gen_post = unpack(post[gi][2][k]) if post[gi][2] is not None else (0, 0, 0, 0)
                t = int(pre[gi][0][k])
                self.first_step.setdefault(e, t)
                state = views.state_dict(p, t, float(pre[gi][1][k]), gen_pre, p.battery.soc if pristine else None)
                self.rows[e].append(views.log_row(p, state, info[k], float(reward[k]), gen_post))
        return res

    def get_log(self, env_id):
        import pandas as pd
        rows = self.rows[int(env_id)]
        if not rows:
            return pd.DataFrame()
        cols = pd.MultiIndex.from_tuples(list(rows[0].keys()), names=["module_name", "module_number", "field"])
        start = self.first_step[int(env_id)]
        return pd.DataFrame([list(r.values()) for r in rows], columns=cols, index=pd.RangeIndex(start, start + len(rows)))

    def flush(self):
        self.rows = {e: [] for e in self.env_ids}
        self.first_step = {}


class RolloutLog:
    """The reference-format log of the envs selected with `rollout(..., log_envs=ids)`: the persistent kernel wrote one
    record per selected env and step into a device buffer ([n_logged, n_steps, MG_N_LOG] per group: state before the step,
    genset status after it, reward, done, flags, info columns -- include/pymgrid_b200.h, MG_LOG_*); `get_log(env_id)`
    brings that env's records to the host in ONE copy and builds the DataFrame `Microgrid.get_log()` returns
    (microgrid/microgrid.py:434-475: same MultiIndex columns, same values, index = the steps taken)."""

    def __init__(self, bm, env_ids, buffers, rows, n_steps, pristine):
        self.bm, self.env_ids, self.n_steps = bm, [int(e) for e in env_ids], n_steps
        self._buffers, self._rows, self._pristine = buffers, rows, pristine      # per group tensors; env id -> (group, row)

    def records(self, env_id):
        """[n_steps, MG_N_LOG] numpy array of one env (columns MG_LOG_*)"""
        gi, row = self._rows[int(env_id)]
        return self._buffers[gi][row].cpu().numpy()

    def get_log(self, env_id, drop_singleton_key=False):
        from . import views
        if self.bm.configs is None:
            raise ValueError("get_log needs per-config MicrogridParams (array-form batches: use records())")
        rec = self.records(env_id)
        p = self.bm.configs[self.bm.env_config[int(env_id)]]
        unpack = lambda w: (int(w) & 0xff, (int(w) >> 8) & 0xff, (int(w) >> 16) & 0xff, (int(w) >> 24) & 0xff)   # noqa: E731
        rows = []
        for k in range(len(rec)):
            r = rec[k]
            if int(r[_cabi.MG_LOG_FLAGS]) & ((1 << 5) | (1 << 6)):      # a rejected step logs nothing (the reference raises there)
                break
            state = views.state_dict(p, int(r[_cabi.MG_LOG_STEP]), float(r[_cabi.MG_LOG_CHARGE]), unpack(r[_cabi.MG_LOG_GENSET_BEFORE]),
                                     p.battery.soc if (self._pristine and k == 0) else None)
            rows.append(views.log_row(p, state, r[_cabi.MG_LOG_INFO:_cabi.MG_LOG_INFO + MG_N_INFO], float(r[_cabi.MG_LOG_REWARD]),
                                      unpack(r[_cabi.MG_LOG_GENSET_AFTER])))
        stop = int(rec[len(rows) - 1][_cabi.MG_LOG_STEP]) + 1 if rows else 0
        return views.log_frame(rows, stop, drop_singleton_key)


class BatchedMicrogrid:
    def __init__(self, configs: Sequence[MicrogridParams], env_config, device=None, obs_order="gym_sorted",
                 with_info=False, with_flags=True, remove_redundant_gensets=True, action_order=None,
                 obs_dtype=torch.float64):
        """configs: distinct parameter sets; env_config[i] = index of env i's parameter set.
        obs_dtype: torch.float64 (the reference's type, bit-exact) or torch.float32 (the same values rounded to nearest:
        half the dominant memory traffic, for consumers that feed fp32 policies; everything else stays f64)."""
        self.configs = list(configs)
        env_config = np.asarray(env_config, dtype=np.int64)
        if env_config.ndim != 1 or len(env_config) == 0 or env_config.min() < 0 or env_config.max() >= len(self.configs):
            raise ValueError("env_config must be a non-empty 1-D array of indices into configs")
        T = len(self.configs[0])
        if any(len(p) != T for p in self.configs):
            raise ValueError("all configs must have time series of the same length")

        # ---- series tables (deduplicated by content) -------------------------------------------------------
        def table(rows, width):
            keys, index, uniq = {}, [], []
            for r in rows:
                if r is None:
                    index.append(0)
                    continue
                k = r.tobytes()
                if k not in keys:
                    keys[k] = len(uniq)
                    uniq.append(r)
                index.append(keys[k])
            arr = np.stack(uniq) if uniq else np.zeros((0, T) + ((width,) if width > 1 else ()))
            return arr, index
        load_np, load_idx = table([p.load_ts for p in self.configs], 1)
        pv_np, pv_idx = table([p.pv_ts for p in self.configs], 1)
        grid_np, grid_idx = table([p.grid.time_series if p.grid is not None else None for p in self.configs], 4)

        # ---- priority lists + config records ---------------------------------------------------------------
        plist_rows, plist_key, cfg_recs, action_tables = [], {}, [], []
        for k, p in enumerate(self.configs):
            pls = priority_lists(p.has_genset, p.has_grid,
                                 p.genset.running_min_production if p.genset is not None else None,
                                 remove_redundant_gensets)
            action_tables.append(pls)
            key = tuple(pls)
            if key not in plist_key:
                plist_key[key] = len(plist_rows)
                for pl in pls:
                    rec = MgPriorityList()
                    for j in range(_cabi.MG_PLIST_WIDTH):
                        rec.module[j], rec.action[j] = (pl[j] if j < len(pl) else (_cabi.MG_MOD_NONE, 0))
                    rec.n_elements = len(pl)
                    plist_rows.append(rec)
            cfg_recs.append(config_record(p, load_idx[k], pv_idx[k], grid_idx[k], plist_key[key], len(pls)))
        cfg_np = np.frombuffer(bytes((MgConfig * len(cfg_recs))(*cfg_recs)), dtype=np.dtype(MgConfig)).copy()
        plist_np = np.frombuffer(bytes((MgPriorityList * len(plist_rows))(*plist_rows)), dtype=np.uint8).copy()
        names = {p.renewable_name for p in self.configs}
        if len(names) != 1:
            raise ValueError("all configs must name the renewable module the same way")
        if obs_order == "gym_sorted":
            # gym.spaces.Dict sorts the module names: a renewable called 'PV' (MicrogridGenerator) comes before 'battery',
            # one called 'pv' / 'renewable' after 'load'; other positions have no row layout in the kernel
            name = next(iter(names))
            if name < "battery":
                obs_order = "gym_sorted_pv_first"
            elif not name > "load":
                raise NotImplementedError(f"renewable module name {name!r} sorts between 'battery' and 'load': use obs_order='container'")
        status = None
        if any(p.grid is not None and p.grid.status is not None for p in self.configs):
            status = [None if p.grid is None else (p.grid.status if p.grid.status is not None else p.grid.time_series[:, 3])
                      for p in self.configs]
        self._setup(cfg_np=cfg_np, plist_np=plist_np, action_tables=action_tables, env_config=env_config,
                    cfg_arch=np.array([p.arch for p in self.configs], dtype=np.int64),
                    cfg_step=np.array([p.current_step for p in self.configs], dtype=np.int32),
                    cfg_charge=np.array([p.battery.current_charge for p in self.configs], dtype=np.float64),
                    cfg_soc=np.array([np.nan if p.battery.soc is None else p.battery.soc for p in self.configs], dtype=np.float64),
                    cfg_genset=np.array([0 if p.genset is None else (p.genset.current_status | (p.genset.goal_status << 8)
                                         | (p.genset.steps_until_up << 16) | (p.genset.steps_until_down << 24))
                                         for p in self.configs], dtype=np.int64).astype(np.int32),
                    load_np=load_np, pv_np=pv_np, grid_np=grid_np, cfg_status=status, device=device, obs_order=obs_order,
                    with_info=with_info, with_flags=with_flags, action_order=action_order, obs_dtype=obs_dtype)
        if any(f.noise_std != 0 for p in self.configs for f in p.forecasters.values()):
            self.set_forecast_noise(seed=0)

    def _setup(self, cfg_np, plist_np, action_tables, env_config, cfg_arch, cfg_step, cfg_charge, cfg_genset, load_np,
               pv_np, grid_np, cfg_status, device, obs_order, with_info, with_flags, action_order,
               obs_dtype=torch.float64, cfg_soc=None):
        """Common construction from array-form inputs (also used by the vectorised generator front end):
        cfg_np structured MgConfig records; cfg_arch [n_cfg, 3]; cfg_* initial state per config; series tables;
        cfg_status: None or per-config 0/1 status rows ([n_cfg, T] array or list with None for grid-less configs);
        cfg_soc: None or the soc each config's battery was constructed with (nan = derive from the charge)."""
        if not torch.cuda.is_available():
            raise EngineError("BatchedMicrogrid needs a CUDA device: there is no CPU path")
        self._lib = _cabi.lib()
        self._noise = None              # forecast noise: (device records, seed, env id offset) once set_forecast_noise ran
        self._noise_calls = 0
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.obs_order = obs_order
        if obs_dtype not in (torch.float64, torch.float32):
            raise ValueError("obs_dtype must be torch.float64 or torch.float32")
        self.obs_dtype = obs_dtype
        self.n_envs = len(env_config)
        self.env_config = env_config
        self.action_tables = action_tables
        T = load_np.shape[1]
        self.series_len = T
        dev, f64 = self.device, torch.float64
        self.max_horizon = int(cfg_arch[:, 2].max())
        Tp = T + self.max_horizon + 1
        self.load_raw = torch.from_numpy(np.ascontiguousarray(load_np)).to(dev)
        self.pv_raw = torch.from_numpy(np.ascontiguousarray(pv_np)).to(dev)
        self.grid_raw = torch.from_numpy(np.ascontiguousarray(grid_np)).to(dev) if len(grid_np) else None
        self.load_nrm = torch.empty((len(load_np), Tp), dtype=f64, device=dev)
        self.pv_nrm = torch.empty((len(pv_np), Tp), dtype=f64, device=dev)
        self.grid_nrm = torch.empty((len(grid_np), Tp, 4), dtype=f64, device=dev) if len(grid_np) else None
        self.bounds = torch.empty((len(load_np) + len(pv_np) + 4 * len(grid_np), 2), dtype=f64, device=dev)
        self.cfg = torch.from_numpy(cfg_np.view(np.uint8).reshape(-1).copy()).to(dev)
        self.n_cfg = len(cfg_np)
        self.plist = torch.from_numpy(np.ascontiguousarray(plist_np)).to(dev)
        self._scaled = bool(cfg_np["series_scaled"].any())
        scaled_or_status = self._scaled or cfg_status is not None

        # ---- architecture groups ----------------------------------------------------------------------------
        env_arch_rows = cfg_arch[env_config]
        uniq, first, inverse = np.unique(env_arch_rows, axis=0, return_index=True, return_inverse=True)
        order_of_first = np.argsort(first)                       # groups in order of first appearance
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[order_of_first] = np.arange(len(uniq))
        env_arch = rank[inverse.reshape(-1)]
        order = [tuple(int(x) for x in uniq[i]) for i in order_of_first]
        if len(order) > MG_MAX_GROUPS:
            raise ValueError(f"more than {MG_MAX_GROUPS} architecture groups")
        self.env_group = env_arch
        self.env_slot = np.empty(self.n_envs, dtype=np.int64)
        self.groups: List[Group] = []
        # group-major flat outputs; reward and done share one allocation so one D2H copy returns both
        self._out = torch.zeros(self.n_envs * 9, dtype=torch.uint8, device=dev)
        self.reward = self._out[:self.n_envs * 8].view(f64)
        self.done = self._out[self.n_envs * 8:]
        start = 0
        n_actions_cfg = None if isinstance(action_tables, dict) else np.array([len(t) for t in action_tables])
        for gi, arch in enumerate(order):
            ids = np.nonzero(env_arch == gi)[0]
            # envs of one config sit next to each other so that a 64-env tile shares its time-series windows, and configs
            # that share a grid table (the four tariff x CO2 tables of MicrogridGenerator grids) are neighbours too: the
            # emitters fetch a window once per run of rows that read it at the same step
            ids = ids[np.lexsort((env_config[ids], cfg_np["grid_series"][env_config[ids]]))]
            self.env_slot[ids] = np.arange(len(ids))
            has_genset, has_grid, H = arch
            n_act = 1 + has_grid + 2 * has_genset
            obs_dim = (1 + H) * (2 + 4 * has_grid) + 2 + 4 * has_genset
            if scaled_or_status and (obs_dim > 192 or H % 2 == 0):
                raise ValueError("profile-times-scale series / per-env grid status need an odd forecast horizon and obs_dim <= 192")
            names = [m for m, present in (("genset", has_genset), ("battery", 1), ("grid", has_grid)) if present]
            if action_order is None:
                act_names = sorted(names) if obs_order.startswith("gym_sorted") else names
            else:
                act_names = [m for m in action_order if m in names]
            cols, col = {}, 0
            for m in act_names:
                cols[m] = col
                col += 2 if m == "genset" else 1
            cfg_ids = env_config[ids]
            g = Group(arch=arch, env_ids=ids, n_act=n_act, obs_dim=obs_dim, act_cols=cols)
            g.cfg_index = torch.from_numpy(cfg_ids.astype(np.int32)).to(dev)
            g.step = torch.from_numpy(cfg_step[cfg_ids].astype(np.int32)).to(dev)
            g.charge = torch.from_numpy(cfg_charge[cfg_ids].astype(np.float64)).to(dev)
            if cfg_soc is not None:
                # BatteryModule reports the soc it was constructed with until its first update (battery_module.py:89, 125-130);
                # only batches where that differs from charge / max_capacity (by an ulp) carry the extra array
                derived = cfg_charge[cfg_ids].astype(np.float64) / cfg_np["bat_max_capacity"][cfg_ids]
                soc = np.where(np.isnan(cfg_soc[cfg_ids]), derived, cfg_soc[cfg_ids])
                if (soc != derived).any():
                    g.soc_reported = torch.from_numpy(np.ascontiguousarray(soc, dtype=np.float64)).to(dev)
            if has_genset:
                g.genset = torch.from_numpy(cfg_genset[cfg_ids].astype(np.int32)).to(dev)
            if has_grid and cfg_status is not None:
                W = (Tp + 31) // 32
                if isinstance(cfg_status, np.ndarray):
                    rows = cfg_status[cfg_ids]
                else:
                    rows = np.stack([cfg_status[c] for c in cfg_ids])
                bits = np.zeros((len(ids), W * 32), dtype=np.uint8)
                bits[:, :T] = rows.astype(np.uint8)
                packed = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)
                g.status_bits = torch.from_numpy(packed.view(np.int32).copy()).to(dev)
                g.status_words = W
            g.obs = torch.empty((len(ids), obs_dim), dtype=obs_dtype, device=dev)
            g.reward = self.reward[start:start + len(ids)]
            g.done = self.done[start:start + len(ids)]
            g.info = torch.zeros((len(ids), MG_N_INFO), dtype=f64, device=dev) if with_info else None
            g.flags = torch.zeros(len(ids), dtype=torch.int32, device=dev) if with_flags else None
            g.n_actions = (int(n_actions_cfg[np.unique(cfg_ids)].max()) if n_actions_cfg is not None
                           else len(action_tables[(int(has_genset), int(has_grid))]))
            self.groups.append(g)
            start += len(ids)
        self._handle = None
        self._soc_pristine = True      # no battery has updated yet (see _mark_stepped)
        self._create()

    # ------------------------------------------------------------------------------------------------------
    def _layout(self):
        L = MgLayout()
        L.abi_version, L.n_groups = MG_ABI_VERSION, len(self.groups)
        for gi, g in enumerate(self.groups):
            m = L.groups[gi]
            m.has_genset, m.has_grid, m.horizon = g.arch
            m.obs_order = OBS_ORDERS[self.obs_order]
            m.n_act, m.obs_dim, m.n_envs = g.n_act, g.obs_dim, g.n_envs
            m.act_col_genset = g.act_cols.get("genset", 0)
            m.act_col_battery = g.act_cols["battery"]
            m.act_col_grid = g.act_cols.get("grid", 0)
            m.step, m.charge, m.genset, m.cfg_index = _ptr(g.step), _ptr(g.charge), _ptr(g.genset), _ptr(g.cfg_index)
            m.env_initial_step, m.env_final_step = _ptr(g.env_initial_step), _ptr(g.env_final_step)
            m.grid_status_bits, m.status_words = _ptr(g.status_bits), g.status_words
        L.n_cfg, L.series_len, L.max_horizon = self.n_cfg, self.series_len, self.max_horizon
        L.n_load, L.n_pv = self.load_raw.shape[0], self.pv_raw.shape[0]
        L.n_grid = 0 if self.grid_raw is None else self.grid_raw.shape[0]
        L.cfg, L.load_raw, L.pv_raw, L.grid_raw = _ptr(self.cfg), _ptr(self.load_raw), _ptr(self.pv_raw), _ptr(self.grid_raw)
        L.load_nrm, L.pv_nrm, L.grid_nrm, L.bounds = _ptr(self.load_nrm), _ptr(self.pv_nrm), _ptr(self.grid_nrm), _ptr(self.bounds)
        L.plist, L.n_plist = _ptr(self.plist), self.plist.numel() // C.sizeof(MgPriorityList)
        L.flags = (1 if self._scaled else 0) | (2 if self.obs_dtype == torch.float32 else 0)
        return L

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _create(self):
        if self._handle is not None:
            self._lib.mg_destroy(self._handle)
            self._handle = None
        L = self._layout()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.mg_create(C.byref(L), self._stream(), C.byref(h)), "mg_create")
        self._handle = h
        if self._soc_pristine and any(g.soc_reported is not None for g in self.groups):
            socs = (C.c_void_p * len(self.groups))(*[_ptr(g.soc_reported) or None for g in self.groups])
            _cabi.check(self._lib.mg_set_reported_soc(h, socs), "mg_set_reported_soc")
        # tuning options survive a re-creation of the handle
        for option, value in getattr(self, "_options", {}).items():
            _cabi.check(self._lib.mg_set_option(self._handle, option, value), "mg_set_option")

    def _mark_stepped(self):
        """Called by everything that steps (or binds a stepping launcher): the handle drops the constructed-with soc values at
        its first step; this flag keeps a later re-creation of the handle (set_trajectories) from installing them again."""
        self._soc_pristine = False

    def set_forecast_noise(self, seed=0, env_offset=0, records=None):
        """Turn on the Gaussian-noise forecasters of the configs (`MicrogridParams.forecasters`; reference:
        GaussianNoiseForecaster, forecast/forecaster.py:220-262): after every step / reset / observe the forecast
        entries of the freshly written observation rows get N(0, std_k) added and are clipped to their bounds
        (`mg_forecast_noise`).  The draw is a pure function of (seed, call number, env, the env's step, element);
        `env_offset` shifts the env ids so that shards of one batch on different GPUs draw differently.
        `records`: optional explicit list of MgForecastNoise, one per config (array-form front ends).
        `rollout` applies the same noise to the slots of its observation ring once the persistent kernel has written them
        (mg_forecast_noise_at, one call number per step): the ring holds what step-by-step stepping would have produced."""
        if records is None:
            records = [forecast_noise_record(p) for p in self.configs]
        raw = np.frombuffer(bytes((MgForecastNoise * len(records))(*records)), dtype=np.uint8).copy()
        dev_records = torch.from_numpy(raw).to(self.device)
        base, bases = int(env_offset), []
        for g in self.groups:
            bases.append(base)
            base += g.n_envs
        self._noise = (dev_records, int(seed) & (2 ** 64 - 1), (C.c_int64 * len(bases))(*bases))

    def clear_forecast_noise(self):
        self._noise = None

    def _apply_noise(self, obs_bufs):
        """enqueue mg_forecast_noise on the observation buffers a step / reset / observe call has just written"""
        if self._noise is None or all(o is None for o in obs_bufs):
            return
        records, seed, bases = self._noise
        ptrs = (C.c_void_p * len(obs_bufs))(*[_ptr(o) for o in obs_bufs])
        self._noise_calls += 1
        _cabi.check(self._lib.mg_forecast_noise(self._handle, records.data_ptr(), ptrs, bases, seed, self._noise_calls,
                                                self._stream()), "mg_forecast_noise")

    def _set_option(self, option, value):
        if not hasattr(self, "_options"):
            self._options = {}
        self._options[option] = int(value)
        _cabi.check(self._lib.mg_set_option(self._handle, option, int(value)), "mg_set_option")

    def set_rollout_specialised(self, on):
        """Owner / emitter warp-specialised persistent kernel for `rollout` (default on): two warps run the physics one step
        ahead while the other two emit the observation rows."""
        self._set_option(_cabi.MG_OPT_ROLLOUT_SPECIALISED, bool(on))

    def set_rollout_ring(self, on):
        """Batches with per-env series (MicrogridGenerator grids): keep every env's normalised load / pv windows in shared
        memory across the steps of `rollout` (default on; MG_OPT_ROLLOUT_RING).  Off = normalise whole windows per row."""
        self._set_option(_cabi.MG_OPT_ROLLOUT_RING, bool(on))

    def set_emit_image(self, on):
        """How observation rows leave the SM (MG_OPT_EMIT_IMAGE): True = shared-memory images + TMA bulk stores, False = per-lane
        16-byte stores with run detection, "auto" (default) = the library chooses per launch (include/pymgrid_b200.h)."""
        self._set_option(_cabi.MG_OPT_EMIT_IMAGE, 2 if on == "auto" else int(bool(on)))

    def set_image_shape(self, index):
        """Which instantiated (rows per bulk store, buffers per warp, rows gathered together) shape the image kernels use
        (MG_OPT_IMAGE_SHAPE; -1 = the library's choice)."""
        self._set_option(_cabi.MG_OPT_IMAGE_SHAPE, int(index))

    def set_step_overlap(self, mode):
        """Chain consecutive single-step launches with programmatic dependent launch (MG_OPT_STEP_OVERLAP): 0 off, 1 the next
        launch's physics runs under this launch's observation stream, 2 the observation streams overlap as well (needs
        rotating observation buffers: three or more)."""
        self._set_option(_cabi.MG_OPT_STEP_OVERLAP, int(mode))

    @property
    def action_dtype(self):
        return torch.float32 if getattr(self, "_options", {}).get(_cabi.MG_OPT_ACTIONS_F32) else torch.float64

    def set_action_dtype(self, dtype):
        """The element type of every continuous `actions` tensor given to step / rollout / HostIO / HostRollout from now on:
        torch.float64 (default, the reference's) or torch.float32 -- a policy network's output as it is, and half the bytes
        over PCIe for host-resident actions.  float32 values are widened exactly; the arithmetic stays f64, so the step is
        the reference's step for `np.float64(action)` (MG_OPT_ACTIONS_F32).  Launchers bound earlier (`prepare_step`,
        HostIO, HostRollout) keep the tensors they were built with: build them after this call."""
        if dtype not in (torch.float64, torch.float32):
            raise ValueError("action dtype must be torch.float64 or torch.float32")
        self._set_option(_cabi.MG_OPT_ACTIONS_F32, dtype == torch.float32)

    def set_ragged(self, on=True):
        """Tell the library that the envs of a tile are at unrelated steps (independent resets, per-env episode windows):
        no two rows share a window, and the image emitters are the faster ones (MG_OPT_RAGGED_HINT).  Set automatically by
        `set_trajectories` and by masked resets."""
        self._set_option(_cabi.MG_OPT_RAGGED_HINT, bool(on))

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                self._lib.mg_destroy(self._handle)
        except Exception:
            pass

    def set_trajectories(self, initial_step, final_step):
        """Per-env episode windows (reference: microgrid/trajectory/*, Microgrid._set_trajectory microgrid.py:221-225).
        The handle is kept (mg_set_trajectories swaps the window arrays in place): launchers bound earlier by
        `prepare_step` / `prepare_rollout` / `host_io` / `host_rollout` stay valid and see the new windows."""
        initial_step = np.asarray(initial_step, dtype=np.int32)
        final_step = np.asarray(final_step, dtype=np.int32)
        previous = [(g.env_initial_step, g.env_final_step) for g in self.groups]
        for g in self.groups:
            g.env_initial_step = torch.from_numpy(np.ascontiguousarray(initial_step[g.env_ids])).to(self.device)
            g.env_final_step = torch.from_numpy(np.ascontiguousarray(final_step[g.env_ids])).to(self.device)
        lo = (C.c_void_p * len(self.groups))(*[_ptr(g.env_initial_step) for g in self.groups])
        hi = (C.c_void_p * len(self.groups))(*[_ptr(g.env_final_step) for g in self.groups])
        _cabi.check(self._lib.mg_set_trajectories(self._handle, lo, hi), "mg_set_trajectories")
        # kernels already enqueued may still read the previous arrays: release them only once the device has caught up
        self._retired_windows = previous
        torch.cuda.current_stream(self.device).synchronize()
        self._retired_windows = None
        self.set_ragged(True)

    # ------------------------------------------------------------------------------------------------------
    @property
    def single_group(self):
        return len(self.groups) == 1

    def _per_group(self, x, name):
        if x is None:
            return [None] * len(self.groups)
        if isinstance(x, torch.Tensor):
            if not self.single_group:
                raise ValueError(f"{name}: pass one tensor per architecture group ({len(self.groups)} groups)")
            return [x]
        if len(x) != len(self.groups):
            raise ValueError(f"{name}: expected {len(self.groups)} per-group tensors")
        return list(x)

    def _io(self, actions=None, dactions=None, obs=True, mask=None, reward_total=None):
        io = (MgStepIO * len(self.groups))()
        acts = self._per_group(actions, "actions")
        dacts = self._per_group(dactions, "actions")
        masks = self._per_group(mask, "mask")
        obs_bufs = [g.obs for g in self.groups] if obs is True else self._per_group(obs, "obs") if obs is not False else [None] * len(self.groups)
        for gi, g in enumerate(self.groups):
            a, d, m, o = acts[gi], dacts[gi], masks[gi], obs_bufs[gi]
            if a is not None:
                if a.dtype != self.action_dtype or a.shape != (g.n_envs, g.n_act) or not a.is_contiguous() or a.device != self.device:
                    raise ValueError(f"group {gi}: actions must be a contiguous {self.action_dtype} [{g.n_envs}, {g.n_act}] tensor on "
                                     f"{self.device}")
            if d is not None:
                if d.dtype != torch.int32 or d.shape != (g.n_envs,) or not d.is_contiguous() or d.device != self.device:
                    raise ValueError(f"group {gi}: discrete actions must be a contiguous int32 [{g.n_envs}] tensor on {self.device}")
            if o is not None and (o.dtype != self.obs_dtype or o.shape != (g.n_envs, g.obs_dim) or not o.is_contiguous()):
                raise ValueError(f"group {gi}: obs buffer must be a contiguous {self.obs_dtype} [{g.n_envs}, {g.obs_dim}] tensor")
            if m is not None and (m.dtype != torch.uint8 or m.shape != (g.n_envs,)):
                raise ValueError(f"group {gi}: mask must be uint8 [{g.n_envs}]")
            io[gi].actions, io[gi].dactions, io[gi].obs = _ptr(a), _ptr(d), _ptr(o)
            io[gi].reward, io[gi].done, io[gi].info, io[gi].flags = _ptr(g.reward), _ptr(g.done), _ptr(g.info), _ptr(g.flags)
            io[gi].mask = _ptr(m)
            if reward_total is not None:      # one f64 accumulator shared by every group: the whole batch's reward
                io[gi].reward_total = _ptr(reward_total)
        return io, obs_bufs

    def _result(self, obs_bufs):
        if self.single_group:
            g = self.groups[0]
            return obs_bufs[0], g.reward, g.done, g.info
        return obs_bufs, [g.reward for g in self.groups], [g.done for g in self.groups], [g.info for g in self.groups]

    def prepare_step(self, actions, normalized=True, obs=True, discrete=False):
        """Bind the argument block of a step ONCE for fixed buffers and return a zero-argument launcher: the per-call
        host cost drops to one ctypes call (used by HostIO and by loops that step the same buffers repeatedly)."""
        io, obs_bufs = self._io(dactions=actions, obs=obs) if discrete else self._io(actions=actions, obs=obs)
        self._mark_stepped()
        lib, handle, norm = self._lib, self._handle, int(bool(normalized))
        stream_of, dev = torch.cuda.current_stream, self.device
        noise = self._apply_noise if self._noise is not None else None   # (a captured graph replays one call number)
        if discrete:
            def launch():
                rc = lib.mg_step_discrete(handle, io, stream_of(dev).cuda_stream)
                if rc:
                    _cabi.check(rc, "mg_step_discrete")
                if noise:
                    noise(obs_bufs)
        else:
            def launch():
                rc = lib.mg_step(handle, io, norm, stream_of(dev).cuda_stream)
                if rc:
                    _cabi.check(rc, "mg_step")
                if noise:
                    noise(obs_bufs)
        launch.keepalive = (io, obs_bufs, actions)
        return launch

    def step(self, actions, normalized=True, obs=True, reward_total=None):
        """Microgrid.run for every env (reference microgrid.py:227-325).  actions: float64 [n, n_act] per group,
        columns per `Group.act_cols`.  Returns (obs, reward, done, info) as device tensors (lists for > 1 group).
        reward_total: optional f64 [1] device tensor that the kernel adds the batch's summed reward to (logging)."""
        io, obs_bufs = self._io(actions=actions, obs=obs, reward_total=reward_total)
        self._mark_stepped()
        _cabi.check(self._lib.mg_step(self._handle, io, int(bool(normalized)), self._stream()), "mg_step")
        self._apply_noise(obs_bufs)
        return self._result(obs_bufs)

    def step_discrete(self, actions, obs=True):
        """DiscreteMicrogridEnv.step for every env (reference envs/discrete/discrete.py:109-143)."""
        io, obs_bufs = self._io(dactions=actions, obs=obs)
        self._mark_stepped()
        _cabi.check(self._lib.mg_step_discrete(self._handle, io, self._stream()), "mg_step_discrete")
        self._apply_noise(obs_bufs)
        return self._result(obs_bufs)

    def reset(self, mask=None, obs=True):
        """Microgrid.reset (reference microgrid.py:205-225): step = initial_step; battery / genset state is kept."""
        io, obs_bufs = self._io(obs=obs, mask=mask)
        if mask is not None and getattr(self, "_options", {}).get(_cabi.MG_OPT_RAGGED_HINT) is None:
            self.set_ragged(True)      # some envs restart while the others go on: the batch leaves lock-step
        _cabi.check(self._lib.mg_reset(self._handle, io, self._stream()), "mg_reset")
        self._apply_noise(obs_bufs)
        return obs_bufs[0] if self.single_group else obs_bufs

    def observe(self, obs=True):
        io, obs_bufs = self._io(obs=obs)
        _cabi.check(self._lib.mg_observe(self._handle, io, self._stream()), "mg_observe")
        self._apply_noise(obs_bufs)
        return obs_bufs[0] if self.single_group else obs_bufs

    def rbc_actions(self):
        """Per group: the discrete action (priority-list index) rule-based control uses for every env
        (reference: RuleBasedControl._get_priority_list, algos/rbc/rbc.py:31-44)."""
        from .priority_list import rbc_priority_list
        idx = [self.action_tables[k].index(rbc_priority_list(p)) for k, p in enumerate(self.configs)]
        idx = np.asarray(idx, dtype=np.int32)
        return [torch.from_numpy(idx[self.env_config[g.env_ids]]).to(self.device) for g in self.groups]

    def rollout_rbc(self, n_steps, **kw):
        """Rule-based control for n_steps on the device: one persistent kernel, the same priority list every step
        (reference: RuleBasedControl.run, algos/rbc/rbc.py:64-93, without its early `break` on done)."""
        acts = self.rbc_actions()
        return self.rollout(acts if len(acts) > 1 else acts[0], discrete=True, constant_actions=True, n_steps=n_steps, **kw)

    def prepare_rollout(self, actions, **kw):
        """Bind a rollout once and return a zero-argument launcher (`launcher.result` holds the output buffers)."""
        return self.rollout(actions, bind_only=True, **kw)

    def rollout(self, actions, normalized=True, discrete=False, ring=1, keep_obs=True, reward_sum=False, out=None,
                constant_actions=False, n_steps=None, reward_total=None, bind_only=False, log_envs=None):
        """n_steps consecutive steps in one persistent kernel.  actions: per group [n_steps, n, n_act] float64
        (or [n_steps, n] int32 when discrete).  Returns dict(reward=[n_steps, n], done=..., obs_ring=[ring, n, D]);
        pass a previous return value (list of dicts) as `out` to reuse its buffers.
        log_envs: global ids of envs whose per-step log the kernel records (reference: Microgrid.get_log); the log comes
        back as `self.last_log` (a RolloutLog: `.get_log(env_id)` is the reference's DataFrame)."""
        if out is not None and isinstance(out, dict):
            out = [out]
        acts = self._per_group(actions, "actions")
        if constant_actions:
            if not discrete or n_steps is None:
                raise ValueError("constant_actions needs discrete=True and n_steps")
        else:
            n_steps = acts[0].shape[0]
        io = (MgRolloutIO * len(self.groups))()
        outs = []
        for gi, g in enumerate(self.groups):
            a = acts[gi]
            io[gi].dactions_const = int(constant_actions)
            want = (g.n_envs,) if constant_actions else (n_steps, g.n_envs) if discrete else (n_steps, g.n_envs, g.n_act)
            if tuple(a.shape) != want or a.dtype != (torch.int32 if discrete else self.action_dtype) or not a.is_contiguous():
                raise ValueError(f"group {gi}: rollout actions must be contiguous {want} of {torch.int32 if discrete else self.action_dtype}")
            r = out[gi] if out is not None else dict(reward=torch.empty((n_steps, g.n_envs), dtype=torch.float64, device=self.device),
                     done=torch.empty((n_steps, g.n_envs), dtype=torch.uint8, device=self.device),
                     obs_ring=torch.empty((ring, g.n_envs, g.obs_dim), dtype=self.obs_dtype, device=self.device) if keep_obs else None,
                     reward_sum=torch.empty(g.n_envs, dtype=torch.float64, device=self.device) if reward_sum else None)
            outs.append(r)
            if discrete:
                io[gi].dactions = _ptr(a)
            else:
                io[gi].actions = _ptr(a)
            io[gi].obs_ring, io[gi].reward, io[gi].done = _ptr(r["obs_ring"]), _ptr(r["reward"]), _ptr(r["done"])
            io[gi].reward_sum, io[gi].flags = _ptr(r["reward_sum"]), _ptr(g.flags)
            io[gi].reward_total = _ptr(reward_total)
        log_keep = None
        if log_envs is not None:
            log_envs = [int(e) for e in log_envs]
            rows, buffers, slots = {}, [], []
            for gi, g in enumerate(self.groups):
                mine = [e for e in log_envs if self.env_group[e] == gi]
                slot = np.full(g.n_envs, -1, dtype=np.int32)
                for r, e in enumerate(mine):
                    slot[int(self.env_slot[e])] = r
                    rows[e] = (gi, r)
                buf = torch.zeros((max(len(mine), 1), n_steps, MG_N_LOG), dtype=torch.float64, device=self.device)
                slot_dev = torch.from_numpy(slot).to(self.device)
                if mine:
                    io[gi].log_slot, io[gi].log = _ptr(slot_dev), _ptr(buf)
                buffers.append(buf)
                slots.append(slot_dev)
            self.last_log = RolloutLog(self, log_envs, buffers, rows, n_steps, self._soc_pristine)
            log_keep = (buffers, slots)
        lib, handle, norm, dev = self._lib, self._handle, int(bool(normalized)), self.device
        self._mark_stepped()

        def launch():
            st = torch.cuda.current_stream(dev).cuda_stream
            rc = (lib.mg_rollout_discrete(handle, io, n_steps, ring, st) if discrete
                  else lib.mg_rollout(handle, io, n_steps, ring, norm, st))
            if rc:
                _cabi.check(rc, "mg_rollout")
        if self._noise is not None and keep_obs:
            plain_launch, engine = launch, self

            def launch():
                # Gaussian-noise forecasts: the persistent kernel writes the oracle rows; the slots that survive in the ring
                # then get the noise of THEIR step (call numbers continue the per-step sequence, so the ring equals what
                # n_steps x (mg_step + mg_forecast_noise) leaves behind)
                base = [g.step.clone() for g in engine.groups]
                plain_launch()
                records, seed, bases = engine._noise
                first_call = engine._noise_calls
                engine._noise_calls += n_steps
                step_base = (C.c_void_p * len(base))(*[_ptr(b) for b in base])
                for s_idx in range(max(0, n_steps - ring), n_steps):
                    ptrs = (C.c_void_p * len(outs))(*[_ptr(r["obs_ring"][s_idx % ring]) for r in outs])
                    _cabi.check(engine._lib.mg_forecast_noise_at(engine._handle, records.data_ptr(), ptrs, bases, seed, first_call + s_idx + 1,
                                                                 step_base, s_idx + 1, engine._stream()), "mg_forecast_noise_at")
                launch.keepalive_noise = base
        result = outs[0] if self.single_group else outs
        if bind_only:      # the caller launches (repeatedly) with minimal host overhead; buffers are kept alive here
            launch.keepalive = (io, acts, outs, reward_total, log_keep)
            launch.result = result
            return launch
        launch.keepalive = (io, acts, outs, reward_total, log_keep)
        launch()
        return result

    def recorder(self, env_ids):
        """Opt-in reference-format log for a subset of envs (see LogRecorder)."""
        return LogRecorder(self, env_ids)

    def host_io(self, normalized=True, discrete=False, obs=True, use_graph=False):
        """Pinned host staging for a host-resident control loop (see HostIO)."""
        return HostIO(self, normalized=normalized, discrete=discrete, obs=obs, use_graph=use_graph)

    def host_rollout(self, n_steps, chunk=64, **kw):
        """Rollout with host-resident actions and results, copies overlapped with the kernel (see HostRollout)."""
        return HostRollout(self, n_steps, chunk=chunk, **kw)

    # ------------------------------------------------------------------------------------------------------
    @property
    def launch_count(self):
        return int(self._lib.mg_launch_count(self._handle))

    @property
    def last_kernel(self):
        """name of the kernel family the last step / rollout launched (which emitters the library chose)"""
        return self._lib.mg_last_kernel(self._handle).decode()

    def state_dict(self):
        """The reference's serialisable state (base_module.py:852-868, genset_module.py:426-427), per group."""
        return [dict(step=g.step.clone(), charge=g.charge.clone(), genset=None if g.genset is None else g.genset.clone())
                for g in self.groups]

    def load_state_dict(self, state):
        # BatteryModule.current_charge setter (battery_module.py:360-362): the soc follows the restored charge
        self._mark_stepped()
        _cabi.check(self._lib.mg_set_reported_soc(self._handle, None), "mg_set_reported_soc")
        for g, s in zip(self.groups, state):
            g.step.copy_(s["step"])
            g.charge.copy_(s["charge"])
            if g.genset is not None:
                g.genset.copy_(s["genset"])

    def genset_status(self, gi=0):
        w = self.groups[gi].genset
        return torch.stack([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, (w >> 24) & 0xff], dim=1)

    def scatter_to_env_order(self, per_group):
        """Concatenate per-group [n_g, ...] tensors back into global env order (host-side convenience)."""
        out = torch.empty((self.n_envs,) + tuple(per_group[0].shape[1:]), dtype=per_group[0].dtype, device=self.device)
        for g, x in zip(self.groups, per_group):
            out[torch.from_numpy(g.env_ids).to(self.device)] = x
        return out

    # ------------------------------------------------------------------------------------------------------
    @classmethod
    def from_pymgrid25(cls, batch, scenarios=None, forecast_horizon=None, **kw):
        """`batch` envs tiled over pymgrid25 scenarios: env i -> scenarios[i % len(scenarios)]."""
        from .scenario import load_pymgrid25
        scenarios = list(range(25)) if scenarios is None else list(scenarios)
        configs = [load_pymgrid25(n) for n in scenarios]
        if forecast_horizon is not None:
            for p in configs:
                p.forecast_horizon = forecast_horizon
        env_config = np.arange(batch) % len(scenarios)
        return cls(configs, env_config, **kw)
