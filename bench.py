#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched microgrid step on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE configs[2] -- 65 536 microgrids per GPU tiled from all 25 pymgrid25
scenarios (env i -> scenario i mod 25; three architecture groups fused in one launch), year-rollout style:
every step reads that step's own pre-generated U[0,1) actions, advances every env one timestep and writes reward,
done, state and the FULL normalised observation.  A "step" is one pass of the hot path over the whole batch.

  value      whole-job env-steps/s with actions resident in HBM (CUDA events, max over ranks).  Observation
             buffers rotate through a ring larger than L2 so the stores reach HBM.
  e2e        the same metric through the public API `BatchedMicrogrid.step` with HOST buffers: per step the
             actions go pinned-host -> device and reward + done come back device -> pinned-host inside the
             timed region (observations stay on the device, where a policy consumes them).
  roofline   HBM: algorithmic bytes per launch (SURVEY.md 8d per-env-step figure x envs) / average launch time
             measured with CUDA events in this run, against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  the C oracle port of the reference step (oracle/mg_oracle.c) on this box's host cores, bounded sample.

`--impl reference` times that CPU port alone (the reference itself is pure Python and cannot travel to the GPU
box; its measured 1e3 steps/s/core is quoted in BASELINE.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 65536
METRIC = "microgrid env-steps/sec at batch 65536 (pymgrid25)"
UNIT = "env-steps/s"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def algorithmic_bytes(has_genset, has_grid, horizon, discrete=False, obs_bytes=8):
    """act + state r/w + reward + done + obs, f64 parity mode, per env-step.  +8 for the per-env
    step counter (read + write), which this engine keeps per env so that envs need not run in lock-step."""
    act = 4 if discrete else 8 * (1 + has_grid + 2 * has_genset)
    state = 2 * (8 + 4 * has_genset) + 8
    obs_dim = (1 + horizon) * (2 + 4 * has_grid) + 2 + 4 * has_genset
    return act + state + 8 + 1 + obs_bytes * obs_dim


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region.  Polls NVML in-process (sub-millisecond period, a
    timed region lasts only tens of milliseconds); falls back to the nvidia-smi query of B200_PROFILING.md."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self._halt = index, [], set(), threading.Event()
        self.sm_max, self.source = None, "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv, self.source = None, "nvidia-smi"

    def _poll_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.sm.append(float(out[0]))
        self.sm_max = float(out[1])
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                self._poll_nvml() if self._nv else self._poll_smi()
            except Exception:
                pass
            self._halt.wait(0.0005 if self._nv else 0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


WORKLOADS = {
    "pymgrid25": "configs[2]: 65536 grids/GPU tiled from all 25 pymgrid25 configs, year-rollout style steps with full obs",
    "replicas": "configs[1]: replicas of pymgrid25 microgrid_0, continuous normalised step (default batch 4096)",
    "discrete": "configs[3]: DiscreteMicrogridEnv step (priority-list actions), forecast_horizon=24, the 15 pymgrid25 grids with a GridModule tiled",
    "generator": "configs[4]: heterogeneous MicrogridGenerator grids (profile*scale series, weak-grid outages), one parameter set per env",
}
GRID_SCENARIOS = [0, 4, 6, 11, 12, 14, 16, 1, 8, 9, 10, 13, 18, 22, 24]
WORKLOADS["composed"] = ("beyond BASELINE's configs: the general-dispatch path (mgc_run) on microgrids outside the fused module "
                         "set -- pymgrid25 microgrid_1 with its load, pv and battery each split in two (2 loads, 2 renewables, "
                         "2 batteries, genset, grid; H=23), replicated")


def composed_modules():
    """pymgrid25 microgrid_1 (genset + grid) re-cut into a module list the fused kernels do not cover"""
    from pymgrid_b200 import modules as M
    from pymgrid_b200.scenario import load_pymgrid25
    p = load_pymgrid25(1)
    load, pv, b, g, gr = -p.load_ts, p.pv_ts, p.battery, p.genset, p.grid
    half = dict(min_capacity=b.min_capacity / 2, max_capacity=b.max_capacity / 2, max_charge=b.max_charge / 2,
                max_discharge=b.max_discharge / 2, efficiency=b.efficiency, battery_cost_cycle=b.battery_cost_cycle, init_soc=0.6)
    ts = dict(forecaster="oracle", forecast_horizon=23)
    return [M.LoadModule(0.6 * load, **ts), M.LoadModule(0.4 * load, **ts), M.RenewableModule(0.7 * pv, **ts),
            M.RenewableModule(0.3 * pv, **ts), M.BatteryModule(**half), M.BatteryModule(**half),
            M.GensetModule(g.running_min_production, g.running_max_production, g.genset_cost, g.co2_per_unit, g.cost_per_unit_co2),
            M.GridModule(gr.max_import, gr.max_export, gr.time_series, cost_per_unit_co2=gr.cost_per_unit_co2, **ts)]


def run_composed(args):
    """`--workload composed`: the same contract for the general-dispatch kernel (single GPU or independent shards)."""
    import torch
    from pymgrid_b200.compose import ComposedBatch, Composition
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    B, K, W, R = args.batch or BATCH_PER_GPU, args.steps, args.warmup, args.ring
    comp = Composition(composed_modules(), loss_load_cost=10.0, overgeneration_cost=1.0)
    batch = ComposedBatch([comp], np.zeros(B, dtype=np.int64), device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    Kc = min(K, 512)                                     # steps per launch; a fresh action block per launch
    chunks = [Kc] * (K // Kc) + ([K % Kc] if K % Kc else [])
    actions = torch.rand((Kc, B, comp.n_act), dtype=torch.float64, device=dev, generator=gen)
    state0 = [t.clone() for t in (batch.step_counter, batch.fstate, batch.istate)]

    def restore():
        for t, s0 in zip((batch.step_counter, batch.fstate, batch.istate), state0):
            t.copy_(s0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    batch.rollout(actions[:max(W, 3)], ring=R)           # warm-up steps (also sizes the kernel's local memory)
    outs = {n: batch.rollout(actions[:n], ring=R) for n in set(chunks)}      # untimed: the output buffers of every launch shape
    restore()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launch0 = batch.launch_count
    ev0.record()
    for n in chunks:
        batch.rollout(actions[:n], ring=R, out=outs[n])
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = batch.launch_count - launch0
    value = world * B * K / (ms * 1e-3)
    # e2e: the caller's loop with HOST buffers -- actions pinned-host -> device, reward + done device -> pinned-host, per launch
    Ke = min(K, 256)
    h_act = torch.rand((Ke, B, comp.n_act), dtype=torch.float64).pin_memory()
    h_rew = torch.empty((Ke, B), dtype=torch.float64).pin_memory()
    h_done = torch.empty((Ke, B), dtype=torch.uint8).pin_memory()
    d_act = torch.empty_like(h_act, device=dev)
    out_e = batch.rollout(d_act, ring=R)                 # untimed: output buffers
    restore()
    barrier()
    ev0.record()
    d_act.copy_(h_act, non_blocking=True)
    out = batch.rollout(d_act, ring=R, out=out_e)
    h_rew.copy_(out["reward"], non_blocking=True)
    h_done.copy_(out["done"], non_blocking=True)
    ev1.record()
    barrier()
    ms_e = max_over_ranks(ev0.elapsed_time(ev1))
    if rank == 0:
        n_bat = sum(s.kind == "battery" for s in comp.slots)
        n_gen = sum(s.kind == "genset" for s in comp.slots)
        per_step = 8 * comp.n_act + 2 * (4 + 16 * n_bat + 16 * n_gen) + 9 + 8 * comp.obs_dim      # act + state r/w + reward, done + obs
        peak, peak_src = measured_peak()
        achieved = B * per_step * K / (ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "pymgrid25 microgrid_1 parameters + series (bundled) re-cut into 8 modules, synthetic U[0,1) actions",
            "config": {"workload": WORKLOADS["composed"], "batch_per_gpu": B, "global_batch": world * B, "obs_dim": comp.obs_dim,
                       "n_act": comp.n_act, "path": "mgc_run", "steps_per_launch": Kc,
                       "l2": f"obs ring of {R} buffers = {R * B * comp.obs_dim * 8 / 1e6:.0f} MB, action block = "
                             f"{actions.numel() * 8 / 1e6:.0f} MB per GPU (L2 126 MB)",
                       "parallelism": f"batch sharded over {world} GPU(s), no collective on the step path"},
            "gpu_launches": launches, "clocks": clocks,
            "e2e": {"value": world * B * Ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * comp.n_act * 8,
                    "d2h_bytes_per_step": B * 9, "steps": Ke, "api": "ComposedBatch.rollout with pinned host actions / results",
                    "note": "one H2D copy, one launch, two D2H copies, serialised"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "mgc_kernel", "bytes_per_step": B * per_step,
                         "bytes_per_launch": B * per_step * K / max(launches, 1), "steps_per_launch": K / max(launches, 1)},
        }
        if world == 1 and not args.no_cpu:
            import time
            from oracle.compose import ComposedOracle
            orc = ComposedOracle(composed_modules(), loss_load_cost=10.0, overgeneration_cost=1.0)
            rng = np.random.default_rng(0)
            n, t0 = 1500, time.perf_counter()
            for _ in range(n):
                orc.run({"battery": list(rng.random(2)), "genset": [rng.random(2)], "grid": [rng.random()]})
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"1 env x {n} steps of the same microgrid in {dt:.2f} s (pure-Python oracle of the "
                                              f"general dispatch, oracle/compose.py; obs + log row every step)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def build_engine(batch, device, rank=0, world=1, workload="pymgrid25", obs_f32=False):
    """This rank's contiguous slice of the global batch world*batch (global env numbering, independent of the sharding)."""
    from pymgrid_b200.sharding import shard_range, sharded_pymgrid25
    import torch
    kw = dict(device=device, with_info=False, with_flags=False, obs_dtype=torch.float32 if obs_f32 else torch.float64)
    if workload == "pymgrid25":     # env i -> scenario i mod 25
        return sharded_pymgrid25(world * batch, rank, world, **kw)
    from pymgrid_b200.engine import BatchedMicrogrid
    from pymgrid_b200.scenario import load_pymgrid25
    lo, hi = shard_range(world * batch, rank, world)
    if workload == "replicas":
        return BatchedMicrogrid([load_pymgrid25(0)], np.zeros(hi - lo, dtype=np.int64), **kw)
    if workload == "discrete":
        configs = []
        for n in GRID_SCENARIOS:
            p = load_pymgrid25(n)
            p.forecast_horizon = 24
            configs.append(p)
        return BatchedMicrogrid(configs, np.arange(lo, hi) % len(configs), **kw)
    from pymgrid_b200 import generator
    gb = generator.sample(hi - lo, seed=1000 + rank)        # every rank samples its own shard
    return generator.engine_from_batch(gb, **kw)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


_CPU_CACHE = {}


def cpu_port_rate(n_envs, n_steps, threads, reps=1, seed=7):
    """Time the C oracle port on a bounded sample of the same workload (env i -> scenario i mod 25, U[0,1) actions,
    full normalised observation computed every step): `reps` consecutive rollouts of n_steps (the state carries on;
    reps * n_steps must stay below the 8760-step year).  Returns (env-steps/s, seconds)."""
    from oracle.oracle import OracleBatch
    from pymgrid_b200.scenario import load_pymgrid25
    assert reps * n_steps <= 8759
    if "configs" not in _CPU_CACHE:
        _CPU_CACHE["configs"] = [load_pymgrid25(n) for n in range(25)]
    configs = _CPU_CACHE["configs"]
    plist = [configs[e % 25] for e in range(n_envs)]
    key = (n_steps, n_envs, seed)
    if key not in _CPU_CACHE:
        _CPU_CACHE[key] = np.random.default_rng(seed).random((n_steps, n_envs, 4))
    actions = _CPU_CACHE[key]
    ob = OracleBatch(plist)
    t0 = time.perf_counter()
    for _ in range(reps):
        ob.rollout(actions, normalized=True, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_envs * n_steps * reps / dt, dt


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_envs, n_steps = 16384, 250         # one "step" of this arm = 4 096 000 env-steps of the workload (~1 CPU-second)
    for _ in range(args.warmup):
        cpu_port_rate(n_envs, 25, threads)
    total, total_t = 0, 0.0
    for k in range(args.steps):
        _, dt = cpu_port_rate(n_envs, n_steps, threads)
        total += n_envs * n_steps
        total_t += dt
    value = total / total_t
    sample = f"{n_envs} envs (25 pymgrid25 scenarios tiled) x {n_steps} steps per bench step, full obs every step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "pymgrid25 scenario parameters + series (bundled), synthetic U[0,1) actions",
        "config": {"workload": "pymgrid25 tiled, CPU port of Microgrid.run (C oracle), bounded sample", "batch": n_envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--batch", type=int, default=None, help="envs per GPU (default 65536; replicas 4096; generator 131072)")
    ap.add_argument("--workload", default="pymgrid25", choices=tuple(WORKLOADS), help="default = the BASELINE metric's workload")
    ap.add_argument("--ring", type=int, default=4, help="observation buffers rotated so stores reach HBM")
    ap.add_argument("--path", default="rollout", choices=("rollout", "graph", "eager"),
                    help="headline path: the persistent rollout kernel (BASELINE configs[2] is a year rollout with pre-generated "
                         "actions), one mg_step launch per step replayed from a CUDA graph, or plain launches from Python")
    ap.add_argument("--single-path", action="store_true", help="time only the headline path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--obs-f32", action="store_true",
                    help="NON-CANONICAL secondary mode: observations written as float32 (half the dominant bytes); reported with dtype f64+f32obs")
    ap.add_argument("--ragged", action="store_true",
                    help="start every env at its own random step (envs that reset independently): no two rows of a tile share a window")
    ap.add_argument("--no-ring", action="store_true",
                    help="per-env series batches (generator workload): use the persistent kernel that normalises whole windows per "
                         "row instead of the one that keeps sliding windows in shared memory (A/B)")
    ap.add_argument("--emit", default="image", choices=("image", "lsu"),
                    help="row emitter: shared-memory images + TMA bulk stores (default) or the per-lane 16-byte store emitters (A/B)")
    ap.add_argument("--image-shape", type=int, default=None, help="MG_OPT_IMAGE_SHAPE index (tuning)")
    ap.add_argument("--no-specialised", action="store_true", help="persistent kernel without the owner / emitter warp split (A/B)")
    ap.add_argument("--preheat", type=float, default=0.2, help="seconds of untimed steps before the warm-up (0 under ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "composed":
        return run_composed(args)

    import torch
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if args.batch is None:
        args.batch = {"replicas": 4096, "generator": 131072}.get(args.workload, BATCH_PER_GPU)
    B, K, W, R = args.batch, args.steps, args.warmup, args.ring
    discrete = args.workload == "discrete"

    bm = build_engine(B, dev, rank, world, args.workload, args.obs_f32)
    if args.no_ring:
        bm.set_rollout_ring(False)
    if args.emit == "lsu":
        bm.set_emit_image(False)
    if args.image_shape is not None:
        bm.set_image_shape(args.image_shape)
    if args.no_specialised or (args.emit == "lsu" and args.ragged):
        bm.set_rollout_specialised(False)      # (LSU emitters, envs at unrelated steps: the plain persistent kernel is the faster one)
    groups = bm.groups

    def rand_actions(steps, g):
        if discrete:
            return torch.randint(0, g.n_actions, (steps, g.n_envs), dtype=torch.int32, device=dev, generator=gen)
        return torch.rand((steps, g.n_envs, g.n_act), dtype=torch.float64, device=dev, generator=gen)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    # every step reads its own actions from HBM: a ring of A steps (> 2x L2) reused cyclically
    A = min(W + K, 256)
    acts = [rand_actions(A, g) for g in groups]
    rings = [torch.empty((R, g.n_envs, g.obs_dim), dtype=bm.obs_dtype, device=dev) for g in groups]
    ring_bytes = sum(r.numel() * r.element_size() for r in rings)
    act_bytes = sum(a.numel() * 8 for a in acts)
    if args.ragged:
        for g in groups:
            g.step.copy_(torch.randint(0, 8760 - 2 * (W + K) - 64 if 2 * (W + K) < 4000 else 100, (g.n_envs,), dtype=torch.int32, device=dev, generator=gen))
    state0 = bm.state_dict()
    launchers = {}   # one pre-bound launcher per (action slot, obs slot)

    def one_step(s):
        key = (s % A, s % R)
        if key not in launchers:
            launchers[key] = bm.prepare_step([a[key[0]] for a in acts], obs=[r[key[1]] for r in rings], discrete=discrete)
        launchers[key]()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def time_path(path):
        """W untimed warm-up steps, then EXACTLY K timed steps between barrier + synchronize; returns (ms, launches, clocks)."""
        bm.load_state_dict(state0)
        if path == "rollout":    # persistent kernel: the K steps run in ceil(K / 2048) launches
            Kc = min(K, 2048)
            gen.manual_seed(3 + rank)
            timed = [rand_actions(Kc, g) for g in groups]
            chunks = [Kc] * (K // Kc) + ([K % Kc] if K % Kc else [])
            bm.rollout([a[:max(W, 3)] for a in timed], ring=R, keep_obs=True, discrete=discrete)             # warm-up steps
            # untimed: allocate the outputs and bind the argument blocks, so that a timed launch is one C call
            binds = {n: bm.prepare_rollout([a[:n] for a in timed], ring=R, keep_obs=True, discrete=discrete) for n in set(chunks)}
            bm.load_state_dict(state0)
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launch0 = bm.launch_count
            ev0.record(stream)
            for n in chunks:
                binds[n]()
            ev1.record(stream)
        elif path == "graph":    # one mg_step launch per step, replayed from a CUDA graph
            chunk = K if K <= 256 else max(d for d in range(1, 257) if K % d == 0)
            for s in range(W + chunk):      # make every launcher of the chunk exist before capture
                one_step(s)
            bm.load_state_dict(state0)
            torch.cuda.synchronize()
            launch0 = bm.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for s in range(W, W + chunk):
                    one_step(s)
            captured = bm.launch_count - launch0
            bm.load_state_dict(state0)
            for s in range(W):
                one_step(s)
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launch0 = bm.launch_count - captured * (K // chunk)
            ev0.record(stream)
            for _ in range(K // chunk):
                graph.replay()
            ev1.record(stream)
        else:                    # eager: one mg_step launch per step from Python
            for s in range(W + min(K, A * R)):
                one_step(s)
            bm.load_state_dict(state0)
            for s in range(W):
                one_step(s)
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launch0 = bm.launch_count
            ev0.record(stream)
            for s in range(W, W + K):
                one_step(s)
            ev1.record(stream)
        barrier()
        clocks = sampler.stop()
        return max_over_ranks(ev0.elapsed_time(ev1)), bm.launch_count - launch0, clocks

    with torch.cuda.stream(stream):
        if args.preheat > 0:     # untimed: bring the clocks to their loaded state
            t0 = time.perf_counter()
            s = 0
            while time.perf_counter() - t0 < args.preheat:
                for _ in range(64):
                    one_step(s)
                    s += 1
                stream.synchronize()
        ms, launches, clocks = time_path(args.path)
        others = {}
        if not args.single_path:
            for p in ("rollout", "graph", "eager"):
                if p != args.path:
                    ms_p, l_p, _ = time_path(p)
                    others[p] = {"value": world * B * K / (ms_p * 1e-3), "us_per_step": 1e3 * ms_p / K, "gpu_launches": l_p}
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the public API with host buffers -------------------------------------------------
    # (1) HostIO.step(): one step per call -- actions pinned-host -> device (one copy), fused kernel, reward + done
    #     device -> pinned-host (one copy), serialised on one stream: the figure a host-side control loop sees.
    Ke = min(K, 200)
    bm.load_state_dict(state0)
    hio = bm.host_io(normalized=True, obs=[r[0] for r in rings], discrete=discrete)
    for a, g in zip(hio.actions, groups):          # the caller's actions, in pinned host memory
        a.copy_(torch.randint(0, g.n_actions, tuple(a.shape), dtype=torch.int32) if discrete else torch.rand(tuple(a.shape), dtype=torch.float64))
    with torch.cuda.stream(stream):
        for s in range(3):
            hio.step()
        barrier()
        ev0.record(stream)
        for s in range(Ke):
            hio.step()
        ev1.record(stream)
        barrier()
    h2d, d2h = hio.h2d_bytes, hio.d2h_bytes
    e2e_step_value = world * B * Ke / (max_over_ranks(ev0.elapsed_time(ev1)) * 1e-3)
    e2e = {"value": e2e_step_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
           "api": "BatchedMicrogrid.host_io().step()",
           "note": "actions written into pinned host memory by the caller, one H2D copy, fused kernel, one D2H copy of "
                   "reward+done, every step; observations stay on the device"}
    # (2) HostRollout.run(): the workload's own call (a year rollout with pre-generated actions) with HOST buffers --
    #     every step's actions cross the bus host -> device and every step's reward + done come back, in chunks of
    #     `chunk` steps, the copies of neighbouring chunks overlapped with the persistent kernel on three streams.
    #     Same bytes per step as (1); this is the headline e2e figure when it runs (any failure keeps (1) and says so).
    # Collectives (barrier, max over ranks) stay outside the try blocks so that a failure on one rank cannot hang the others.
    Kr, chunk = min(K, 1024), 64
    del hio
    errors = {}
    for pipeline in ("native", "torch"):       # mg_rollout_host (one C-ABI call); else the same schedule from torch streams
        err, hr, ms_local, launch_e = None, None, float("inf"), 0
        try:
            bm.load_state_dict(state0)
            hr = bm.host_rollout(Kr, chunk=chunk, normalized=True, discrete=discrete, ring=R, pipeline=pipeline)
            for a, g in zip(hr.actions, groups):
                if discrete:
                    a.copy_(torch.randint(0, g.n_actions, tuple(a.shape), dtype=torch.int32))
                else:
                    a.uniform_(0.0, 1.0)
            with torch.cuda.stream(stream):
                hr.run(min(Kr, 3 * chunk) if Kr % chunk == 0 else Kr)      # warm-up (untimed)
                bm.load_state_dict(state0)
        except Exception as ex:      # keep the per-step figure; never lose the bench line to this path
            err = f"{type(ex).__name__}: {ex}"
        barrier()
        if err is None:
            try:
                with torch.cuda.stream(stream):
                    launch_e = bm.launch_count
                    ev0.record(stream)
                    hr.run()
                    ev1.record(stream)
                torch.cuda.synchronize()
                launch_e = bm.launch_count - launch_e
                if not all(bool(torch.isfinite(r).all()) for r in hr.reward):
                    raise RuntimeError("non-finite reward came back from the host rollout")
                ms_local = ev0.elapsed_time(ev1)
            except Exception as ex:
                err = f"{type(ex).__name__}: {ex}"
                ms_local = float("inf")
        barrier()
        ms_e = max_over_ranks(ms_local)
        if ms_e != float("inf"):
            name = "mg_rollout_host" if pipeline == "native" else "mg_rollout + torch streams"
            e2e = {"value": world * B * Kr / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": hr.h2d_bytes_per_step,
                   "d2h_bytes_per_step": hr.d2h_bytes_per_step, "steps": Kr, "chunk_steps": chunk,
                   "gpu_launches": launch_e, "us_per_step": 1e3 * ms_e / Kr,
                   "pcie_gbs": {"h2d": hr.h2d_bytes_per_step * Kr / (ms_e * 1e-3) / 1e9, "d2h": hr.d2h_bytes_per_step * Kr / (ms_e * 1e-3) / 1e9},
                   "api": f"BatchedMicrogrid.host_rollout(n_steps).run() -> {name}",
                   "note": "year-rollout call with HOST buffers: every step's actions go pinned-host -> device and every step's "
                           "reward + done come back device -> pinned-host inside the timed region, in chunks of 64 steps; copy-in, "
                           "persistent kernel and copy-out of neighbouring chunks overlap on three streams (PCIe-bound); "
                           "observations go to the device ring",
                   "per_step_call": {"value": e2e_step_value, "api": "BatchedMicrogrid.host_io().step()", "steps": Ke,
                                     "note": "one H2D + kernel + one D2H per step, serialised (a host-side control loop)"}}
        else:
            errors[pipeline] = err or "failed on another rank"
        hr = None
        if ms_e != float("inf"):
            break
    if errors:
        e2e["host_rollout_errors"] = errors

    if rank == 0:
        bytes_per_launch = sum(g.n_envs * algorithmic_bytes(*g.arch, discrete=discrete, obs_bytes=4 if args.obs_f32 else 8) for g in groups)
        if args.workload == "generator" and args.path != "rollout":
            # + the env's own parameter record and status word(s), re-read by every single-step launch (+~176 B
            # there, 336 B with this engine's record); inside the persistent kernel they stay cache-resident and are not counted
            bytes_per_launch += sum(g.n_envs * (336 + 8 * g.arch[1]) for g in groups)
        peak, peak_src = measured_peak()
        bytes_per_step = bytes_per_launch
        steps_per_launch = K / max(launches, 1)
        bytes_per_launch = bytes_per_step * steps_per_launch
        achieved = bytes_per_launch / (ms * 1e-3 / max(launches, 1)) / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tr = json.load(f)["mg_rollout_kernel" if args.path == "rollout" else "mg_step_kernel"]
            if tr["dram_bytes_per_step"] and B == BATCH_PER_GPU and args.workload == "pymgrid25" and not args.obs_f32 and not args.ragged:
                traffic, traffic_src = tr["dram_bytes_per_step"] * steps_per_launch, tr["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if not args.obs_f32 else "f64 arithmetic, f32 observation output (non-canonical)",
            "data": "pymgrid25 scenario parameters + series (bundled), synthetic U[0,1) actions",
            "config": {"workload": WORKLOADS[args.workload],
                       "batch_per_gpu": B, "global_batch": world * B, "forecast_horizon": 23, "path": args.path, "ragged_steps": bool(args.ragged), "emit": args.emit, "image_shape": args.image_shape, "specialised": not args.no_specialised,
                       "l2": f"inputs larger than L2: obs ring of {R} buffers = {ring_bytes / 1e6:.0f} MB and action ring = {act_bytes / 1e6:.0f} MB per GPU (L2 126 MB)",
                       "parallelism": f"batch sharded over {world} GPU(s), no collective on the step path"},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "mg_step_kernel" if args.path != "rollout" else "mg_rollout_kernel",
                         "bytes_per_launch": bytes_per_launch, "bytes_per_step": bytes_per_step, "steps_per_launch": steps_per_launch,
                         "write_only_ceiling_gbs": 5450.0,
                         "note": "peak is the read+write copy bandwidth; this path is ~97% stores, plain 16-byte stores measured 5.45 TB/s on this GPU (tools/microbench.py)"},
            "other_paths": others,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n_envs, n_steps, reps = 16384, 250, 20          # 81.9e6 env-steps: ~20 s of CPU work at ~4e6 steps/s/core
            cpu_port_rate(n_envs, 25, threads)
            rate, dt = cpu_port_rate(n_envs, n_steps, threads, reps=reps)
            rate1, dt1 = cpu_port_rate(1024, 250, 1, reps=8)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_envs} envs x {n_steps * reps} steps of the same workload = {n_envs * n_steps * reps / 1e6:.1f}e6 "
                                              f"env-steps in {dt:.2f} s wall on {threads} threads (C oracle port of Microgrid.run, full obs every step)",
                                    "single_core": rate1,
                                    "python_reference_note": "the unmodified Python reference measures ~1e3 env-steps/s/core (BASELINE.md); it cannot travel to this box"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
