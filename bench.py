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
  cpu_baseline  the C oracle port of the reference step (oracle/mg_oracle.c) on this box's host cores, bounded sample;
             `python_reference` beside it is the unmodified Python reference timed in the build container
             (tools/time_python_reference.py -> profiles/python_reference_timing.json; it cannot travel to the GPU box).
  configs    short in-run measurements of BASELINE configs[1], [3] and the per-GPU shard of [4] (value, us/step, roofline
             fraction), so that the driver-run line carries every config; under torchrun every rank runs its shard.

A timed region that would last less than MIN_TIMED_MS is repeated (state restored and a fresh action block between
repetitions, both untimed) and the MEDIAN is reported, with `repeats` in the line: `--steps 20` is 0.2 ms of GPU work.

`--impl reference` times the CPU port alone.
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
MIN_TIMED_MS = 50.0           # a timed region shorter than this is repeated and the median reported
MAX_REPEATS = 400
L2_BYTES = 126e6


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

    PERIOD_S = 0.0005

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
            self._halt.wait(self.PERIOD_S if self._nv else 0.1)

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
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
    # e2e: the caller's call with HOST buffers -- actions pinned-host -> device, reward + done device -> pinned-host, in
    # chunks whose copies overlap the launches (ComposedBatch.host_rollout)
    Ke, chunk_e = min(max(K, 256), 512), 32
    h_act = torch.rand((Ke, B, comp.n_act), dtype=torch.float64).pin_memory()
    h_rew = torch.empty((Ke, B), dtype=torch.float64).pin_memory()
    h_done = torch.empty((Ke, B), dtype=torch.uint8).pin_memory()
    batch.host_rollout(h_act[:2 * chunk_e], h_rew[:2 * chunk_e], h_done[:2 * chunk_e], chunk=chunk_e, ring=R)    # untimed: buffers, streams
    restore()
    barrier()
    launch_e = batch.launch_count
    ev0.record()
    batch.host_rollout(h_act, h_rew, h_done, chunk=chunk_e, ring=R)
    ev1.record()
    barrier()
    ms_e = max_over_ranks(ev0.elapsed_time(ev1))
    launch_e = batch.launch_count - launch_e
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
                    "d2h_bytes_per_step": B * 9, "steps": Ke, "chunk_steps": chunk_e, "gpu_launches": launch_e,
                    "api": "ComposedBatch.host_rollout(actions, reward, done) with pinned host tensors",
                    "note": "every step's actions go pinned-host -> device and every step's reward + done come back inside the "
                            "timed region, in chunks of 32 steps; copy-in, kernel and copy-out of neighbouring chunks overlap "
                            "on three streams"},
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
    """The reference arm: the CPU implementation of the path (the C port of Microgrid.run, all host threads) on the headline
    workload's own configuration -- 65 536 envs tiled over the 25 pymgrid25 scenarios, full observation every step -- one
    bench step = a bounded sample of 125 consecutive env steps of that batch."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_envs, n_steps = BATCH_PER_GPU, 125         # one "step" of this arm = 8 192 000 env-steps of the workload
    for _ in range(args.warmup):
        cpu_port_rate(n_envs, 10, threads)
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
        "config": {"workload": WORKLOADS["pymgrid25"], "batch_per_gpu": n_envs, "global_batch": n_envs, "forecast_horizon": 23,
                   "implementation": "CPU port of Microgrid.run (C oracle, oracle/mg_oracle.c), all host threads, bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": python_reference_figure()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def python_reference_figure():
    """The unmodified Python reference's Microgrid.run rate, measured in the build container (it cannot travel)."""
    try:
        with open(os.path.join(ROOT, "profiles", "python_reference_timing.json")) as f:
            d = json.load(f)
        return {"value": d["env_steps_per_s_per_core"], "unit": UNIT + " per core", "cores": 1,
                "where": "build container (no GPU), tools/time_python_reference.py -> profiles/python_reference_timing.json",
                "what": d["what"]}
    except Exception:
        return {"value": None, "note": "profiles/python_reference_timing.json missing; BASELINE.md quotes ~1e3 env-steps/s/core"}


def pin_rank_to_gpu_cpus(local_rank, world):
    """Multi-GPU runs: keep this rank (and the pinned host buffers it is about to allocate) on the cores next to its GPU.
    The ranks whose GPUs report the same CPU set split it evenly.  Returns a description for the JSON line."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        n_cpu = os.cpu_count() or 1
        words = (n_cpu + 63) // 64

        def cpus_of(i):
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), words)
            return tuple(c for c in range(n_cpu) if (mask[c // 64] >> (c % 64)) & 1)
        sets = [cpus_of(i) for i in range(world)]
        mine = sets[local_rank]
        allowed = set(os.sched_getaffinity(0))
        mine = tuple(c for c in mine if c in allowed)
        sharers = [i for i in range(world) if sets[i] == sets[local_rank]]
        k, n = sharers.index(local_rank), len(sharers)
        chunk = mine[k * len(mine) // n:(k + 1) * len(mine) // n]
        if len(chunk) < 2:
            return {"pinned": False, "reason": f"only {len(mine)} cores for {n} ranks"}
        os.sched_setaffinity(0, chunk)
        return {"pinned": True, "cores": len(chunk), "first_core": chunk[0], "ranks_sharing_the_cpu_set": n, "source": "nvmlDeviceGetCpuAffinity"}
    except Exception as ex:
        return {"pinned": False, "reason": f"{type(ex).__name__}: {ex}"}


class Harness:
    """Process-group plumbing and the timing discipline shared by every measurement of this file."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank, self.local_rank, self.world = dist_env()
        self.affinity = pin_rank_to_gpu_cpus(self.local_rank, self.world)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"      # errors only ...
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # ... and off stdout (the version banner too): rank 0 prints ONE JSON line
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local_rank}"))
            self.dist = dist
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.args = args

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def timed(self, run, prepare, min_ms=MIN_TIMED_MS, max_reps=MAX_REPEATS, gate=False):
        """Time `run(rep)` (enqueues EXACTLY the K steps on self.stream) between barrier + synchronize on both sides, CUDA
        events on the launching stream, max over ranks; `prepare(rep)` (state restore, untimed) precedes every repetition.
        Repeats until the timed regions add up to min_ms; returns (median ms, all ms).
        gate=True (device-timed paths): an untimed ~0.1 ms spin kernel runs in front of the start event, so that the event and
        the first launch are both queued before the GPU reaches them -- the region then holds the device's time for the K
        steps, not the host's latency between `record` and the first launch call (10-20 us of Python per call, which is as
        long as a whole step).  The host-buffer legs are timed without it: there the host's work is part of the metric."""
        times, total = [], 0.0
        while True:
            rep = len(times)
            prepare(rep)
            self.barrier()
            if gate:
                try:
                    with self.torch.cuda.stream(self.stream):
                        self.torch.cuda._sleep(200000)
                except Exception:
                    pass
            self.ev0.record(self.stream)
            run(rep)
            self.ev1.record(self.stream)
            self.barrier()
            ms = self.max_over_ranks(self.ev0.elapsed_time(self.ev1))      # identical on every rank: so is the loop count
            times.append(ms)
            total += ms
            if total >= min_ms or len(times) >= max_reps:
                break
        return float(np.median(times)), times

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure_workload(hx, workload, B, K, W, R, paths=("rollout", "graph", "eager"), with_e2e=True, ragged=False, min_ms=MIN_TIMED_MS):
    """One workload on this rank's shard: the persistent-rollout path, the graph-replayed and the eager single-step paths,
    and (with_e2e) the host-buffer legs.  Returns a dict; every collective inside is executed by every rank."""
    torch, args, dev, world, rank, stream = hx.torch, hx.args, hx.dev, hx.world, hx.rank, hx.stream
    discrete = workload == "discrete"
    bm = build_engine(B, dev, rank, world, workload, args.obs_f32)
    if args.no_ring:
        bm.set_rollout_ring(False)
    if args.emit != "auto":
        bm.set_emit_image(args.emit == "image")
    if args.image_shape is not None:
        bm.set_image_shape(args.image_shape)
    if ragged:
        bm.set_ragged(True)                    # (what set_trajectories / a masked reset do on their own)
    if args.no_specialised or (args.emit == "lsu" and ragged):
        bm.set_rollout_specialised(False)      # (LSU emitters, envs at unrelated steps: the plain persistent kernel is the faster one)
    groups = bm.groups
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)

    def rand_actions(steps, g):
        if discrete:
            return torch.randint(0, g.n_actions, (steps, g.n_envs), dtype=torch.int32, device=dev, generator=gen)
        return torch.rand((steps, g.n_envs, g.n_act), dtype=torch.float64, device=dev, generator=gen)
    rings = [torch.empty((R, g.n_envs, g.obs_dim), dtype=bm.obs_dtype, device=dev) for g in groups]
    ring_bytes = sum(r.numel() * r.element_size() for r in rings)
    if ragged:
        for g in groups:
            hi = 8760 - 2 * (W + K) - 64 if 2 * (W + K) < 4000 else 100
            g.step.copy_(torch.randint(0, hi, (g.n_envs,), dtype=torch.int32, device=dev, generator=gen))
    state0 = bm.state_dict()
    step_bytes = sum(g.n_envs * (4 if discrete else 8 * g.n_act) for g in groups)      # action bytes of one step
    # Action blocks of Kc steps: every repetition of a timed region reads a block that the previous repetitions did not
    # leave in L2 (the blocks together exceed 2 x L2), every step its own actions.
    Kc = min(K, 2048)
    n_blocks = int(min(16, max(2, -(-2 * L2_BYTES // (Kc * step_bytes)))))
    blocks = [[rand_actions(Kc, g) for g in groups] for _ in range(n_blocks)]
    act_bytes = n_blocks * Kc * step_bytes
    chunks = [Kc] * (K // Kc) + ([K % Kc] if K % Kc else [])
    out = {"batch_per_gpu": B, "l2": f"inputs larger than L2: obs ring of {R} buffers = {ring_bytes / 1e6:.0f} MB, {n_blocks} action blocks of "
                                     f"{Kc} steps = {act_bytes / 1e6:.0f} MB per GPU, rotated between repetitions (L2 126 MB)"}

    launchers = {}

    def one_step(s):      # single-step launch s: actions of step s % (n_blocks Kc), obs slot s % R
        key = (s % (n_blocks * Kc), s % R)
        if key not in launchers:
            b, k = divmod(key[0], Kc)
            launchers[key] = bm.prepare_step([a[k] for a in blocks[b]], obs=[r[key[1]] for r in rings], discrete=discrete)
        launchers[key]()

    def restore(rep=0):
        bm.load_state_dict(state0)

    results = {}
    with torch.cuda.stream(stream):
        if args.preheat > 0:     # untimed: bring the clocks to their loaded state
            t0, s = time.perf_counter(), 0
            while time.perf_counter() - t0 < args.preheat:
                for _ in range(32):
                    one_step(s % 64)
                    s += 1
                stream.synchronize()
        for path in paths:
            restore()
            if path == "rollout":    # persistent kernel: the K steps run in ceil(K / 2048) launches
                bm.rollout([a[:max(W, 3)] for a in blocks[0]], ring=R, keep_obs=True, discrete=discrete)         # warm-up steps
                outs = {n: bm.rollout([a[:n] for a in blocks[0]], ring=R, keep_obs=True, discrete=discrete) for n in set(chunks)}
                outs = {n: (o if isinstance(o, list) else [o]) for n, o in outs.items()}
                # untimed: bind the argument blocks, so that a timed launch is one C call
                binds = {(b, n): bm.prepare_rollout([a[:n] for a in blocks[b]], ring=R, keep_obs=True, discrete=discrete, out=outs[n])
                         for b in range(n_blocks) for n in set(chunks)}
                launch0 = bm.launch_count

                def run(rep):
                    for n in chunks:
                        binds[rep % n_blocks, n]()
                per_rep = len(chunks)
            elif path in ("graph", "graph_overlap2"):    # one mg_step launch per step, replayed from a CUDA graph
                # ("graph": the library's default chaining of consecutive step launches, MG_OPT_STEP_OVERLAP = 1 -- the next
                #  launch's physics runs under this launch's observation stream; "graph_overlap2": the observation streams
                #  overlap too, which the rotating observation buffers of this loop allow)
                bm.set_step_overlap(2 if path == "graph_overlap2" else 1)
                launchers.clear()
                chunk = K if K <= 256 else max(d for d in range(1, 257) if K % d == 0)
                for s in range(W + chunk):      # make every launcher of the chunk exist before capture
                    one_step(s)
                restore()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    for s in range(W, W + chunk):
                        one_step(s)
                restore()
                for s in range(W):
                    one_step(s)

                def run(rep):
                    for _ in range(K // chunk):
                        graph.replay()
                per_rep = K
            else:                    # eager: one mg_step launch per step from Python
                for s in range(W + min(K, 256)):
                    one_step(s)
                restore()
                for s in range(W):
                    one_step(s)

                def run(rep):
                    for s in range(W, W + K):
                        one_step(s)
                per_rep = K
            sampler = ClockSampler(hx.local_rank)
            sampler.start()
            ms, all_ms = hx.timed(run, restore, min_ms=min_ms if path != "eager" else min(min_ms, 20.0), gate=True)
            clocks = sampler.stop()
            results[path] = {"ms": ms, "repeats": len(all_ms), "ms_min": min(all_ms), "ms_max": max(all_ms), "launches": per_rep, "clocks": clocks,
                             "kernel": bm.last_kernel}
            if path == "graph_overlap2":
                bm.set_step_overlap(1)
                launchers.clear()
    out["paths"] = results
    nbytes = sum(g.n_envs * algorithmic_bytes(*g.arch, discrete=discrete, obs_bytes=4 if args.obs_f32 else 8) for g in groups)
    out["bytes_per_step"] = nbytes
    out["cfg_bytes_single_step"] = sum(g.n_envs * (336 + 8 * g.arch[1]) for g in groups) if workload == "generator" else 0
    out["obs_dims"] = [g.obs_dim for g in groups]

    if not with_e2e:
        del bm
        return out
    # ---- end to end through the public API with host buffers -------------------------------------------------
    # (1) HostIO.step(): one step per call -- actions pinned-host -> device (one copy), fused kernel, reward + done
    #     device -> pinned-host (one copy), serialised on one stream: the figure a host-side control loop sees.
    Ke = min(K, 200)
    restore()
    hio = bm.host_io(normalized=True, obs=[r[0] for r in rings], discrete=discrete)
    for a, g in zip(hio.actions, groups):          # the caller's actions, in pinned host memory
        a.copy_(torch.randint(0, g.n_actions, tuple(a.shape), dtype=torch.int32) if discrete else torch.rand(tuple(a.shape), dtype=torch.float64))
    with torch.cuda.stream(stream):
        for s in range(3):
            hio.step()

        def run_hio(rep):
            for s in range(Ke):
                hio.step()
        ms_hio, reps_hio = hx.timed(run_hio, restore, min_ms=min(min_ms, 30.0), max_reps=50)
    h2d, d2h = hio.h2d_bytes, hio.d2h_bytes
    e2e_step_value = world * B * Ke / (ms_hio * 1e-3)
    e2e = {"value": e2e_step_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke, "repeats": len(reps_hio),
           "api": "BatchedMicrogrid.host_io().step()",
           "note": "actions written into pinned host memory by the caller, one H2D copy, fused kernel, one D2H copy of "
                   "reward+done, every step; observations stay on the device"}
    # (1b) the same loop with the observations ALSO copied to the host every step (a host-side consumer of obs): PCIe-bound
    obs_host = None
    try:
        h_obs = [torch.empty((g.n_envs, g.obs_dim), dtype=bm.obs_dtype, pin_memory=True) for g in groups]
        n_o = min(Ke, 24)
        with torch.cuda.stream(stream):
            def run_obs(rep):
                for s in range(n_o):
                    hio.step()
                    for h, r in zip(h_obs, rings):
                        h.copy_(r[0], non_blocking=True)
            ms_o, _ = hx.timed(run_obs, restore, min_ms=min(min_ms, 30.0), max_reps=5)
        obs_bytes = sum(h.numel() * h.element_size() for h in h_obs)
        obs_host = {"value": world * B * n_o / (ms_o * 1e-3), "unit": UNIT, "d2h_bytes_per_step": d2h + obs_bytes, "steps": n_o,
                    "note": "HostIO.step() plus a device -> pinned-host copy of EVERY observation row each step: what a host-side "
                            "consumer of the observations would see (the obs are 97% of the bytes; PCIe is the bound)"}
        del h_obs
    except Exception as ex:
        obs_host = {"error": f"{type(ex).__name__}: {ex}"}
    # (2) HostRollout.run(): the workload's own call (a year rollout with pre-generated actions) with HOST buffers --
    #     every step's actions cross the bus host -> device and every step's reward + done come back, in chunks of
    #     `chunk` steps, the copies of neighbouring chunks overlapped with the persistent kernel on three streams.
    #     Same bytes per step as (1); this is the headline e2e figure when it runs (any failure keeps (1) and says so).
    # Collectives (barrier, max over ranks) stay outside the try blocks so that a failure on one rank cannot hang the others.
    Kr = min(max(K, 256), 1024)      # its own length: a host-buffer rollout shorter than ~256 steps measures pipeline fill and drain
    chunk = max(1, min(64, Kr // (8 if Kr >= 128 else 4)))     # >= 4-8 chunks: the three-stream pipeline is exercised at any K
    # (short rollouts take fewer, larger chunks: a chunk costs ~12 API calls on the host, which must stay ahead of the GPU)
    del hio
    errors = {}
    for pipeline in ("native", "torch"):       # mg_rollout_host (one C-ABI call); else the same schedule from torch streams
        err, hr, launch_e = None, None, 0
        try:
            restore()
            hr = bm.host_rollout(Kr, chunk=chunk, normalized=True, discrete=discrete, ring=R, pipeline=pipeline)
            for a, g in zip(hr.actions, groups):
                if discrete:
                    a.copy_(torch.randint(0, g.n_actions, tuple(a.shape), dtype=torch.int32))
                else:
                    a.uniform_(0.0, 1.0)
            with torch.cuda.stream(stream):
                hr.run(min(Kr, 3 * chunk) if Kr % chunk == 0 else Kr)      # warm-up (untimed)
                restore()
        except Exception as ex:      # keep the per-step figure; never lose the bench line to this path
            err = f"{type(ex).__name__}: {ex}"
        ok = hx.max_over_ranks(0.0 if err is None else 1.0) == 0.0
        ms_e, reps_e = float("inf"), []
        if ok:
            state = {"err": None}

            def run_hr(rep):
                if state["err"] is None:
                    try:
                        hr.run()
                    except Exception as ex:
                        state["err"] = f"{type(ex).__name__}: {ex}"
            with torch.cuda.stream(stream):
                launch_e = bm.launch_count
                ms_e, reps_e = hx.timed(run_hr, restore, min_ms=min_ms, max_reps=20)
                launch_e = (bm.launch_count - launch_e) // max(len(reps_e), 1)
            err = state["err"]
            if err is None and not all(bool(torch.isfinite(r).all()) for r in hr.reward):
                err = "non-finite reward came back from the host rollout"
            ok = hx.max_over_ranks(0.0 if err is None else 1.0) == 0.0
        if ok:
            name = "mg_rollout_host" if pipeline == "native" else "mg_rollout + torch streams"
            e2e = {"value": world * B * Kr / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": hr.h2d_bytes_per_step,
                   "d2h_bytes_per_step": hr.d2h_bytes_per_step, "steps": Kr, "chunk_steps": chunk, "repeats": len(reps_e),
                   "gpu_launches": launch_e, "us_per_step": 1e3 * ms_e / Kr,
                   "pcie_gbs": {"h2d": hr.h2d_bytes_per_step * Kr / (ms_e * 1e-3) / 1e9, "d2h": hr.d2h_bytes_per_step * Kr / (ms_e * 1e-3) / 1e9,
                                "note": "per rank (GB/s of this rank's host <-> device copies; the slowest rank sets the time)"},
                   "api": f"BatchedMicrogrid.host_rollout(n_steps).run() -> {name}",
                   "note": "year-rollout call with HOST buffers: every step's actions go pinned-host -> device and every step's "
                           f"reward + done come back device -> pinned-host inside the timed region, in chunks of {chunk} steps; copy-in, "
                           "persistent kernel and copy-out of neighbouring chunks overlap on three streams (PCIe-bound); "
                           "observations stay in the device ring (97% of the bytes: see obs_to_host for a host-side consumer of them)",
                   "per_step_call": {"value": e2e_step_value, "api": "BatchedMicrogrid.host_io().step()", "steps": Ke,
                                     "note": "one H2D + kernel + one D2H per step, serialised (a host-side control loop)"}}
        else:
            errors[pipeline] = err or "failed on another rank"
        hr = None
        if ok:
            break
    if errors:
        e2e["host_rollout_errors"] = errors
    # (3) the same host-buffer rollout with FLOAT32 actions (what a policy network emits; widened exactly on the device, the
    #     arithmetic stays f64: bm.set_action_dtype) -- half the host -> device bytes.  Reported beside the f64 figure, which
    #     stays the headline (the reference's callers pass float64).
    if not discrete:
        err, hr = None, None
        try:
            restore()
            bm.set_action_dtype(torch.float32)
            hr = bm.host_rollout(Kr, chunk=chunk, normalized=True, ring=R)
            for a in hr.actions:
                a.uniform_(0.0, 1.0)
            with torch.cuda.stream(stream):
                hr.run(min(Kr, 3 * chunk) if Kr % chunk == 0 else Kr)
                restore()
        except Exception as ex:
            err = f"{type(ex).__name__}: {ex}"
        if hx.max_over_ranks(0.0 if err is None else 1.0) == 0.0:
            state = {"err": None}

            def run_f32(rep):
                if state["err"] is None:
                    try:
                        hr.run()
                    except Exception as ex:
                        state["err"] = f"{type(ex).__name__}: {ex}"
            with torch.cuda.stream(stream):
                ms_f, reps_f = hx.timed(run_f32, restore, min_ms=min_ms, max_reps=20)
            err = state["err"]
            if err is None:
                e2e["float32_actions"] = {"value": world * B * Kr / (ms_f * 1e-3), "unit": UNIT,
                                          "h2d_bytes_per_step": hr.h2d_bytes_per_step, "d2h_bytes_per_step": hr.d2h_bytes_per_step,
                                          "steps": Kr, "chunk_steps": chunk, "repeats": len(reps_f),
                                          "note": "the same call after set_action_dtype(torch.float32): the actions cross the bus as "
                                                  "float32 and are widened exactly on the device (bit-identical to float64 actions of "
                                                  "the same values)"}
        if err is not None:
            e2e["float32_actions"] = {"error": err}
        hr = None
        bm.set_action_dtype(torch.float64)
    e2e["obs_to_host"] = obs_host
    out["e2e"] = e2e
    del bm
    return out


def roofline_block(m, path, K, args, workload):
    """roofline object of one measured path: algorithmic bytes per launch / average launch duration, vs the measured copy peak"""
    r = m["paths"][path]
    peak, peak_src = measured_peak()
    per_step = m["bytes_per_step"] + (m["cfg_bytes_single_step"] if path != "rollout" else 0)
    # (+ the env's own parameter record and status word(s), re-read by every single-step launch of the generator workload:
    #  SURVEY.md 8d; inside the persistent kernel they stay cache-resident and are not counted)
    launches = r["launches"]
    steps_per_launch = K / max(launches, 1)
    achieved = per_step * K / (r["ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "kernel": r["kernel"],
            "bytes_per_launch": per_step * steps_per_launch, "bytes_per_step": per_step, "steps_per_launch": steps_per_launch}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--batch", type=int, default=None, help="envs per GPU (default 65536; replicas 4096; generator 131072)")
    ap.add_argument("--workload", default="pymgrid25", choices=tuple(WORKLOADS), help="default = the BASELINE metric's workload")
    ap.add_argument("--ring", type=int, default=4, help="observation buffers rotated so stores reach HBM")
    ap.add_argument("--path", default="rollout", choices=("rollout", "graph", "graph_overlap2", "eager"),
                    help="headline path: the persistent rollout kernel (BASELINE configs[2] is a year rollout with pre-generated "
                         "actions), one mg_step launch per step replayed from a CUDA graph, or plain launches from Python")
    ap.add_argument("--single-path", action="store_true", help="time only the headline path (no other paths, no e2e, no configs block)")
    ap.add_argument("--no-configs", action="store_true", help="skip the in-run measurements of the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--obs-f32", action="store_true",
                    help="NON-CANONICAL secondary mode: observations written as float32 (half the dominant bytes); reported with dtype f64+f32obs")
    ap.add_argument("--ragged", action="store_true",
                    help="start every env at its own random step (envs that reset independently): no two rows of a tile share a window")
    ap.add_argument("--no-ring", action="store_true",
                    help="per-env series batches (generator workload): use the persistent kernel that normalises whole windows per "
                         "row instead of the one that keeps sliding windows in shared memory (A/B)")
    ap.add_argument("--emit", default="auto", choices=("auto", "image", "lsu"),
                    help="row emitter: the library's choice per launch (default), shared-memory images + TMA bulk stores, or the per-lane "
                         "16-byte store emitters (A/B)")
    ap.add_argument("--image-shape", type=int, default=None, help="MG_OPT_IMAGE_SHAPE index (tuning)")
    ap.add_argument("--no-specialised", action="store_true", help="persistent kernel without the owner / emitter warp split (A/B)")
    ap.add_argument("--min-timed-ms", type=float, default=MIN_TIMED_MS, help="repeat a timed region until it adds up to this much")
    ap.add_argument("--clock-period-ms", type=float, default=0.5, help="NVML polling period of the clock sampler during timed regions")
    ap.add_argument("--preheat", type=float, default=0.2, help="seconds of untimed steps before the warm-up (0 under ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    ClockSampler.PERIOD_S = args.clock_period_ms * 1e-3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "composed":
        return run_composed(args)

    hx = Harness(args)
    rank, world = hx.rank, hx.world
    if args.batch is None:
        args.batch = {"replicas": 4096, "generator": 131072}.get(args.workload, BATCH_PER_GPU)
    B, K, W, R = args.batch, args.steps, args.warmup, args.ring
    paths = (args.path,) if args.single_path else (args.path,) + tuple(p for p in ("rollout", "graph", "graph_overlap2", "eager") if p != args.path)
    m = measure_workload(hx, args.workload, B, K, W, R, paths=paths, with_e2e=not args.single_path, ragged=args.ragged, min_ms=args.min_timed_ms)
    head = m["paths"][args.path]
    value = world * B * K / (head["ms"] * 1e-3)
    # ---- the other BASELINE configs, measured briefly in the same run (every rank runs its shard) ----------------------
    configs = None
    if args.workload == "pymgrid25" and not args.single_path and not args.no_configs and not args.ragged and not args.obs_f32:
        configs = {}
        Kc = 256      # (their own length, whatever --steps: a 20-step launch of the per-env-series kernel is half prologue)
        for name, wl, Bc, cpaths in (("configs[1]", "replicas", 4096, ("graph", "rollout")), ("configs[3]", "discrete", BATCH_PER_GPU, ("rollout", "graph")),
                                     ("configs[4]", "generator", 131072, ("rollout",))):
            try:
                mc = measure_workload(hx, wl, Bc, Kc, W, R, paths=cpaths, with_e2e=False, min_ms=min(args.min_timed_ms, 30.0))
                entry = {"workload": WORKLOADS[wl], "batch_per_gpu": Bc, "global_batch": world * Bc, "steps": Kc}
                for pth in cpaths:
                    r = mc["paths"][pth]
                    rb = roofline_block(mc, pth, Kc, args, wl)
                    entry[pth] = {"value": world * Bc * Kc / (r["ms"] * 1e-3), "unit": UNIT, "us_per_step": 1e3 * r["ms"] / Kc, "repeats": r["repeats"],
                                  "kernel": r["kernel"],
                                  "roofline_frac": rb["frac"], "achieved_gbs": rb["achieved"], "bytes_per_step": rb["bytes_per_step"]}
                entry["headline_path"] = cpaths[0]
                configs[name] = entry
            except Exception as ex:      # (collectives inside measure_workload are unconditional; a failure here is deterministic on every rank)
                configs[name] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        roof = roofline_block(m, args.path, K, args, args.workload)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tr = json.load(f)[roof["kernel"].split(" ")[0]]
            if tr["dram_bytes_per_step"] and B == BATCH_PER_GPU and args.workload == "pymgrid25" and not args.obs_f32 and not args.ragged:
                traffic, traffic_src = tr["dram_bytes_per_step"] * roof["steps_per_launch"], tr["source"]
        except Exception:
            pass
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "write_only_ceilings_gbs": {"lsu_16_byte_stores": 5780.0, "tma_bulk_stores_4_rows": 7200.0},
                     "note": "peak is the read+write copy bandwidth (torch copy); this path is ~97% stores: tools/microbench_store.cu measured 5.78 TB/s "
                             "for per-lane 16-byte stores and 7.2 TB/s for TMA bulk stores of 4 rows on this GPU (profiles/r02_microbench_store.txt)"})
        others = {p: {"value": world * B * K / (r["ms"] * 1e-3), "us_per_step": 1e3 * r["ms"] / K, "gpu_launches": r["launches"], "repeats": r["repeats"]}
                  for p, r in m["paths"].items() if p != args.path}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if not args.obs_f32 else "f64 arithmetic, f32 observation output (non-canonical)",
            "data": "pymgrid25 scenario parameters + series (bundled), synthetic U[0,1) actions",
            "repeats": head["repeats"],
            "timing": {"rule": f"the K-step timed region is repeated until the repetitions add up to {args.min_timed_ms:.0f} ms (state restored and a fresh "
                               "action block in between, untimed); value / ms_per_step are the MEDIAN repetition, max over ranks each; an untimed "
                               "0.1 ms spin kernel in front of the start event keeps the host's launch latency out of the device-timed regions",
                       "ms_min": head["ms_min"], "ms_median": head["ms"], "ms_max": head["ms_max"]},
            "config": {"workload": WORKLOADS[args.workload],
                       "batch_per_gpu": B, "global_batch": world * B, "forecast_horizon": 23, "path": args.path, "ragged_steps": bool(args.ragged),
                       "emit": args.emit, "image_shape": args.image_shape, "specialised": not args.no_specialised, "l2": m["l2"],
                       "parallelism": f"batch sharded over {world} GPU(s), no collective on the step path", "cpu_affinity": hx.affinity},
            "gpu_launches": head["launches"],
            "clocks": head["clocks"],
            "e2e": m.get("e2e"),
            "roofline": roof,
            "other_paths": others,
        }
        if configs is not None:
            line["configs"] = configs
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
                                    "python_reference": python_reference_figure()}
        print(json.dumps(line))
    hx.finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
