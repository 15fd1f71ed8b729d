#!/usr/bin/env python
"""A/B of the row emitters on one GPU: every (emitter, image shape, warp split) variant of the persistent kernel and of the
graph-replayed single-step kernel on the bench workloads, one engine per workload, CUDA-event timing.

    python tools/tune_emitters.py [--steps 400] [--workloads pymgrid25,ragged,generator,replicas,discrete] [--variants all|default]

Prints one JSON line per (workload, variant): us/step, algorithmic TB/s, fraction of MEASURED_PEAKS.json's copy bandwidth.
Design data for pymgrid_b200/csrc/mg_engine.cu (which variant the host selects by default); bench.py is the judged number."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--workloads", default="pymgrid25,ragged,generator,replicas,discrete")
    ap.add_argument("--variants", default="all")
    ap.add_argument("--step-path", action="store_true", help="also time graph-replayed single steps")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peak, _ = bench.measured_peak()
    variants = [("lsu", 0, True), ("lsu", 0, False)] + [("image", s, ws) for s in range(6) for ws in (True, False)]
    if args.variants == "image_ws":
        variants = [("image", s, True) for s in range(6)]
    if args.variants == "lsu_ws":
        variants = [("lsu", 0, True), ("image", 0, True), ("image", 1, True)]
    if args.variants == "default":
        variants = [("lsu", 0, True), ("image", 0, True), ("image", 0, False)]
    K = args.steps
    for wl in args.workloads.split(","):
        ragged = wl == "ragged"
        name = "pymgrid25" if ragged else wl
        B = {"replicas": 4096, "generator": 131072}.get(name, 65536)
        bm = bench.build_engine(B, dev, 0, 1, name)
        discrete = name == "discrete"
        gen = torch.Generator(device=dev)
        gen.manual_seed(3)
        if ragged:
            for g in bm.groups:
                g.step.copy_(torch.randint(0, 8760 - 2 * K - 64, (g.n_envs,), dtype=torch.int32, device=dev, generator=gen))
        state0 = bm.state_dict()
        acts = [torch.randint(0, g.n_actions, (K, g.n_envs), dtype=torch.int32, device=dev, generator=gen) if discrete
                else torch.rand((K, g.n_envs, g.n_act), dtype=torch.float64, device=dev, generator=gen) for g in bm.groups]
        nbytes = sum(g.n_envs * bench.algorithmic_bytes(*g.arch, discrete=discrete) for g in bm.groups)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rings = [torch.empty((4, g.n_envs, g.obs_dim), dtype=torch.float64, device=dev) for g in bm.groups]
        # physics only (no observation rows): what the owner threads alone sustain
        bm.load_state_dict(state0)
        launch = bm.prepare_rollout(acts, ring=1, keep_obs=False, discrete=discrete)
        launch()
        bm.load_state_dict(state0)
        torch.cuda.synchronize()
        ev0.record()
        launch()
        ev1.record()
        torch.cuda.synchronize()
        print(json.dumps({"workload": wl, "batch": B, "path": "rollout, no observations", "us_per_step": round(1e3 * ev0.elapsed_time(ev1) / K, 3)}), flush=True)
        wl_variants = [v + (True,) for v in variants]
        if name == "generator":      # per-env series: also without the shared-memory rings (whole windows normalised per row)
            wl_variants += [v + (False,) for v in variants if v[0] == "image" and v[1] == 0]
        for emit, shape, ws, ring in wl_variants:
            bm.set_emit_image(emit == "image")
            bm.set_image_shape(shape)
            bm.set_rollout_specialised(ws)
            bm.set_rollout_ring(ring)
            try:
                bm.load_state_dict(state0)
                launch = bm.prepare_rollout(acts, ring=4, keep_obs=True, discrete=discrete)
                bm.rollout([a[:8] for a in acts], ring=4, discrete=discrete)
                bm.load_state_dict(state0)
                torch.cuda.synchronize()
                if hasattr(bm._lib, "mg_debug_role_cycles"):
                    import ctypes
                    bm._lib.mg_debug_role_cycles((ctypes.c_ulonglong * 8)(), 1)
                ev0.record()
                launch()
                ev1.record()
                torch.cuda.synchronize()
                us = 1e3 * ev0.elapsed_time(ev1) / K
                line = {"workload": wl, "batch": B, "path": "rollout", "emit": emit, "shape": shape, "specialised": ws, "ring": ring,
                        "us_per_step": round(us, 3), "tbs": round(nbytes / us / 1e6, 3), "frac": round(nbytes / us / 1e3 / peak, 3)}
                if hasattr(bm._lib, "mg_debug_role_cycles"):     # -DMG_ROLE_TIMERS build: average cycles per warp-step by role
                    import ctypes
                    buf = (ctypes.c_ulonglong * 8)()
                    bm._lib.mg_debug_role_cycles(buf, 1)
                    n = max(buf[5], 1)
                    line["role_cycles_per_warp_step"] = {"owner_physics": round(buf[0] / n), "owner_wait_empty": round(buf[1] / n),
                                                         "emitter_wait_full": round(buf[2] / n), "emitter_wait_buffer": round(buf[3] / n),
                                                         "emitter_busy_incl_buffer_wait": round(buf[4] / n)}
            except Exception as ex:     # a variant that cannot launch (shared memory) must not lose the others
                line = {"workload": wl, "emit": emit, "shape": shape, "specialised": ws, "error": f"{type(ex).__name__}: {ex}"}
            print(json.dumps(line), flush=True)
            for overlap in (0, 1, 2) if (args.step_path and ws and ring) else ():
                try:
                    bm.set_step_overlap(overlap)
                    bm.load_state_dict(state0)
                    n = min(K, 128)
                    launchers = [bm.prepare_step([a[s] for a in acts], obs=[r[s % 4] for r in rings], discrete=discrete) for s in range(n)]
                    for f in launchers[:4]:
                        f()
                    bm.load_state_dict(state0)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        for f in launchers:
                            f()
                    bm.load_state_dict(state0)
                    graph.replay()
                    bm.load_state_dict(state0)
                    torch.cuda.synchronize()
                    ev0.record()
                    graph.replay()
                    ev1.record()
                    torch.cuda.synchronize()
                    us = 1e3 * ev0.elapsed_time(ev1) / n
                    line = {"workload": wl, "batch": B, "path": f"graph, step overlap {overlap}", "emit": emit, "shape": shape,
                            "us_per_step": round(us, 3), "tbs": round(nbytes / us / 1e6, 3), "frac": round(nbytes / us / 1e3 / peak, 3)}
                except Exception as ex:
                    line = {"workload": wl, "path": "graph", "emit": emit, "shape": shape, "error": f"{type(ex).__name__}: {ex}"}
                print(json.dumps(line), flush=True)
                bm.set_step_overlap(1)
        del bm, acts, rings
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
