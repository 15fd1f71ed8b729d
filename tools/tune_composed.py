"""Time the composed-path kernel (mgc_run) on bench.py's composed workload: gather emission vs per-element decode, with and
without observation rows (owner phase alone).  GPU box only; prints one line per variant."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from pymgrid_b200.compose import ComposedBatch, Composition
    dev = torch.device("cuda:0")
    B, K = args.batch, args.steps
    for gather in ("1", "0"):
        os.environ["PYMGRID_B200_COMPOSE_GATHER"] = gather
        comp = Composition(bench.composed_modules(), loss_load_cost=10.0, overgeneration_cost=1.0)
        batch = ComposedBatch([comp], np.zeros(B, dtype=np.int64), device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(2)
        actions = torch.rand((K, B, comp.n_act), dtype=torch.float64, device=dev, generator=gen)
        state0 = [t.clone() for t in (batch.step_counter, batch.fstate, batch.istate)]
        for obs in (True, False):
            kw = dict(ring=4) if obs else dict(obs=False)
            out = batch.rollout(actions, **kw)
            times = []
            for _ in range(args.reps):
                for t, s0 in zip((batch.step_counter, batch.fstate, batch.istate), state0):
                    t.copy_(s0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                batch.rollout(actions, out=out, **kw)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) * 1e3 / K)
            us = float(np.median(times))
            print(f"gather={gather} obs={int(obs)} B={B} K={K}: {us:8.2f} us/step  {B / us * 1e6:.3e} env-steps/s", flush=True)
        del batch, out
    return 0


if __name__ == "__main__":
    sys.exit(main())
