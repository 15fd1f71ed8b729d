"""Cost of the Gaussian-noise forecaster post-pass (mg_forecast_noise) at the bench batch: 65 536 pymgrid25 envs with noise on
every time-series module, single steps issued back to back, with and without the post-pass.  GPU box only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymgrid_b200.engine import BatchedMicrogrid  # noqa: E402
from pymgrid_b200.params import ForecasterParams  # noqa: E402
from pymgrid_b200.scenario import load_pymgrid25  # noqa: E402


def main():
    B, K = 65536, 200
    configs = [load_pymgrid25(n) for n in range(25)]
    for p in configs:
        p.forecasters = dict(load=ForecasterParams(0.1, True, True), pv=ForecasterParams(0.1, True, True))
        if p.grid is not None:
            p.forecasters["grid"] = ForecasterParams(0.05, True, False)
    env_config = np.arange(B) % 25
    bm = BatchedMicrogrid(configs, env_config, device="cuda:0", with_info=False)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    acts = [[torch.rand((g.n_envs, g.n_act), dtype=torch.float64, device="cuda", generator=gen) for g in bm.groups] for _ in range(8)]
    rings = [torch.empty((4, g.n_envs, g.obs_dim), dtype=torch.float64, device="cuda") for g in bm.groups]
    state0 = bm.state_dict()
    obs_bytes = sum(g.n_envs * g.obs_dim * 8 for g in bm.groups)
    for label in ("noise", "oracle"):
        if label == "oracle":
            bm.clear_forecast_noise()
        for s in range(8):
            bm.step(acts[s % 8], obs=[r[s % 4] for r in rings])
        bm.load_state_dict(state0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(K):
            bm.step(acts[s % 8], obs=[r[s % 4] for r in rings])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        print(f"{label:7s}: {us:8.2f} us/step  ({B / us * 1e6:.3e} env-steps/s; obs rows {obs_bytes / 1e6:.1f} MB/step)", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
