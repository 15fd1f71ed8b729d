#!/usr/bin/env python
"""Time the UNMODIFIED Python reference's Microgrid.run loop (BASELINE configs[0] / north_star's "reference CPU
Microgrid.run loop") in the build container, where /root/reference exists; the GPU box has no copy of it.

    python tools/time_python_reference.py            # writes profiles/python_reference_timing.json

pymgrid25 scenarios 0 (grid only), 1 (genset + grid) and 2 (genset only): `Microgrid.run(sample_action(), normalized=True)`
for N steps each after a short warm-up, one process, one core.  bench.py quotes the file in `cpu_baseline.python_reference`
(a build-container figure with its provenance; the same-box CPU baseline is the C port)."""
import json
import os
import platform
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from oracle.ref_loader import load_reference  # noqa: E402

load_reference()
from pymgrid import Microgrid  # noqa: E402


def main(n_steps=3000):
    out = {"what": "unmodified reference (Total-RD/pymgrid @ /root/reference) Microgrid.run(control, normalized=True), "
                   "random actions from Microgrid.sample_action, one Python process on one core, build container",
           "python": platform.python_version(), "numpy": np.__version__, "cpu": platform.processor() or platform.machine(),
           "n_steps": n_steps, "scenarios": {}}
    rates = []
    for n in (0, 1, 2):
        m = Microgrid.from_scenario(microgrid_number=n)
        m.reset()
        np.random.seed(n)
        actions = [m.sample_action() for _ in range(n_steps + 50)]
        for a in actions[:50]:
            m.run(a, normalized=True)
        t0 = time.perf_counter()
        for a in actions[50:]:
            m.run(a, normalized=True)
        dt = time.perf_counter() - t0
        out["scenarios"][str(n)] = {"steps_per_s": n_steps / dt, "seconds": dt}
        rates.append(n_steps / dt)
    out["env_steps_per_s_per_core"] = float(np.mean(rates))
    path = os.path.join(ROOT, "profiles", "python_reference_timing.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
