"""Pin the numpy restatement of GaussianNoiseForecaster (oracle/forecast_noise.py) to the reference: with the legacy
numpy seed the reference used, the restatement must reproduce the recorded noisy observations BIT FOR BIT -- standard
deviation table (relative / increasing), draw order of the modules, padding past the end of the series, clipping."""
import numpy as np
import pytest

from oracle.forecast_noise import NoisyModule
from oracle.oracle import OracleGrid
from pymgrid_b200.scenario import load_pymgrid25
from pymgrid_b200 import views
from tests.helpers import jump_to

NOISE_CASES = (   # the cases of tests/golden/make_golden.py: scenario, {module: (std, increase_uncertainty, relative_noise)}
    (0, dict(load=(40.0, False, False), pv=(0.15, True, True), grid=(0.05, True, False))),
    (2, dict(load=(0.1, True, True), pv=(25.0, False, False))),
    (1, dict(load=(0.2, False, True), pv=(0.3, True, True), grid=(0.02, False, True))),
)
RESET_ORDER = ("load", "pv", "grid")     # Microgrid.reset walks the module listing (microgrid.py:205-219)
RUN_ORDER = ("load", "grid", "pv")       # Microgrid.run: fixed, controllable, flex (microgrid.py:255-314)


def noisy_modules(p, spec):
    series = dict(load=(p.load_ts, True), pv=(p.pv_ts, True))
    if p.grid is not None:
        series["grid"] = (p.grid.time_series, False)
    return {name: NoisyModule(series[name][0], p.forecast_horizon, series[name][1], *spec[name], p.initial_step, p.final_step)
            for name in spec}


def splice(obs, p, blocks):
    out = obs.copy()
    for name, sl in views.obs_slices(p, "gym_sorted").items():
        if name in blocks:
            out[sl] = blocks[name]
    return out


@pytest.mark.parametrize("ci", range(len(NOISE_CASES)))
@pytest.mark.parametrize("seg", (0, 1))
def test_restatement_reproduces_reference_draws(golden, ci, seg):
    z = golden["noisy_forecast"]
    n, spec = NOISE_CASES[ci]
    key = f"c{ci}_{seg}"
    p = jump_to(load_pymgrid25(n), int(z[key + "_t0"]))
    mods = noisy_modules(p, spec)
    np.testing.assert_allclose([np.mean(mods[name].std) for name in spec], z[key + "_noise_std"], rtol=0, atol=0)
    o = OracleGrid(p)
    np.random.seed(500 + ci)
    clean = o.reset()
    t = int(z[key + "_t0"])
    blocks = {name: mods[name].observe(t) for name in RESET_ORDER if name in mods}
    np.testing.assert_array_equal(splice(clean, p, blocks), z[key + "_reset_obs"])
    for k, a in enumerate(z[key + "_a"]):
        clean, r, _, _, _ = o.run(a)
        t += 1
        blocks = {name: mods[name].observe(t) for name in RUN_ORDER if name in mods}
        noisy = splice(clean, p, blocks)
        np.testing.assert_array_equal(noisy, z[key + "_o"][k], err_msg=f"step {k}")
        assert r == z[key + "_r"][k]                 # the forecaster never touches the physics
        assert noisy.min() >= 0.0 and noisy.max() <= 1.0
    if seg == 1:   # across the end: padded rows carry no noise and sit at the clipped fill value
        sl = views.obs_slices(p, "gym_sorted")["load"]
        assert np.array_equal(noisy[sl][2:], clean[sl][2:]) and not np.array_equal(z[key + "_o"][0][sl], o.reset()[sl])
