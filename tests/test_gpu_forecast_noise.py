"""GaussianNoiseForecaster on the GPU (mg_forecast_noise).  The reference draws from numpy's global generator, the
kernel from a counter-based one, so parity has two legs:
  * against a numpy restatement of the KERNEL's generator (tests/helpers.py): every element within 1e-12 (libm vs CUDA
    log / sincos differ in the last bits), untouched entries bit-exact, rows past the end untouched;
  * against the reference-pinned restatement of GaussianNoiseForecaster (oracle/forecast_noise.py, bit-exact with the
    reference under its seed): same distribution per element -- mean, standard deviation and the mass the clip piles up
    on the bounds."""
import numpy as np
import pytest
import torch

from oracle.forecast_noise import NoisyModule
from oracle.oracle import OracleGrid
from pymgrid_b200 import views
from pymgrid_b200.params import ForecasterParams
from pymgrid_b200.scenario import load_pymgrid25
from tests.helpers import engine_noisy_row, jump_to

pytestmark = pytest.mark.gpu
SPEC = {0: dict(load=ForecasterParams(40.0, False, False), pv=ForecasterParams(0.15, True, True),
                grid=ForecasterParams(0.05, True, False)),
        2: dict(load=ForecasterParams(0.1, True, True), pv=ForecasterParams(25.0, False, False)),
        1: dict(load=ForecasterParams(0.2, False, True), pv=ForecasterParams(0.3, True, True),
                grid=ForecasterParams(0.02, False, True))}


def noisy_params(n, t0=0):
    p = jump_to(load_pymgrid25(n), t0)
    p.forecasters = dict(SPEC[n])
    return p


def engine(configs, env_config, **kw):
    from pymgrid_b200.engine import BatchedMicrogrid
    return BatchedMicrogrid(configs, env_config, device="cuda:0", **kw)


def sigma_tables(p):
    from pymgrid_b200.engine import forecast_noise_record
    rec = forecast_noise_record(p)
    sigma = dict(load=[rec.load_sigma], pv=[rec.pv_sigma], grid=list(rec.grid_sigma[:]))
    inc = dict(load=bool(rec.load_increase), pv=bool(rec.pv_increase), grid=bool(rec.grid_increase))
    return sigma, inc


@pytest.mark.parametrize("order", ("gym_sorted", "container"))
def test_kernel_equals_generator_restatement(order):
    """three architectures, start of the year and across the end of the series, step + reset + observe"""
    configs = [noisy_params(n, t0) for n in (0, 1, 2) for t0 in (0, 8752)]
    B = 60
    env_config = np.arange(B) % len(configs)
    bm = engine(configs, env_config, obs_order=order)
    clean = engine([jump_to(load_pymgrid25(n), t0) for n in (0, 1, 2) for t0 in (0, 8752)], env_config, obs_order=order)
    bm.set_forecast_noise(seed=0x1234567890abcdef, env_offset=(1 << 33) + 5)
    rng = np.random.default_rng(3)
    bases = np.cumsum([0] + [g.n_envs for g in bm.groups])
    call = 0
    for k in range(16):
        acts = [torch.from_numpy(rng.random((g.n_envs, g.n_act))).cuda() for g in bm.groups]
        if k == 7:
            got, want = bm.reset(), clean.reset()
        elif k == 11:
            got, want = bm.observe(), clean.observe()
        else:
            got, want = bm.step(acts)[0], clean.step(acts)[0]
        call += 1
        for gi, g in enumerate(bm.groups):
            assert torch.equal(g.step, clean.groups[gi].step) and torch.equal(g.reward, clean.groups[gi].reward)
            steps = g.step.cpu().numpy()
            a, b = got[gi].cpu().numpy(), want[gi].cpu().numpy()
            for slot in range(g.n_envs):
                p = configs[env_config[g.env_ids[slot]]]
                sigma, inc = sigma_tables(p)
                exp = engine_noisy_row(b[slot], p, order, sigma, inc, (1 << 33) + 5 + int(bases[gi]) + slot, int(steps[slot]),
                                       0x1234567890abcdef, call, 8760)
                np.testing.assert_allclose(a[slot], exp, rtol=0, atol=1e-12, err_msg=f"call {k} group {gi} slot {slot}")
                same = exp == b[slot]
                assert np.array_equal(a[slot][same], b[slot][same])      # untouched entries are bit-exact
                assert same.sum() < len(exp) or steps[slot] >= 8759
    # every env ran to the end of its series in the second half of the configs: rows there carry no noise at all
    assert any((g.step.cpu().numpy() >= 8759).any() for g in bm.groups)


@pytest.mark.parametrize("n", (0, 1, 2))
def test_distribution_matches_the_reference_forecaster(n):
    """8 192 replicas at one step vs 8 192 draws of the reference-pinned restatement: per-element mean and standard
    deviation (the clip included) agree within sampling error; current values and constant columns stay put."""
    B, t0 = 8192, 100
    p = noisy_params(n, t0)
    bm = engine([p], np.zeros(B, dtype=np.int64))
    bm.set_forecast_noise(seed=99)
    got = bm.observe()
    got = (got if isinstance(got, torch.Tensor) else got[0]).cpu().numpy()
    clean = OracleGrid(jump_to(load_pymgrid25(n), t0)).observe()
    series = dict(load=(p.load_ts, True), pv=(p.pv_ts, True))
    if p.grid is not None:
        series["grid"] = (p.grid.time_series, False)
    np.random.seed(4)
    for name, sl in views.obs_slices(p, "gym_sorted").items():
        if name not in series:
            np.testing.assert_array_equal(got[:, sl], np.broadcast_to(clean[sl], (B, sl.stop - sl.start)))
            continue
        f = p.forecasters[name]
        mod = NoisyModule(series[name][0], 23, series[name][1], f.noise_std, f.increase_uncertainty, f.relative_noise,
                          p.initial_step, p.final_step)
        ref = np.stack([mod.observe(t0) for _ in range(B)])
        C = mod.ts.shape[1]
        np.testing.assert_array_equal(got[:, sl][:, :C], ref[:, :C])          # current values: no noise
        assert got[:, sl].min() >= 0.0 and got[:, sl].max() <= 1.0
        gm, rm, gs, rs = got[:, sl].mean(0), ref.mean(0), got[:, sl].std(0), ref.std(0)
        se = np.maximum(rs, 1e-12) / np.sqrt(B)
        assert (np.abs(gm - rm) <= 6 * np.sqrt(2) * se + 1e-12).all(), name
        assert (np.abs(gs - rs) <= 0.06 * rs + 1e-12).all(), name
        # the clip's point masses
        for bound in (0.0, 1.0):
            assert (np.abs((got[:, sl] == bound).mean(0) - (ref == bound).mean(0)) <= 0.03).all(), (name, bound)
        const = rs == 0
        assert np.array_equal(got[:, sl][:, const], ref[:, const])              # constant columns and the like
        assert (~const).sum() >= 23


def test_noise_is_reproducible_and_independent():
    p = noisy_params(0, 50)
    env_config = np.zeros(2048, dtype=np.int64)
    a, b = engine([p], env_config), engine([p], env_config)
    a.set_forecast_noise(seed=7)
    b.set_forecast_noise(seed=7)
    oa, ob = a.observe().clone(), b.observe().clone()
    assert torch.equal(oa, ob)                                  # same seed, same call number
    assert not torch.equal(oa, a.observe())                     # next call draws again
    b.set_forecast_noise(seed=8)
    b._noise_calls = 0
    assert not torch.equal(oa, b.observe())
    sl = views.obs_slices(p, "gym_sorted")["load"]
    x = oa[:, sl][:, 1:].cpu().numpy()
    x = x - x.mean(0)
    corr = np.corrcoef(x.T)                                     # forecast rows are independent of each other ...
    assert np.abs(corr - np.eye(len(corr))).max() < 0.12
    assert abs(np.corrcoef(x[:-1, 0], x[1:, 0])[0, 1]) < 0.12    # ... and so are neighbouring envs


def test_f32_observations_take_noise_too():
    p = noisy_params(1, 10)
    env_config = np.zeros(256, dtype=np.int64)
    a = engine([p], env_config, obs_dtype=torch.float32)
    b = engine([p], env_config)
    a.set_forecast_noise(seed=3)
    b.set_forecast_noise(seed=3)
    oa, ob = a.observe(), b.observe()
    assert oa.dtype == torch.float32
    # f32 rows start from the rounded clean value, so they agree with the f64 path to f32 precision
    assert torch.allclose(oa.double(), ob, rtol=0, atol=2e-7)


@pytest.mark.parametrize("emit", ["auto", "image"])
def test_rollout_ring_carries_the_noise_of_each_step(emit):
    """`rollout` with Gaussian-noise forecasters: the persistent kernel writes the oracle rows and mg_forecast_noise_at gives
    every surviving ring slot the noise of ITS step (call numbers continue the per-step sequence) -- the ring equals, bit for
    bit, the rows that the same engine produces step by step with mg_step + mg_forecast_noise; also across the end of the
    series, where forecast rows turn into padding and carry no noise."""
    configs = [noisy_params(n, t0) for n in (0, 1, 2) for t0 in (0, 8740)]
    B, n_steps, ring = 300, 12, 5
    env_config = np.arange(B) % len(configs)
    a, b = engine(configs, env_config), engine(configs, env_config)
    for bm in (a, b):
        bm.set_forecast_noise(seed=77, env_offset=9)
        if emit == "image":
            bm.set_emit_image(True)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    acts = [torch.rand((n_steps, g.n_envs, g.n_act), dtype=torch.float64, device="cuda", generator=gen) for g in a.groups]
    out = a.rollout(acts, ring=ring)
    rows = []
    for k in range(n_steps):
        obs, _, _, _ = b.step([x[k].contiguous() for x in acts])
        rows.append([o.clone() for o in obs])
    noisy_somewhere = False
    for gi in range(len(a.groups)):
        for k in range(n_steps - ring, n_steps):
            assert torch.equal(out[gi]["obs_ring"][k % ring], rows[k][gi]), (gi, k)
        clean = engine([jump_to(load_pymgrid25(n), t0) for n in (0, 1, 2) for t0 in (0, 8740)], env_config)
    clean_out = clean.rollout(acts, ring=ring)
    for gi in range(len(a.groups)):
        noisy_somewhere |= not torch.equal(clean_out[gi]["obs_ring"], out[gi]["obs_ring"])
        assert torch.equal(clean_out[gi]["reward"], out[gi]["reward"])        # the physics never sees the noise
    assert noisy_somewhere and a._noise_calls == b._noise_calls == n_steps
