"""Host-side checks of the forecast-noise plumbing that need no GPU: the Philox restatement the GPU test compares the
kernel with (known-answer vectors of the published algorithm), and the per-config MgForecastNoise records against the
reference-pinned numpy restatement of GaussianNoiseForecaster."""
import numpy as np

from oracle.forecast_noise import NoisyModule
from pymgrid_b200.params import ForecasterParams
from pymgrid_b200.scenario import load_pymgrid25
from tests.helpers import engine_noise_normals, philox4x32_10


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10"""
    kat = (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)))
    for ctr, key, out in kat:
        assert tuple(int(v) for v in philox4x32_10(np.array(ctr, dtype=np.uint64), key)) == out


def test_generator_normals_are_standard_normal():
    z = np.concatenate([engine_noise_normals(env, 17, 72, seed=5, call=1) for env in range(400)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert abs(np.mean(z ** 3)) < 0.05 and abs(np.mean(z ** 4) - 3) < 0.15
    assert not np.array_equal(engine_noise_normals(0, 17, 8, 5, 1), engine_noise_normals(0, 17, 8, 5, 2))
    assert not np.array_equal(engine_noise_normals(0, 17, 8, 5, 1), engine_noise_normals(1, 17, 8, 5, 1))


def test_noise_records_match_the_reference_restatement():
    from pymgrid_b200.engine import forecast_noise_record
    p = load_pymgrid25(1)
    p.forecasters = dict(load=ForecasterParams(0.2, False, True), pv=ForecasterParams(0.3, True, True),
                         grid=ForecasterParams(0.02, False, True))
    rec = forecast_noise_record(p)
    mods = dict(load=NoisyModule(p.load_ts, 23, True, 0.2, False, True, p.initial_step, p.final_step),
                pv=NoisyModule(p.pv_ts, 23, True, 0.3, True, True, p.initial_step, p.final_step),
                grid=NoisyModule(p.grid.time_series, 23, False, 0.02, False, True, p.initial_step, p.final_step))
    assert rec.load_sigma == mods["load"].sigma_normalised()[0, 0] and rec.load_increase == 0
    assert rec.pv_sigma == mods["pv"].sigma_normalised()[0, 0] and rec.pv_increase == 1
    np.testing.assert_array_equal(np.array(rec.grid_sigma[:]), mods["grid"].sigma_normalised()[0])
    assert rec.grid_sigma[1] == 0.0       # export price is constant 0: the forecaster's clip pins it
    # with increase_uncertainty the table grows like 1 + log(1 + k)
    np.testing.assert_allclose(mods["pv"].sigma_normalised()[:, 0], rec.pv_sigma * (1 + np.log(1 + np.arange(23))), rtol=1e-15)
