"""MicrogridGenerator grids (BASELINE config 5): the (profile, scale) representation against REAL generator grids
recorded from the reference (tests/golden/generator.npz), CPU side through the oracle."""
import numpy as np

from oracle.oracle import OracleGrid
from pymgrid_b200 import generator
from tests.helpers import generator_params


def pv_first(obs_sorted_lower, p):
    """oracle 'gym_sorted' order (battery, genset, grid, load, pv) -> the order with the renewable module named 'PV'."""
    rows = 1 + p.forecast_horizon
    n_state = 2 + 4 * p.has_genset + 4 * rows * p.has_grid
    head, load, pv = np.split(obs_sorted_lower, [n_state, n_state + rows])
    return np.concatenate([pv, head, load])


def test_real_generator_grids_in_profile_scale_form(golden):
    z = golden["generator"]
    n = int(z["n"])
    assert n == 24
    archs = set()
    for i in range(n):
        p = generator_params(z, i)
        archs.add((p.has_genset, p.has_grid))
        assert p.scaled and len(p) == 8760 and p.final_step == 8760 and p.forecast_horizon == 23
        o = OracleGrid(p)
        for k, a in enumerate(z[f"g{i}_a"]):
            obs, r, d, _, _ = o.run(a)
            assert r == z[f"g{i}_r"][k] and d == bool(z[f"g{i}_d"][k]), (i, k)
            np.testing.assert_array_equal(pv_first(obs, p), z[f"g{i}_o"][k], err_msg=f"grid {i} step {k}")
        st = o.state
        np.testing.assert_array_equal(np.array([st["t"], st["charge"], *st["genset"]], dtype=float), z[f"g{i}_s"][-1])
    assert archs == {(True, False), (False, True), (True, True)}


def test_sampler_distributions_and_explicit_form():
    gb = generator.sample(4000, seed=3)
    assert abs(gb.has_grid.mean() - 0.67) < 0.03 and abs((~gb.has_grid).mean() - 0.33) < 0.03
    assert (gb.has_genset | gb.has_grid).all() and (gb.has_genset[gb.grid_weak]).all()
    assert abs(gb.grid_weak[gb.has_grid].mean() - 0.5) < 0.04
    assert gb.status[~gb.grid_weak].min() == 1 and gb.status[gb.grid_weak].mean() < 1
    assert (gb.bat_power == np.ceil(gb.bat_capacity / 4)).all() and ((gb.bat_soc0 >= 0.2) & (gb.bat_soc0 <= 1)).all()
    peak = gb.profiles["load"].max(axis=1)[gb.load_profile] * gb.load_scale
    assert (np.abs(peak - np.round(peak)) < 1e-6).all() and peak.min() >= 100 and peak.max() <= 100000
    p = gb.to_params(int(np.nonzero(gb.has_grid & gb.has_genset)[0][0]))
    assert p.arch == (1, 1, 23) and p.renewable_name == "PV" and p.grid.status is not None
    o = OracleGrid(p)
    obs, r, d, _, flags = o.run(np.array([1.0, 0.5, 0.5, 0.5]))
    assert np.isfinite(r) and not d and (obs >= 0).all() and (obs <= 1).all()


def test_vectorised_config_records_equal_the_scalar_builder():
    """generator.batch_config_records (numpy, one record per env) == engine.config_record(to_params(i)) field for field."""
    from pymgrid_b200 import _cabi
    from pymgrid_b200.engine import config_record
    gb = generator.sample(300, seed=9)
    cfg, tables, plist_rows, load_tab, pv_tab, grid_np = generator.batch_config_records(gb)
    n_co2 = gb.profiles["co2"].shape[0]
    skip = {"reserved", "plist_offset"}
    for i in range(gb.n):
        p = gb.to_params(i)
        grid_series = (int(gb.tariff[i]) - 1) * n_co2 + int(gb.co2_profile[i]) if p.has_grid else 0
        want = config_record(p, int(gb.load_profile[i]), int(gb.pv_profile[i]), grid_series, 0, len(tables[(int(p.has_genset), int(p.has_grid))]))
        for name, _ in _cabi.MgConfig._fields_:
            if name in skip:
                continue
            assert cfg[name][i] == getattr(want, name), (i, name, cfg[name][i], getattr(want, name))
        if p.has_grid:
            np.testing.assert_array_equal(grid_np[grid_series], p.grid.time_series)
        np.testing.assert_array_equal(load_tab[gb.load_profile[i]], p.load_ts)
        np.testing.assert_array_equal(pv_tab[gb.pv_profile[i]], p.pv_ts)
