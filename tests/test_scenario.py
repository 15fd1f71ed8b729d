"""Scenario ingestion: bundled pymgrid25.npz == our reader of the reference's YAML/csv.gz == the live reference objects."""
import os

import numpy as np
import pytest

from pymgrid_b200.scenario import load_pymgrid25, read_reference_scenario, reference_scenario_path

DATA_ROOT = "/root/reference/src/pymgrid/data"


def test_bundle_has_25_scenarios_with_expected_architectures():
    archs = [load_pymgrid25(n).arch for n in range(25)]
    genset_only = {2, 3, 5, 7, 15, 17, 19, 20, 21, 23}
    grid_only = {0, 4, 6, 11, 12, 14, 16}
    for n, a in enumerate(archs):
        want = (1, 0, 23) if n in genset_only else (0, 1, 23) if n in grid_only else (1, 1, 23)
        assert a == want, n
    p = load_pymgrid25(0)
    assert len(p) == 8760 and p.final_step == 8759 and p.obs_dim == 146 and p.n_act == 2
    assert load_pymgrid25(1).obs_dim == 150 and load_pymgrid25(2).obs_dim == 54
    assert (p.load_ts <= 0).all() and (p.pv_ts >= 0).all()
    with pytest.raises(ValueError):
        load_pymgrid25(25)


@pytest.mark.reference
@pytest.mark.parametrize("n", range(25))
def test_bundle_matches_reference_files_and_objects(n):
    import warnings
    warnings.simplefilter("ignore")
    from oracle.from_reference import params_from_reference
    from oracle.ref_loader import load_reference
    load_reference()
    from pymgrid import Microgrid
    b = load_pymgrid25(n)
    f = read_reference_scenario(reference_scenario_path(n, DATA_ROOT))
    r = params_from_reference(Microgrid.from_scenario(n))
    for other in (f, r):
        np.testing.assert_array_equal(b.load_ts, other.load_ts)
        np.testing.assert_array_equal(b.pv_ts, other.pv_ts)
        assert (b.grid is None) == (other.grid is None) and (b.genset is None) == (other.genset is None)
        if b.grid is not None:
            np.testing.assert_array_equal(b.grid.time_series, other.grid.time_series)
            for k in ("max_import", "max_export", "cost_per_unit_co2"):
                assert getattr(b.grid, k) == getattr(other.grid, k)
        for k in ("min_capacity", "max_capacity", "max_charge", "max_discharge", "efficiency", "battery_cost_cycle",
                  "current_charge"):
            assert getattr(b.battery, k) == getattr(other.battery, k), k
        if b.genset is not None:
            for k in ("running_min_production", "running_max_production", "genset_cost", "co2_per_unit",
                      "cost_per_unit_co2", "start_up_time", "wind_down_time", "allow_abortion", "current_status",
                      "goal_status", "steps_until_up", "steps_until_down"):
                assert getattr(b.genset, k) == getattr(other.genset, k), k
        for k in ("loss_load_cost", "overgeneration_cost", "forecast_horizon", "initial_step", "final_step"):
            assert getattr(b, k) == getattr(other, k), k


@pytest.mark.reference
def test_oracle_matches_live_reference_random_steps():
    """Independent of the golden files: live reference vs oracle on fresh random actions."""
    import warnings
    warnings.simplefilter("ignore")
    from oracle.from_reference import params_from_reference
    from oracle.oracle import OracleGrid
    from oracle.ref_loader import load_reference
    load_reference()
    from pymgrid import Microgrid
    for n in (0, 1, 2, 9):
        m = Microgrid.from_scenario(n)
        o = OracleGrid(params_from_reference(m))
        rng = np.random.default_rng(42 + n)
        names = [k for k in ("genset", "battery", "grid") if hasattr(m.modules, k)]
        for _ in range(120):
            ctrl, flat = {}, []
            for k in names:
                a = rng.random(2) if k == "genset" else rng.random()
                ctrl[k] = [a]
                flat += list(np.atleast_1d(a))
            obs, r, d, _ = m.run(ctrl)
            oobs, orr, od, _, _ = o.run(flat)
            ref = np.concatenate([np.asarray(x).ravel() for k in ("battery", "genset", "grid", "load", "pv") if k in obs
                                  for x in obs[k]])
            assert r == orr and d == od
            np.testing.assert_array_equal(ref, oobs)
