"""Host-side conversions (pymgrid_b200/views.py) against the reference's recorded `get_log()` / `state_series()` and,
when the reference checkout is present, against its live `run()` dicts.  The step outputs fed to the conversions
come from the C oracle here (CPU suite); tests/test_gpu_dropin.py feeds them from the engine."""
import numpy as np
import pytest

from oracle.oracle import OracleGrid
from pymgrid_b200 import views
from pymgrid_b200.scenario import load_pymgrid25


@pytest.mark.parametrize("n", (0, 1, 2))
def test_log_rows_match_reference_get_log(golden, n):
    z = golden["log"]
    p = load_pymgrid25(n)
    o = OracleGrid(p)
    rows = []
    for a in z[f"s{n}_actions"]:
        st = o.state
        pre = views.state_dict(p, st["t"], st["charge"], st["genset"])
        _, r, _, info, _ = o.run(a)
        rows.append(views.log_row(p, pre, info, r, o.state["genset"]))
    cols = ["|".join(map(str, c)) for c in rows[0].keys()]
    assert cols == list(z[f"s{n}_columns"])
    vals = np.array([[float(v) for v in row.values()] for row in rows])
    np.testing.assert_array_equal(vals, z[f"s{n}_values"])
    # state_series() after the 30 steps
    st = o.state
    sd = views.state_dict(p, st["t"], st["charge"], st["genset"])
    flat = [(f"{name}|0|{k}", float(v)) for name, d in sd.items() for k, v in d.items()]
    assert [k for k, _ in flat] == list(z[f"s{n}_state_series_index"])
    np.testing.assert_array_equal(np.array([v for _, v in flat]), z[f"s{n}_state_series_values"])


def test_control_dict_conversion():
    p = load_pymgrid25(1)
    cols = {"genset": 0, "battery": 2, "grid": 3}
    row = views.control_dict_to_row({"genset": [np.array([1.0, 0.25])], "battery": [0.5], "grid": 0.75}, p, cols)
    np.testing.assert_array_equal(row, [1.0, 0.25, 0.5, 0.75])
    with pytest.raises(ValueError):
        views.control_dict_to_row({"genset": [[1.0, 0.2]], "grid": [0.1]}, p, cols)
    with pytest.warns(UserWarning):
        views.control_dict_to_row({"genset": [[1.0, 0.2]], "battery": [0.1], "grid": [0.1], "pv": [0.0]}, p, cols)


@pytest.mark.reference
def test_dict_views_match_live_reference():
    import warnings
    warnings.simplefilter("ignore")
    from oracle.ref_loader import load_reference
    load_reference()
    from pymgrid import Microgrid
    for n in (0, 1, 2):
        m = Microgrid.from_scenario(n)
        p = load_pymgrid25(n)
        o = OracleGrid(p)
        rng = np.random.default_rng(n)
        cols, c = {}, 0
        for name in views.control_names(p):
            cols[name] = c
            c += 2 if name == "genset" else 1
        for _ in range(40):
            ctrl = {name: [rng.random(2) if name == "genset" else rng.random()] for name in views.control_names(p)}
            obs, r, d, info = m.run(ctrl)
            row = views.control_dict_to_row(ctrl, p, cols)
            oobs, orr, od, oinfo, flags = o.run(row)
            mine_obs = views.obs_row_to_dict(oobs, p)
            assert list(mine_obs.keys()) == list(obs.keys())
            for k in obs:
                np.testing.assert_array_equal(np.asarray(obs[k][0], dtype=float), mine_obs[k][0])
            mine_info = views.info_row_to_dict(oinfo, flags, p)
            assert list(mine_info.keys()) == list(info.keys())
            for k in info:
                assert {kk: float(vv) for kk, vv in info[k][0].items()} == mine_info[k][0], k
            assert r == orr and d == od


def test_trajectory_rules():
    """reference: tests/envs/test_trajectory.py -- windows inside [initial, final), fixed length honoured."""
    from pymgrid_b200 import trajectory as tr
    rng = np.random.default_rng(0)
    i, f = tr.StochasticTrajectory()(0, 8759, n=5000, rng=rng)
    assert (i >= 0).all() and (i <= 8759 - 3).all() and (f > i).all() and (f <= 8758).all()
    i, f = tr.FixedLengthStochasticTrajectory(24)(10, 500, n=2000, rng=rng)
    assert (f - i == 24).all() and (i >= 10).all() and (f <= 500).all()
    with pytest.raises(ValueError):
        tr.FixedLengthStochasticTrajectory(600)(10, 500, n=4)
    i, f = tr.DeterministicTrajectory(7, 99)(0, 8759, n=3)
    assert i.tolist() == [7, 7, 7] and f.tolist() == [99, 99, 99]
    assert tr.DeterministicTrajectory(7, 99)(0, 8759) == (7, 99)
