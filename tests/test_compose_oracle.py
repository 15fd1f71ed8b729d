"""The Python restatement of Microgrid.run for ANY module composition (oracle/compose.py) against what the live
reference returned (tests/golden/compose.npz, recorded by tests/golden/make_compose.py): bit for bit."""
import numpy as np
import pytest

from oracle.compose import ComposedOracle, Raised, np_sum
from tests.compose_cases import BALANCE_COLS, load_cases

CASES = load_cases()


def controllable_widths(orc):
    return [(name, [2 if m.kind == "genset" else 1 for m in lst]) for name, lst in orc._of("controllable")]


def test_np_sum_matches_numpy():
    rng = np.random.default_rng(5)
    for n in list(range(0, 40)) + [64, 100, 128]:
        for _ in range(50):
            a = list(10 * rng.random(n) - 3)
            assert np_sum(a) == float(np.sum(a)) or n == 0


@pytest.mark.parametrize("case", CASES, ids=[c.label for c in CASES])
def test_oracle_reproduces_reference(case):
    orc = ComposedOracle(case.modules(), **case.microgrid_kwargs, **case.callable_kwargs)
    assert [(m.name, m.index) for m in orc.listing] == [(n, j) for n, j, _ in case.names]
    reset = orc.reset()
    assert list(reset.keys()) == [k for k in case.json("reset_keys") if k not in ("balance", "other")]
    flat = lambda obs: np.concatenate([np.asarray(obs[m.name][m.index]).ravel() for m in orc.listing] + [np.zeros(0)])  # noqa: E731
    assert np.array_equal(flat(reset), case["obs_reset"])
    n = len(case["rewards"])
    widths = controllable_widths(orc)
    reset_at, n_resets = list(case.spec.get("reset_at", [])), 0
    for k in range(n):
        if k in reset_at:
            assert np.array_equal(flat(orc.reset()), case["reset_obs_rows"][n_resets]), k
            n_resets += 1
        obs, reward, done, info = orc.run(case.control(k, widths), normalized=bool(case["normalized"][k]))
        assert orc.current_step == int(case["steps_after"][k]), k
        assert list(obs.keys()) == case.json("run_keys")
        assert reward == case["rewards"][k], k
        assert done == bool(case["dones"][k]), k
        assert np.array_equal(flat(obs), case["obs"][k]), k
        for i, m in enumerate(orc.listing):
            inf, want = info[m.name][m.index], case["info"][k, i]
            assert inf.get("provided_energy", 0.0) == want[0] and inf.get("absorbed_energy", 0.0) == want[1], (k, m.name)
            assert inf.get("co2_production", inf.get("curtailment", 0.0)) == want[2]
            assert ("absorbed_energy" in inf) == bool(want[4])
            assert orc.log_rows[-1][(m.name, m.index, "reward")] == want[3]
        state = []
        for m in orc.listing:
            if m.kind == "battery":
                state += [m.charge, m.soc]
            elif m.kind == "genset":
                state += [m.cs, m.gs, m.up, m.dn]
        assert np.array_equal(np.array(state, dtype=np.float64), case["states"][k]), k
    raised_at = int(case["raised_at"])
    if raised_at >= 0:
        # the step after the last recorded one is where the reference raised
        with pytest.raises(Raised) as exc:
            # any in-range action: the recorded cases raise independently of the action (no slack module / end of the data)
            ctrl = {name: [np.array([0.5, 0.5]) if w == 2 else 0.5 for w in ws] for name, ws in widths}
            orc.run(ctrl, normalized=True)
        assert exc.value.kind == str(case["raised_type"])
    assert orc.current_step == int(case["current_step"])
    if str(case["log_raises"]):
        return          # the reference's own get_log() fails under a trajectory_func (see tests/golden/make_compose.py)
    # the log frame: same columns in the same order, same values
    cols = [tuple(c) for c in case.json("log_columns")]
    rows = orc.log_rows
    assert len(rows) == len(case["log_values"])
    if rows:
        assert list(rows[0].keys()) == cols
        got = np.array([[float(v) for v in r.values()] for r in rows])
        assert np.array_equal(got, case["log_values"], equal_nan=True)
        bal = np.array([[r[("balance", 0, c)] for c in BALANCE_COLS] for r in rows[:n]])
        assert np.array_equal(bal, case["balance"][:n])
    assert orc.current_step == int(case["current_step"])
