"""Priority lists on composed microgrids: the Python oracle (oracle/compose.py) against DiscreteMicrogridEnv /
RuleBasedControl of the live reference (tests/golden/compose_discrete.npz, tests/golden/make_compose_discrete.py)."""
import json

import numpy as np
import pytest

from oracle.compose import ComposedOracle, oracle_priority_control, oracle_priority_lists, oracle_rbc_list
from tests.compose_cases import load_cases

CASES = load_cases("compose_discrete.npz")


@pytest.mark.parametrize("case", CASES, ids=[c.label for c in CASES])
def test_oracle_priority_lists_and_controls(case):
    orc = ComposedOracle(case.modules(), **case.microgrid_kwargs)
    for flag in (0, 1):
        want = [tuple(tuple(el) for el in pl) for pl in case.json(f"table_{flag}")]
        assert oracle_priority_lists(orc, bool(flag)) == want
    table = oracle_priority_lists(orc, False)
    widths = [(name, [2 if m.kind == "genset" else 1 for m in lst]) for name, lst in orc._of("controllable")]
    orc.reset()
    for k, a in enumerate(case["actions"]):
        control = oracle_priority_control(orc, table[int(a)])
        row = np.concatenate([np.atleast_1d(control[name][j]) for name, ws in widths for j in range(len(ws))])
        assert np.array_equal(row, case["controls"][k]), k
        obs, reward, done, info = orc.run(control, normalized=False)
        assert reward == case["rewards"][k] and done == bool(case["dones"][k]), k
        flat = np.concatenate([np.asarray(obs[m.name][m.index]).ravel() for m in orc.listing])
        assert np.array_equal(flat, case["obs"][k]), k
    # rule-based control
    orc = ComposedOracle(case.modules(), **case.microgrid_kwargs)
    pl = oracle_rbc_list(orc)
    assert [list(el) for el in pl] == case.json("rbc_list")
    orc.reset()
    got = []
    for _ in range(len(case["rbc_rewards"])):
        _, r, _, _ = orc.run(oracle_priority_control(orc, pl), normalized=False)
        got.append(r)
    assert np.array_equal(np.array(got), case["rbc_rewards"])
