"""Pin the C oracle (oracle/mg_oracle.c) to the reference: golden vectors recorded from the live reference
(tests/golden/make_golden.py) must be reproduced BIT FOR BIT (integers and floats alike)."""
import numpy as np
import pytest

from oracle.oracle import MOD_BATTERY, MOD_GENSET, MOD_GRID, OracleBatch, OracleGrid  # noqa: F401
from pymgrid_b200.scenario import load_pymgrid25
from tests.helpers import custom_params, fuzz_params, jump_to, overfull_params, state_from_oracle


def run_and_compare(o, actions, rewards, dones, obs, infos, states, normalized=True):
    for k in range(len(actions)):
        ob, r, d, info, _ = o.run(actions[k], normalized=normalized)
        assert r == rewards[k], (k, r, rewards[k])
        assert d == bool(dones[k]), k
        np.testing.assert_array_equal(ob, obs[k], err_msg=f"obs step {k}")
        np.testing.assert_array_equal(info, infos[k], err_msg=f"info step {k}")
        np.testing.assert_array_equal(state_from_oracle(o), states[k], err_msg=f"state step {k}")


@pytest.mark.parametrize("n", range(25))
def test_pymgrid25_steps(golden, n):
    z = golden["pymgrid25_steps"]
    p = load_pymgrid25(n)
    o = OracleGrid(p)
    g = lambda k: z[f"s{n}_{k}"]  # noqa: E731
    run_and_compare(o, g("a0"), g("r0"), g("d0"), g("o0"), g("i0"), g("s0"))
    run_and_compare(o, g("au"), g("ru"), g("du"), g("ou"), g("iu"), g("su"), normalized=False)
    # across the end of the series: forecast padding, done at final_step - 1, last valid step t = T - 1
    o2 = OracleGrid(jump_to(p, int(z["end_from"])))
    np.testing.assert_array_equal(o2.reset(), g("reset_obs"))
    run_and_compare(o2, g("a1"), g("r1"), g("d1"), g("o1"), g("i1"), g("s1"))
    assert g("d1")[-1] == 1 and g("d1")[0] == 0


@pytest.mark.parametrize("n", (0, 1, 2))
def test_full_year(golden, n):
    z = golden["pymgrid25_year"]
    p = load_pymgrid25(n)
    actions = np.random.default_rng(0).random((8760, p.n_act))
    b = OracleBatch([p])
    rewards, dones, _ = b.rollout(actions[:, None, :])
    np.testing.assert_array_equal(rewards[:, 0], z[f"s{n}_rewards"])
    assert int(np.argmax(dones[:, 0])) == int(z[f"s{n}_first_done"]) == 8758
    t, charge, gen = b.state()
    np.testing.assert_array_equal(np.array([t[0], charge[0], *gen[0]], dtype=np.float64), z[f"s{n}_final_state"])


def test_legacy_seed_known_answer(golden):
    """np.random.seed(0); sample_action(strict_bound=True) on scenario 0."""
    z = golden["pymgrid25_year"]
    o = OracleGrid(load_pymgrid25(0))
    _, r, _, info, _ = o.run(z["legacy_s0_action"])
    assert r == float(z["legacy_s0_reward"]) == -544.8242524518755
    np.testing.assert_array_equal(info, z["legacy_s0_info"])
    assert info[0] == 304.403798960378 and info[9] == 826.3271668700909 and info[4] == 339.9448144937339


def test_stepping_past_the_end_is_flagged():
    from oracle.oracle import lib  # noqa: F401
    p = jump_to(load_pymgrid25(0), 8759)
    o = OracleGrid(p)
    _, r, d, _, err = o.run(np.array([0.5, 0.5]))
    assert d and err & 0x7f == 0 and o.state["t"] == 8760
    _, r, d, _, err = o.run(np.array([0.5, 0.5]))
    assert err & (1 << 5) and np.isnan(r) and o.state["t"] == 8760


def test_genset_state_machine(golden):
    """Exhaustive: U, D in 0..4, both abortion settings, both initial states, all 6-step goal sequences
    (the reference's TestManyStatusChanges, tests/microgrid/modules/module_tests/test_genset_long_status_changes.py:217-262)."""
    from oracle.oracle import lib, OrcGrid
    import ctypes as C
    z = golden["genset_machine"]
    params, goals, states = z["params"], z["goals"], z["states"]
    L = lib()
    for (U, D, abort, init), gl, st in zip(params, goals, states):
        g = OrcGrid()
        g.start_up_time, g.wind_down_time, g.allow_abortion = int(U), int(D), int(abort)
        g.cs = g.gs = int(init)
        g.up, g.dn = (0, int(D)) if init else (int(U), 0)
        for goal, want in zip(gl, st):
            pred = L.orc_genset_next_status(C.byref(g), int(goal))
            L.orc_genset_update_status(C.byref(g), float(goal))
            assert (g.cs, g.gs, g.up, g.dn, pred) == tuple(int(x) for x in want)
    # Python round(): half to even
    g = OrcGrid()
    g.cs = g.gs = 1
    for f, want in zip(z["frac_goals"], z["frac_status"]):
        L.orc_genset_update_status(C.byref(g), float(f))
        assert g.cs == int(want)


def test_genset_known_answers_from_reference_tests():
    """Literal state tuples from the reference's unit tests:
    test_genset_long_status_changes.py:35-89 (on -> off with wind_down_time=3) and :180-214 (abortion)."""
    from oracle.oracle import lib, OrcGrid
    import ctypes as C
    L = lib()
    g = OrcGrid()
    g.start_up_time, g.wind_down_time, g.allow_abortion = 2, 3, 1
    g.cs = g.gs = 1
    g.up, g.dn = 0, 3
    seq = []
    for _ in range(4):
        L.orc_genset_update_status(C.byref(g), 0.0)
        seq.append((g.cs, g.gs, g.up, g.dn))
    assert seq == [(1, 0, 0, 2), (1, 0, 0, 1), (1, 0, 0, 0), (0, 0, 2, 0)]
    # abort a shut-down half way: goal back to 1 restores the running state
    g.cs = g.gs = 1
    g.up, g.dn = 0, 3
    L.orc_genset_update_status(C.byref(g), 0.0)
    L.orc_genset_update_status(C.byref(g), 1.0)
    assert (g.cs, g.gs, g.up, g.dn) == (1, 1, 0, 3)


@pytest.mark.parametrize("i", range(6))
def test_custom_grids(golden, i):
    z = golden["custom"]
    p = custom_params(z, i)
    o = OracleGrid(p)
    np.testing.assert_array_equal(o.reset(), z[f"c{i}_reset_obs"])
    run_and_compare(o, z[f"c{i}_a"], z[f"c{i}_r"], z[f"c{i}_d"], z[f"c{i}_o"], z[f"c{i}_i"], z[f"c{i}_s"])
    # reset keeps battery charge and genset status
    charge = o.state["charge"]
    np.testing.assert_array_equal(o.reset(), z[f"c{i}_after_reset_obs"])
    assert o.state["charge"] == charge and o.state["t"] == 0


@pytest.mark.parametrize("i", range(40))
def test_fuzz_grids(golden, i):
    """Randomised constructor arguments (tests/golden/make_fuzz.py): continuous steps (normalised, then unnormalised
    beyond every limit), the discrete env's priority-list expansion with slow gensets, and RuleBasedControl's sorted list."""
    from pymgrid_b200 import priority_list as PL
    z = golden["fuzz"]
    g = lambda k: z[f"f{i}_{k}"]  # noqa: E731
    p = fuzz_params(z, i)
    o = OracleGrid(p)
    np.testing.assert_array_equal(o.reset(), g("reset_obs"))
    for seg, normalized in (("n", True), ("u", False)):
        n = len(g(f"{seg}_r"))
        run_and_compare(o, g(f"{seg}_a")[:n], g(f"{seg}_r"), g(f"{seg}_d"), g(f"{seg}_o"), g(f"{seg}_i"), g(f"{seg}_s"),
                        normalized=normalized)
        err = int(g(f"{seg}_err"))
        if err >= 0:     # the reference raised AssertionError (base_module.py:272) on this action
            assert err == n
            _, _, _, _, flags = o.run(g(f"{seg}_a")[err], normalized=normalized)
            assert flags & (1 << 4), flags
            break
    # discrete env: table built on the host, expansion by the oracle
    o = OracleGrid(fuzz_params(z, i))
    np.testing.assert_array_equal(o.reset(), g("d_reset_obs"))
    table = PL.priority_lists(p.has_genset, p.has_grid, p.genset.running_min_production if p.has_genset else None)
    mod, act = g("d_table_mod"), g("d_table_act")
    assert len(table) == len(mod)
    for row, pl in enumerate(table):
        assert [(int(m), int(a)) for m, a in zip(mod[row], act[row]) if m >= 0] == list(pl)
    for k, a in enumerate(g("d_actions")):
        ctrl = o.priority_control(table[a])
        np.testing.assert_array_equal(ctrl, g("d_controls")[k], err_msg=f"control {k}")
        ob, r, d, _, _ = o.run(ctrl, normalized=False)
        assert r == g("d_rewards")[k] and d == bool(g("d_dones")[k])
        np.testing.assert_array_equal(ob, g("d_obs")[k])
    np.testing.assert_array_equal(state_from_oracle(o), g("d_state"))
    # rule based control: the automatically sorted list, then its whole run
    o = OracleGrid(fuzz_params(z, i))
    o.reset()
    pl = PL.rbc_priority_list(p)
    assert [m for m, _ in pl] == list(g("rbc_list_mod")) and [a for _, a in pl] == list(g("rbc_list_act"))
    rewards = []
    for _ in range(len(g("rbc_rewards"))):
        _, r, _, _, _ = o.run(o.priority_control(pl), normalized=False)
        rewards.append(r)
    np.testing.assert_array_equal(np.array(rewards), g("rbc_rewards"))
    np.testing.assert_array_equal(state_from_oracle(o), g("rbc_final_state"))


def test_overfull_battery_is_flagged_where_the_reference_asserts(golden):
    """charge one ulp above max_capacity (reachable by rounding): the reference refuses a continuous charge request
    (AssertionError base_module.py:272) and any priority list that reaches the battery while there is surplus energy
    (AssertionError priority_list.py:124); both set the NEGATIVE_ABSORB flag here.  Discharging and lists that never make
    the battery absorb run normally (tests/golden/fuzz.npz over_*, make_fuzz.overfull_battery)."""
    from pymgrid_b200 import priority_list as PL
    z = golden["fuzz"]
    assert str(z["over_continuous_charge_raised"]) == "base_module.py:272"
    o = OracleGrid(overfull_params(z))
    _, _, _, _, flags = o.run(np.array([-10.0, 0.0]), normalized=False)
    assert flags & (1 << 4)
    o = OracleGrid(overfull_params(z))
    ob, r, _, _, flags = o.run(np.array([5.0, -9.0]), normalized=False)
    assert flags & 0x7f == 0 and r == float(z["over_discharge_reward"])
    np.testing.assert_array_equal(ob, z["over_discharge_obs"])
    np.testing.assert_array_equal(state_from_oracle(o), z["over_discharge_state"])
    table = PL.priority_lists(False, True)
    assert [list(pl) for pl in table] == [[(int(m), int(a)) for m, a in zip(mr, ar) if m >= 0]
                                          for mr, ar in zip(z["over_table_mod"], z["over_table_act"])]
    for a, pl in enumerate(table):
        o = OracleGrid(overfull_params(z))
        ctrl = o.priority_control(pl)
        raised = str(z["over_discrete_raised"][a])
        assert bool(o.list_flags & (1 << 4)) == (raised == "priority_list.py:124"), (a, raised)
        if not raised:
            _, r, _, _, flags = o.run(ctrl, normalized=False)
            assert flags & 0x7f == 0 and r == z["over_discrete_reward"][a]


def _discrete_cases(z):
    return sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_actions")})


def test_discrete_env(golden):
    z = golden["discrete"]
    cases = _discrete_cases(z)
    assert len(cases) == 40
    for tag in cases:
        horizon, n = int(tag.split("_")[0][1:]), int(tag.split("_")[1][1:])
        p = load_pymgrid25(n)
        p.forecast_horizon = horizon
        o = OracleGrid(p)
        assert o.obs_dim == int(z[f"{tag}_obs_dim"])
        np.testing.assert_array_equal(o.reset(), z[f"{tag}_reset_obs"])
        mod, act = z[f"{tag}_table_mod"], z[f"{tag}_table_act"]
        for k, a in enumerate(z[f"{tag}_actions"]):
            plist = [(int(m), int(x)) for m, x in zip(mod[a], act[a]) if m >= 0]
            ctrl = o.priority_control(plist)
            np.testing.assert_array_equal(ctrl, z[f"{tag}_controls"][k], err_msg=f"{tag} control {k}")
            ob, r, d, _, _ = o.run(ctrl, normalized=False)
            assert r == z[f"{tag}_rewards"][k] and d == bool(z[f"{tag}_dones"][k])
            np.testing.assert_array_equal(ob, z[f"{tag}_obs"][k], err_msg=f"{tag} obs {k}")


def test_container_order_is_a_permutation_of_sorted_order():
    p = load_pymgrid25(1)
    a, b = OracleGrid(p, order=0), OracleGrid(p, order=1)
    oa, ob = a.observe(), b.observe()
    rows = 1 + p.forecast_horizon
    bat, gen, grid, load, pv = np.split(oa, np.cumsum([2, 4, 4 * rows, rows]))
    np.testing.assert_array_equal(ob, np.concatenate([load, pv, gen, bat, grid]))


SHAPER_FLAG = 1 << 7


def shaped_segments(z):
    for n in (0, 1, 2, 13):
        for tag, name in (("pv", "pv_curtailment"), ("bat", "battery_discharge")):
            for seg in range(int(z["n_seg"])):
                yield n, name, f"s{n}_{tag}_{seg}"


def shaped_params(n, name, t0):
    p = jump_to(load_pymgrid25(n), int(t0))
    p.reward_shaper = name
    return p


def sequential_module_reward(info):
    """the unshaped step reward the balance log keeps: module rewards added in dispatch order (utils/step.py:18)"""
    r = 0.0
    for col in (12, 13, 14, 15):
        r += info[col]
    return r


def test_reward_shapers(golden):
    """Microgrid(reward_shaping_func=PVCurtailmentShaper / BatteryDischargeShaper): shaped reward bit for bit, the
    unshaped sum still available, and the shaper's assert (mid-step or final) reported as a flag on the same step."""
    z = golden["shaped"]
    n_rows = n_raised = 0
    for n, name, key in shaped_segments(z):
        o = OracleGrid(shaped_params(n, name, z[key + "_t0"]))
        a, r, s, raised = z[key + "_a"], z[key + "_r"], z[key + "_s"], int(z[key + "_raised"])
        for k in range(len(r)):
            _, rew, _, info, err = o.run(a[k])
            assert rew == r[k] and not err & SHAPER_FLAG, (key, k)
            assert rew == z[key + "_log_shaped"][k] and sequential_module_reward(info) == z[key + "_log_reward"][k]
            np.testing.assert_array_equal(state_from_oracle(o), s[k])
        n_rows += len(r)
        if raised >= 0:
            assert raised == len(r)
            assert o.run(a[raised])[4] & SHAPER_FLAG, (key, raised)
            n_raised += 1
    assert n_rows > 1000 and n_raised >= 10


@pytest.mark.parametrize("n", (0, 1, 2, 13))
def test_battery_discharge_shaper_under_priority_lists(golden, n):
    z = golden["shaped"]
    key = f"s{n}_batd"
    o = OracleGrid(shaped_params(n, "battery_discharge", z[key + "_t0"]))
    assert int(z[key + "_raised"]) == -1
    for k, ctrl in enumerate(z[key + "_controls"]):
        _, rew, _, info, err = o.run(ctrl, normalized=False)
        assert rew == z[key + "_r"][k] and not err & SHAPER_FLAG, k
        assert sequential_module_reward(info) == z[key + "_log_reward"][k]
        np.testing.assert_array_equal(state_from_oracle(o), z[key + "_s"][k])
