"""What the reference's API / env / controller tests pin on its shared fixture grid (tests/helpers/modular_microgrid.py:14-55:
genset 10-50 at cost 0.5, battery 0-100 with 50 / 50 power and efficiency 1 at soc 0.5, PV == 50, load == 60, grid import
100 / export 0 at price 1 with raise_errors=True) -- tests/microgrid/test_microgrid.py:12-186, tests/envs/test_trajectory.py,
tests/control/test_rbc.py -- written for this package: the facts are the reference's, the code is ours (plain pytest
functions over one fixture builder).  Collected twice: on the CPU with the engine replaced by the oracle-backed stand-in
(tests/test_reference_suite_fused_host.py) and on the GPU (tests/test_zz_gpu_dropin_more.py).

Not covered: sampling from gym space objects (test_action_space*), object equality and YAML round trips of the reference
classes (TestRBC.test_init, test_trajectory_serialization).
"""
import numpy as np
import pytest

import pymgrid_b200
from pymgrid_b200.algos import RuleBasedControl
from pymgrid_b200.envs import DiscreteMicrogridEnv
from pymgrid_b200.modules import BatteryModule, GensetModule, GridModule, LoadModule, RenewableModule

T = 100


def fixture_modules(length=T):
    return [GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5),
            BatteryModule(min_capacity=0, max_capacity=100, max_charge=50, max_discharge=50, efficiency=1.0, init_soc=0.5),
            RenewableModule(time_series=np.full(length, 50.0)), LoadModule(time_series=np.full(length, 60.0)),
            GridModule(max_import=100, max_export=0, time_series=np.ones((length, 3)), raise_errors=True)]


def fixture_microgrid():
    return pymgrid_b200.Microgrid(fixture_modules())


def windowed_env(trajectory_func):
    return DiscreteMicrogridEnv(fixture_modules(), trajectory_func=trajectory_func)


# ---- Microgrid surface (test_microgrid.py:12-186) ----------------------------------------------------------------------
def test_every_scenario_exposes_its_modules_as_attributes():
    for n in range(25):
        m = pymgrid_b200.Microgrid.from_scenario(n)
        assert all(hasattr(m, name) for name in ("load", "pv", "battery")) and (hasattr(m, "grid") or hasattr(m, "genset")), n


def test_action_templates_cover_exactly_the_controllable_modules():
    m = fixture_microgrid()
    empty = m.get_empty_action()
    assert set(empty) == {"battery", "genset", "grid"} and all(v == [None] for v in empty.values())
    sampled = m.sample_action()
    assert set(sampled) == set(empty)
    for name, values in sampled.items():
        for k, value in enumerate(values):
            arr = np.atleast_1d(np.asarray(value, dtype=float))
            assert arr.shape[0] == m.modules[name][k].action_space.shape[0] and ((0 <= arr) & (arr <= 1)).all()
    for name, modules in m.fixed.iterdict():          # fixed modules have no action space and get no action
        for module in modules:
            assert module.action_space.shape == (0,) and name not in sampled


def test_flex_modules_are_sampled_on_request():
    m = fixture_microgrid()
    with_flex = m.sample_action(sample_flex_modules=True)          # test_microgrid.py:230-235 on its load + PV grids
    assert list(with_flex) == ["renewable", "balancing", "genset", "battery", "grid"]
    assert set(m.get_empty_action(sample_flex_modules=True)) == set(with_flex) and "load" not in with_flex
    np.random.seed(11)
    a = m.sample_action()
    np.random.seed(11)
    b = m.sample_action()
    assert all(np.array_equal(np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float)) for k in a)      # numpy's global generator


def test_step_counter_and_reset():
    m = fixture_microgrid()
    assert m.current_step == 0
    for k in range(4):
        m.run(m.sample_action())
        assert m.current_step == k + 1
    m.reset()
    assert m.current_step == 0
    m.initial_step = 1
    assert m.initial_step == 1
    m.reset()
    assert m.current_step == 1


def test_set_module_attr_and_cost_info():
    m = fixture_microgrid()
    m.set_module_attr("forecast_horizon", 50)
    horizons = {x.forecast_horizon for x in m.modules.iterlist() if hasattr(x, "forecast_horizon")}
    assert horizons == {50}
    with pytest.raises(AttributeError):
        m.set_module_attr("blah", "blah")
    info = fixture_microgrid().get_cost_info()
    for name in ("genset", "battery", "renewable", "load", "grid", "balancing"):
        assert len(info[name]) == 1 and set(info[name][0]) == {"production_marginal_cost", "absorption_marginal_cost"}
        assert all(isinstance(float(v), float) for v in info[name][0].values())


# ---- episode windows (test_trajectory.py) ------------------------------------------------------------------------------
def module_window(env):
    return (env.modules.get_attrs("initial_step", unique=True).item(), env.modules.get_attrs("final_step", unique=True).item())


@pytest.mark.parametrize("func, window", [(None, (0, T)), (lambda lo, hi: (10, 20), (10, 20))])
def test_trajectory_moves_the_modules_window_not_the_envs(func, window):
    env = windowed_env(func)
    assert (env.initial_step, env.final_step) == (0, T)
    env.reset()
    assert (env.initial_step, env.final_step) == (0, T) and module_window(env) == window


def test_random_trajectory_stays_inside_the_series():
    def draw(lo, hi):
        a = int(np.random.randint(lo + 1, hi - 2))
        return a, int(np.random.randint(a + 1, hi))
    env = windowed_env(draw)
    env.reset()
    lo, hi = module_window(env)
    assert (env.initial_step, env.final_step) == (0, T) and 0 < lo < hi < T


@pytest.mark.parametrize("func, error", [
    (lambda lo, hi: (10, 110), ValueError),            # beyond the series
    (lambda lo: (10, 110), TypeError),                 # wrong signature
    (lambda lo, hi: (20, 10), ValueError),             # empty window
    (lambda lo, hi: 20, TypeError),                    # not a pair
    (lambda lo, hi: (10, 20, 30), TypeError),          # too many values
    (lambda lo, hi: ("abc", 10.0), TypeError)])        # not integers
def test_bad_trajectory_functions_are_rejected_at_construction(func, error):
    with pytest.raises(error):
        windowed_env(func)


def test_episode_length_is_the_window_length():
    calls = {"n": 0}

    def growing(lo, hi):
        calls["n"] += 1
        return 10, 11 + calls["n"]
    env = windowed_env(growing)            # one validating call here, like Microgrid._check_trajectory_func
    for expected in range(3, 7):
        env.reset()
        n_steps, done = 0, False
        while not done:
            _, _, done, _ = env.step(env.action_space.sample())
            n_steps += 1
        assert n_steps == expected


# ---- rule-based control (test_rbc.py) ----------------------------------------------------------------------------------
def test_rule_based_control_orders_by_cost_runs_and_resets():
    rbc = RuleBasedControl(fixture_microgrid())
    costs = [el.marginal_cost for el in rbc.priority_list]
    assert costs == sorted(costs)
    assert len(rbc.microgrid.log) == 0
    log = rbc.run(10)
    assert len(log) == 10 and log.equals(rbc.microgrid.log)
    rbc.reset()
    assert len(rbc.microgrid.log) == 0


CHECKS = [v for k, v in sorted(globals().items()) if k.startswith("test_")]
