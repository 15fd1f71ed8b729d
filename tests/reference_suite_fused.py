"""The reference's API / env / controller tests on its shared fixture grid (tests/helpers/modular_microgrid.py:14-55: genset
10-50 @ 0.5, battery 0-100 / 50 / 50, PV == 50, load == 60, grid import 100 / export 0 at price 1, raise_errors on the grid)
restated against this package: tests/microgrid/test_microgrid.py:12-186 (TestMicrogrid), tests/envs/test_trajectory.py,
tests/control/test_rbc.py -- same set-up, same assertions.  Collected twice: on the CPU with the engine replaced by the
oracle-backed stand-in (tests/test_reference_suite_fused_host.py) and on the GPU (tests/test_zz_gpu_dropin_more.py).

Not restated: the MicrogridSpace sampling tests (test_action_space*: gym space objects are not mirrored), test_init of TestRBC
and test_trajectory_serialization (object equality / YAML round trips of the reference classes).
"""
import unittest

import numpy as np

import pymgrid_b200
from pymgrid_b200.algos import RuleBasedControl
from pymgrid_b200.envs import DiscreteMicrogridEnv
from pymgrid_b200.modules import BatteryModule, GensetModule, GridModule, LoadModule, RenewableModule


def get_modular_microgrid(remove_modules=(), timeseries_length=100, modules_only=False, **microgrid_kw):
    modules = dict(
        genset=GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5),
        battery=BatteryModule(min_capacity=0, max_capacity=100, max_charge=50, max_discharge=50, efficiency=1.0, init_soc=0.5),
        renewable=RenewableModule(time_series=50 * np.ones(timeseries_length)),
        load=LoadModule(time_series=60 * np.ones(timeseries_length)),
        grid=GridModule(max_import=100, max_export=0, time_series=np.ones((timeseries_length, 3)), raise_errors=True))
    for module in remove_modules:
        modules.pop(module)
    modules = list(modules.values())
    if modules_only:
        return modules
    return pymgrid_b200.Microgrid(modules, **microgrid_kw)


class TestMicrogrid(unittest.TestCase):
    def test_from_scenario(self):
        for j in range(25):
            with self.subTest(microgrid_number=j):
                microgrid = pymgrid_b200.Microgrid.from_scenario(j)
                self.assertTrue(hasattr(microgrid, "load"))
                self.assertTrue(hasattr(microgrid, "pv"))
                self.assertTrue(hasattr(microgrid, "battery"))
                self.assertTrue(hasattr(microgrid, "grid") or hasattr(microgrid, "genset"))

    def test_empty_action_with_load(self):
        action = get_modular_microgrid().get_empty_action()
        self.assertIn('battery', action)
        self.assertIn('genset', action)
        self.assertIn('grid', action)
        self.assertNotIn('load', action)
        self.assertTrue(all(v == [None] for v in action.values()))

    def test_sample_action(self):
        microgrid = get_modular_microgrid()
        action = microgrid.sample_action()
        for module_name, action_list in action.items():
            for module_num, _act in enumerate(action_list):
                action_arr = np.atleast_1d(np.array(_act))
                self.assertEqual(action_arr.shape[0], microgrid.modules[module_name][module_num].action_space.shape[0])
                self.assertTrue(((0 <= action_arr) & (action_arr <= 1)).all())

    def test_sample_action_all_modules_populated(self):
        microgrid = get_modular_microgrid()
        action = microgrid.sample_action()
        for module_name, module_list in microgrid.fixed.iterdict():
            for module_num, module in enumerate(module_list):
                empty_action_space = module.action_space.shape == (0, )
                try:
                    _ = action[module_name][module_num]
                    has_corresponding_action = True
                except KeyError:
                    has_corresponding_action = False
                self.assertTrue(empty_action_space != has_corresponding_action)  # XOR

    def test_current_step(self):
        microgrid = get_modular_microgrid()
        self.assertEqual(microgrid.current_step, 0)
        for j in range(4):
            microgrid.run(microgrid.sample_action())
            self.assertEqual(microgrid.current_step, j + 1)

    def test_current_step_after_reset(self):
        microgrid = get_modular_microgrid()
        microgrid.run(microgrid.sample_action())
        self.assertEqual(microgrid.current_step, 1)
        microgrid.reset()
        self.assertEqual(microgrid.current_step, 0)

    def test_set_module_attr_forecast_horizon(self):
        microgrid = get_modular_microgrid()
        microgrid.set_module_attr('forecast_horizon', 50)
        fh = [module.forecast_horizon for module in microgrid.modules.iterlist() if hasattr(module, 'forecast_horizon')]
        self.assertEqual(min(fh), max(fh))
        self.assertEqual(min(fh), 50)

    def test_set_module_attr_bad_attr_name(self):
        with self.assertRaises(AttributeError):
            get_modular_microgrid().set_module_attr('blah', 'blah')

    def test_get_cost_info(self):
        cost_info = get_modular_microgrid().get_cost_info()
        for module in ('genset', 'battery', 'renewable', 'load', 'grid', 'balancing'):
            self.assertIn(module, cost_info.keys())
            self.assertEqual(len(cost_info[module]), 1)
            self.assertEqual(set(cost_info[module][0]), {'production_marginal_cost', 'absorption_marginal_cost'})

    def test_set_initial_step(self):
        microgrid = get_modular_microgrid()
        self.assertEqual(microgrid.initial_step, 0)
        microgrid.initial_step = 1
        self.assertEqual(microgrid.initial_step, 1)
        microgrid.reset()
        self.assertEqual(microgrid.current_step, 1)


class TestTrajectory(unittest.TestCase):
    def check_initial_final_steps(self, env, expected_env_initial, expected_env_final, expected_module_initial, expected_module_final):
        self.assertEqual(env.initial_step, expected_env_initial)
        self.assertEqual(env.final_step, expected_env_final)
        env.reset()
        self.assertEqual(env.initial_step, expected_env_initial)
        self.assertEqual(env.final_step, expected_env_final)
        self.assertEqual(env.modules.get_attrs('initial_step', unique=True).item(), expected_module_initial)
        self.assertEqual(env.modules.get_attrs('final_step', unique=True).item(), expected_module_final)

    def env(self, trajectory_func, timeseries_length=100):
        return DiscreteMicrogridEnv(get_modular_microgrid(timeseries_length=timeseries_length, modules_only=True),
                                    trajectory_func=trajectory_func)

    def test_none_trajectory(self):
        self.check_initial_final_steps(self.env(None), 0, 100, 0, 100)

    def test_deterministic_trajectory(self):
        self.check_initial_final_steps(self.env(lambda initial_step, final_step: (10, 20)), 0, 100, 10, 20)

    def test_stochastic_trajectory(self):
        def trajectory_func(initial_step, final_step):
            initial = np.random.randint(low=initial_step + 1, high=final_step - 2)
            final = np.random.randint(low=initial, high=final_step)
            return int(initial), int(max(final, initial + 1))
        env = self.env(trajectory_func)
        env.reset()
        self.assertEqual((env.initial_step, env.final_step), (0, 100))
        ini = env.modules.get_attrs('initial_step', unique=True).item()
        fin = env.modules.get_attrs('final_step', unique=True).item()
        self.assertTrue(0 < ini < fin <= 100)

    def test_bad_trajectory_out_of_range(self):
        with self.assertRaises(ValueError):
            self.env(lambda initial_step, final_step: (10, 110))

    def test_bad_trajectory_bad_signature(self):
        with self.assertRaises(TypeError):
            self.env(lambda initial_step: (10, 110))

    def test_bad_trajectory_initial_gt_final(self):
        with self.assertRaises(ValueError):
            self.env(lambda initial_step, final_step: (20, 10))

    def test_bad_trajectory_scalar_output(self):
        with self.assertRaises(TypeError):
            self.env(lambda initial_step, final_step: 20)

    def test_bad_trajectory_too_many_outputs(self):
        with self.assertRaises(TypeError):
            self.env(lambda initial_step, final_step: (10, 20, 30))

    def test_bad_trajectory_wrong_output_types(self):
        with self.assertRaises(TypeError):
            self.env(lambda initial_step, final_step: ('abc', 10.0))

    def test_correct_trajectory_length(self):
        def trajectory_func(initial_step, final_step):
            trajectory_func.n_resets += 1
            return 10, 11 + trajectory_func.n_resets
        trajectory_func.n_resets = 0
        env = self.env(trajectory_func)
        for correct_trajectory_length in range(3, 7):
            with self.subTest(correct_trajectory_length=correct_trajectory_length):
                env.reset()
                n_steps, done = 0, False
                while not done:
                    _, _, done, _ = env.step(env.action_space.sample())
                    n_steps += 1
                self.assertEqual(n_steps, correct_trajectory_length)


class TestRBC(unittest.TestCase):
    def setUp(self):
        self.rbc = RuleBasedControl(get_modular_microgrid())

    def test_priority_list(self):
        for element_1, element_2 in zip(self.rbc.priority_list[:-1], self.rbc.priority_list[1:]):
            self.assertLessEqual(element_1.marginal_cost, element_2.marginal_cost)

    def run_once(self):
        rbc = self.rbc
        self.assertEqual(len(rbc.microgrid.log), 0)
        n_steps = 10
        log = rbc.run(n_steps)
        self.assertEqual(len(log), n_steps)
        self.assertTrue(log.equals(rbc.microgrid.log))
        return rbc

    def test_run_once(self):
        self.run_once()

    def test_reset_after_run(self):
        rbc = self.run_once()
        rbc.reset()
        self.assertEqual(len(rbc.microgrid.log), 0)


SUITES = (TestMicrogrid, TestTrajectory, TestRBC)
