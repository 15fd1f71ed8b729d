"""The drop-in host layer (pymgrid_b200.Microgrid, envs, algos: dict / DataFrame conversions, module views, naming,
trajectory windows, priority-list bookkeeping) on CPU.

The test FUNCTIONS are the GPU suite's own (tests/test_gpu_dropin.py), re-collected here with the engine replaced by the
oracle-backed stand-in of tests/oracle_engine.py: same golden vectors from the live reference, same assertions.  What
this file proves is the Python above the C-ABI; the kernels are proven by the `-m gpu` run of the very same functions."""
import pytest

import tests.test_gpu_dropin as G
from tests.oracle_engine import install


@pytest.fixture(autouse=True)
def _oracle_backed_engine(monkeypatch):
    install(monkeypatch)


test_microgrid_run_returns_reference_types_and_values = G.test_microgrid_run_returns_reference_types_and_values
test_get_log_matches_reference = G.test_get_log_matches_reference
test_legacy_seed_sample_action_known_answer = G.test_legacy_seed_sample_action_known_answer
test_running_past_the_end_raises_like_the_reference = G.test_running_past_the_end_raises_like_the_reference
test_discrete_env_single_and_batched = G.test_discrete_env_single_and_batched
test_continuous_env_config2_shape = G.test_continuous_env_config2_shape
test_microgrid_from_reference_style_modules = G.test_microgrid_from_reference_style_modules
test_randomised_grids_through_the_drop_in_classes = G.test_randomised_grids_through_the_drop_in_classes
test_battery_soc_before_the_first_update = G.test_battery_soc_before_the_first_update
test_default_module_names_and_trajectory_func = G.test_default_module_names_and_trajectory_func
test_env_from_modules_with_trajectory_func = G.test_env_from_modules_with_trajectory_func
test_rule_based_control_class = G.test_rule_based_control_class
test_rule_based_control_runs_until_done = G.test_rule_based_control_runs_until_done
test_reward_shaping_func_drop_in = G.test_reward_shaping_func_drop_in
test_module_views_on_the_engine = G.test_module_views_on_the_engine

import tests.test_zz_gpu_dropin_more as G2  # noqa: E402

test_set_forecaster_matches_reference = G2.test_set_forecaster_matches_reference
test_set_module_attr_like_the_reference_tests = G2.test_set_module_attr_like_the_reference_tests
test_env_observation_keys_match_reference = G2.test_env_observation_keys_match_reference
test_discrete_env_scenarios_like_the_reference_suite = G2.test_discrete_env_scenarios_like_the_reference_suite
test_env_log_matches_reference = G2.test_env_log_matches_reference
test_discrete_env_remove_action = G2.test_discrete_env_remove_action
