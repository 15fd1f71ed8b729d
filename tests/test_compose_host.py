"""The composed-microgrid path on the CPU: the C source of the CUDA path (pymgrid_b200/csrc/mg_compose.cu +
mg_compose_step.h) built for the host (tests/hostsim) under the package's own Python host layer
(pymgrid_b200/compose.py), against what the live reference returned (tests/golden/compose.npz) -- bit for bit, down to
the get_log() frame -- and against the Python oracle on batches.  The GPU twin is tests/test_zz_gpu_compose.py."""
import ctypes

import numpy as np
import pytest
import torch

from pymgrid_b200 import _cabi
from pymgrid_b200.compose import ComposedBatch, ComposedMicrogrid, Composition, in_fused_scope
from tests import hostsim
from tests import compose_checks as K
from tests.compose_checks import CASES



@pytest.fixture
def lib():
    """the host build of the kernel source, selected for the duration of one test (tests/hostsim.select)"""
    L = ctypes.CDLL(hostsim.build())
    hostsim.select(L)
    yield L
    hostsim.select(None)


@pytest.mark.parametrize("order", ["container", "gym_sorted"])
@pytest.mark.parametrize("case", CASES, ids=[c.label for c in CASES])
def test_composed_microgrid_reproduces_reference(case, order, lib):
    K.check_microgrid_reproduces_reference(case, order, lib)


def test_gym_sorted_order_is_a_permutation_by_name(lib):
    case = next(c for c in CASES if c.label == "custom_names")
    a = ComposedMicrogrid(case.modules(), obs_order="container", **case.microgrid_kwargs)
    b = ComposedMicrogrid(case.modules(), obs_order="gym_sorted", **case.microgrid_kwargs)
    # capital letters sort first: 'Abat', 'PV', then 'unbalanced_energy' (empty), 'wind', 'zload'
    assert [s.name for s in sorted(b.composition.slots, key=lambda s: s.obs_off) if s.obs_len] == ["Abat", "PV", "wind", "zload"]
    ra, rb = a._batch.observe()[0].numpy(), b._batch.observe()[0].numpy()
    for sa, sb in zip(a.composition.slots, b.composition.slots):
        assert np.array_equal(ra[sa.obs_off:sa.obs_off + sa.obs_len], rb[sb.obs_off:sb.obs_off + sb.obs_len])


@pytest.mark.parametrize("label", ["several_of_each", "pairwise_sums", "load_pv_genset"])
def test_batch_matches_oracle_and_rollout_matches_steps(label, lib):
    K.check_batch_matches_oracle_and_rollout_matches_steps(label, lib, n_envs=131, T=9)     # two tiles, ragged last tile


def test_layout_validation_and_scope(lib):
    case = next(c for c in CASES if c.label == "per_module_horizons")
    assert not in_fused_scope(case.modules())
    fused = next(c for c in CASES if c.label == "past_the_end")
    assert in_fused_scope(fused.modules(), add_unbalanced_module=False)
    with pytest.raises(ValueError):     # different compositions in one batch
        ComposedBatch([case.modules(), next(c for c in CASES if c.label == "load_only").modules()])
    comp = Composition(case.modules(), obs_order="container")
    assert comp.n_act == 4 and comp.obs_dim == 4 + 1 + 2 + 4 + 4 * 6
    assert [s.kind for s in comp.dispatch] == ["load", "genset", "battery", "grid", "renewable", "balancing"]
    b = ComposedBatch([comp])
    with pytest.raises(ValueError):
        b.step(np.zeros((1, 3)))
    # a module table the library must refuse: dispatch order broken
    from pymgrid_b200.compose import MgcLayout
    tab = comp.module_table()
    tab[0], tab[1] = tab[1], tab[0]
    L = MgcLayout()
    L.abi_version, L.n_modules, L.modules = 1, len(comp.slots), tab
    h = ctypes.c_void_p()
    assert b._L.mgc_create(ctypes.byref(L), ctypes.byref(h)) != 0
    assert b"mgc_create" in b._L.mg_last_error()


def test_product_path_needs_cuda():
    """no host library selected, no GPU in this container: constructing the product path must fail loudly, never fall back"""
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    hostsim.select(None)
    case = next(c for c in CASES if c.label == "load_pv")
    with pytest.raises(_cabi.EngineError):
        ComposedMicrogrid(case.modules())
    import pymgrid_b200
    with pytest.raises(_cabi.EngineError):
        pymgrid_b200.Microgrid(case.modules())        # routed to the composed path (two-module list), which needs CUDA


@pytest.mark.parametrize("case", K.DISCRETE_CASES, ids=[c.label for c in K.DISCRETE_CASES])
def test_discrete_env_and_rule_based_control_reproduce_reference(case, lib):
    K.check_discrete_env_and_rbc(case, lib)


def test_env_and_controller_constructors_route_to_the_composed_path(lib):
    """pymgrid_b200.envs.DiscreteMicrogridEnv / ContinuousMicrogridEnv / algos.RuleBasedControl / Microgrid take module lists
    outside the fused scope and hand them to the composed classes (same surface)"""
    import pymgrid_b200
    from pymgrid_b200.compose import ComposedContinuousEnv, ComposedDiscreteEnv, ComposedRuleBasedControl
    from pymgrid_b200.envs import ContinuousMicrogridEnv, DiscreteMicrogridEnv
    case = next(c for c in K.DISCRETE_CASES if c.label == "two_batteries_grid")
    env = DiscreteMicrogridEnv(case.modules())
    assert isinstance(env, ComposedDiscreteEnv) and env.action_space.n == 6
    obs = env.reset()
    assert obs.shape == env.observation_space.shape and ((0 <= obs) & (obs <= 1)).all()
    obs, reward, done, info = env.step(env.sample_action())
    assert isinstance(reward, float) and isinstance(done, bool) and set(info) == {"load", "renewable", "battery", "grid", "balancing"}
    cenv = ContinuousMicrogridEnv(case.modules(), batch=7)
    assert isinstance(cenv, ComposedContinuousEnv) and cenv.action_space.shape == (3,)
    assert cenv.action_layout == {("battery", 0): 0, ("battery", 1): 1, ("grid", 0): 2}
    obs, reward, done, _ = cenv.step(cenv.sample_action())
    assert obs.shape == (7, cenv.observation_space.shape[0]) and reward.shape == (7,)
    single = ContinuousMicrogridEnv(case.modules())
    mg = pymgrid_b200.Microgrid(case.modules())
    a = np.array([0.3, 0.8, 0.55])
    o1, r1, d1, _ = single.step(a)
    _, r2, d2, _ = mg.run({"battery": [0.3, 0.8], "grid": [0.55]})
    assert r1 == r2 and d1 == d2
    env3 = DiscreteMicrogridEnv.from_microgrid(mg)
    assert isinstance(env3, ComposedDiscreteEnv)
    assert env3.current_step == mg.current_step == 1          # the copy carries the live state
    env3.step(0)
    assert env3.current_step == 2 and mg.current_step == 1
    assert isinstance(pymgrid_b200.algos.RuleBasedControl(mg), ComposedRuleBasedControl)


def test_quickstart_notebook_replays_value_for_value(lib):
    K.check_quickstart_notebook(lib)


def test_batch_trajectory_windows(lib):
    K.check_batch_trajectory_windows(lib)


def test_set_forecaster_and_set_module_attr(lib):
    K.check_set_forecaster(lib)


def test_env_observation_keys(lib):
    K.check_observation_keys(lib)


def test_standalone_module_steps(lib):
    K.check_standalone_module_steps(lib)


def test_batch_log_recorder(lib):
    K.check_batch_log_recorder(lib)


def test_microgrid_helpers(lib):
    K.check_microgrid_helpers(lib)


def test_streaming_sum_equals_numpy_sum_at_every_length(lib):
    """MgcSum (mg_compose_step.h): the sum of a growing list in numpy's pairwise order, queried at every length up to 128"""
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(0, 129))
        v = (10 * rng.random(n) - 3) * 10.0 ** rng.integers(-3, 4, n)
        out = np.zeros(n + 1)
        lib.mgc_test_growing_sums(v.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
        assert all(out[k] == (float(np.sum(list(v[:k]))) if k else 0.0) for k in range(n + 1))


def test_forecast_noise(lib):
    K.check_forecast_noise(lib)


def test_modules_step_batch(lib):
    K.check_modules_step_batch(lib)


def test_control_dict_conventions(lib):
    K.check_control_dict_conventions(lib)
