"""GPU parity of the composed-microgrid path (mgc_* in libpymgrid_b200.so): the checks tests/test_compose_host.py runs
against the host build of the same C source, here through the CUDA build on cuda:0 -- the reference's recorded outputs
(tests/golden/compose.npz) bit for bit, batches against the Python oracle, one-launch rollouts against single steps.
(The file name sorts last on purpose: the fused path's GPU suites run first.)"""
import numpy as np
import pytest
import torch

from tests import compose_checks as K
from tests.compose_checks import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", ["container", "gym_sorted"])
@pytest.mark.parametrize("case", CASES, ids=[c.label for c in CASES])
def test_composed_microgrid_reproduces_reference_on_gpu(case, order):
    K.check_microgrid_reproduces_reference(case, order, None)


@pytest.mark.parametrize("label", ["several_of_each", "pairwise_sums", "load_pv_genset"])
def test_batch_matches_oracle_and_rollout_matches_steps_on_gpu(label):
    K.check_batch_matches_oracle_and_rollout_matches_steps(label, None, n_envs=389, T=7)    # four tiles, ragged last tile


def test_microgrid_constructor_routes_module_lists_outside_the_fused_scope():
    """pymgrid_b200.Microgrid(modules): the reference's own balance-test grid (a load and a renewable, nothing else;
    tests/microgrid/test_microgrid.py:188-319) runs on the composed path with the reference's known answers"""
    import pymgrid_b200
    from pymgrid_b200.compose import ComposedMicrogrid
    from pymgrid_b200.modules import LoadModule, RenewableModule
    rng = np.random.default_rng(0)
    load_ts = 10 * rng.random(100)
    pv_ts = load_ts + 5 * rng.random(100) * (rng.random(100) > 0.5) - 3 * rng.random(100)
    pv_ts = np.abs(pv_ts)
    mg = pymgrid_b200.Microgrid([LoadModule(time_series=load_ts, raise_errors=True),
                                 RenewableModule(time_series=pv_ts, raise_errors=True)])
    assert isinstance(mg, ComposedMicrogrid)
    for step in range(100):
        obs, reward, done, info = mg.run(mg.get_empty_action())
        loss_load = max(load_ts[step] - pv_ts[step], 0)
        assert -1 * reward == mg.modules.balancing[0].loss_load_cost * loss_load
    log = mg.log
    assert len(log) == 100
    assert np.array_equal(log[("load", 0, "load_met")].to_numpy(), load_ts)
    assert np.array_equal(log[("renewable", 0, "renewable_used")].to_numpy(), np.minimum(load_ts, pv_ts))
    assert np.array_equal(log[("renewable", 0, "curtailment")].to_numpy(), pv_ts - np.minimum(load_ts, pv_ts))


def test_large_batch_energy_balance_and_replica_consistency():
    """65 536 replicas of one composition: identical actions give identical rows, every env balances, and env 0 equals
    the oracle"""
    from oracle.compose import ComposedOracle
    from pymgrid_b200.compose import FLAG_BALANCE, ComposedBatch
    case = next(c for c in CASES if c.label == "several_of_each")
    n, T = 65536, 5
    batch = ComposedBatch([case.modules()], np.zeros(n, dtype=np.int64), obs_order="container", with_info=True,
                          microgrid_kwargs=case.microgrid_kwargs)
    comp = batch.comp
    rng = np.random.default_rng(11)
    a = rng.random((T, 1, comp.n_act))
    actions = torch.from_numpy(np.broadcast_to(a, (T, n, comp.n_act)).copy()).cuda()
    out = batch.rollout(actions, ring=1)
    torch.cuda.synchronize()
    reward = out["reward"].cpu().numpy()
    assert (reward == reward[:, :1]).all()
    obs = out["obs_ring"][0]
    assert bool((obs == obs[:1]).all())
    assert not bool((out["flags"] & FLAG_BALANCE).any())
    orc = ComposedOracle(case.modules(), **case.microgrid_kwargs)
    for k in range(T):
        control = {name: [a[k, 0, s.act_col:s.act_col + s.n_act] if s.n_act == 2 else a[k, 0, s.act_col] for s in slots]
                   for name, slots in comp.controllable()}
        o, r, d, info = orc.run(control, normalized=True)
        assert r == reward[k, 0]
    want = np.concatenate([np.asarray(o[m.name][m.index]).ravel() for m in orc.listing])
    assert np.array_equal(obs[0].cpu().numpy(), want)


# what the reference's TestMicrogridLoadPV family pins (tests/microgrid/test_microgrid.py:188-455), through the CUDA path
from tests.reference_suite_compose import checks  # noqa: E402

for _fn in checks(None):
    globals()[_fn.__name__ + "_on_gpu"] = _fn
del _fn


@pytest.mark.parametrize("case", K.DISCRETE_CASES, ids=[c.label for c in K.DISCRETE_CASES])
def test_discrete_env_and_rule_based_control_reproduce_reference_on_gpu(case):
    K.check_discrete_env_and_rbc(case, None)


def test_quickstart_notebook_replays_value_for_value_on_gpu():
    K.check_quickstart_notebook(None)


def test_batch_trajectory_windows_on_gpu():
    K.check_batch_trajectory_windows(None)


def test_set_forecaster_and_set_module_attr_on_gpu():
    K.check_set_forecaster(None)


def test_env_observation_keys_on_gpu():
    K.check_observation_keys(None)


def test_standalone_module_steps_on_gpu():
    K.check_standalone_module_steps(None)


def test_batch_log_recorder_on_gpu():
    K.check_batch_log_recorder(None)


def test_microgrid_helpers_on_gpu():
    K.check_microgrid_helpers(None)


def test_forecast_noise_on_gpu():
    K.check_forecast_noise(None)


def test_modules_step_batch_on_gpu():
    K.check_modules_step_batch(None)


def test_control_dict_conventions_on_gpu():
    K.check_control_dict_conventions(None)


def _wide_row_modules(rng, horizon):
    """378 observation elements at horizon 30: longer than the 256 the register-resident decode covers"""
    from pymgrid_b200 import modules as M
    T = 120
    ts = dict(forecaster="oracle", forecast_horizon=horizon)
    grid_ts = np.column_stack([rng.random(T) + 0.1, rng.random(T) * 0.5, rng.random(T), (rng.random(T) > 0.2).astype(float)])
    return ([M.LoadModule(10 + 5 * rng.random(T), **ts) for _ in range(4)] + [M.RenewableModule(8 * rng.random(T), **ts) for _ in range(4)]
            + [M.BatteryModule(min_capacity=2, max_capacity=20, max_charge=5, max_discharge=5, efficiency=0.9, init_soc=0.5),
               M.GensetModule(1, 12, 0.4, 2, 0.1), M.GridModule(30, 30, grid_ts, **ts)])


@pytest.mark.parametrize("label", ["several_of_each", "pairwise_sums", "load_pv_genset", "wide_row_h30", "narrow_row_h0", "odd_row_h2"])
def test_gather_emission_equals_per_element_decode(label, monkeypatch):
    """the staged-gather emitters (default) and the per-element decode (PYMGRID_B200_COMPOSE_GATHER=0) write the same rows,
    rewards and states -- per-env windows that start at different steps and run past the end of the series (fill rows,
    forecaster.py:120-149) included; rows of 378 elements (table re-read in blocks), of 18 (one pass)
    and of 39 (two passes, the second partial)"""
    from pymgrid_b200.compose import ComposedBatch
    from tests.compose_checks import _batch_case
    n, T = 391, 12
    outs = []
    for gather in ("1", "0"):
        monkeypatch.setenv("PYMGRID_B200_COMPOSE_GATHER", gather)
        if label.startswith(("wide_row", "narrow_row", "odd_row")):
            mods = _wide_row_modules(np.random.default_rng(4), int(label.rsplit("_h", 1)[1]))
            batch = ComposedBatch([mods[1:] if label.startswith("odd") else mods], np.zeros(n, dtype=np.int64))
        else:
            _, batch, _ = _batch_case(None, label, n, 5)
        comp = batch.comp
        final = int(comp.final_step)
        rng = np.random.default_rng(9)
        start = rng.integers(max(final - 30, 0), final - 2, n).astype(np.int32)      # the horizon crosses the end for most envs
        batch.step_counter.copy_(torch.from_numpy(start).to(batch.device))
        actions = torch.from_numpy(rng.random((T, n, comp.n_act))).to(batch.device)
        out = batch.rollout(actions, ring=3)
        torch.cuda.synchronize()
        outs.append((out["reward"].cpu().numpy(), out["obs_ring"].cpu().numpy(), out["done"].cpu().numpy(),
                     batch.fstate.cpu().numpy(), batch.istate.cpu().numpy(), batch.step_counter.cpu().numpy()))
    assert outs[0][1].shape[2] == {"wide_row_h30": 378, "narrow_row_h0": 18, "odd_row_h2": 39}.get(label, outs[0][1].shape[2])
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b, equal_nan=True)
    assert outs[0][2].any()      # some envs did reach the end of their window


@pytest.mark.parametrize("chunk", [4, 5, 64])
def test_host_rollout_equals_rollout(chunk):
    """ComposedBatch.host_rollout (pinned host buffers, chunked launches overlapped with the copies) returns what one
    device-side rollout returns: every reward, every done, the final state, the OR of the flags, the last rows"""
    from tests.compose_checks import _batch_case
    n, T = 777, 23
    _, ref, _ = _batch_case(None, "several_of_each", n, 5)
    _, batch, _ = _batch_case(None, "several_of_each", n, 5)
    comp = ref.comp
    rng = np.random.default_rng(21)
    h_act = torch.from_numpy(rng.random((T, n, comp.n_act)) * 1.3 - 0.1).pin_memory()      # some actions out of range: flags
    want = ref.rollout(h_act.to(ref.device), ring=1)
    h_rew = torch.full((T, n), np.nan, dtype=torch.float64).pin_memory()
    h_done = torch.full((T, n), 7, dtype=torch.uint8).pin_memory()
    for _ in range(2):                                       # the second call reuses streams and buffers
        for a in ("step_counter", "fstate", "istate"):
            getattr(batch, a).copy_(getattr(_batch_case(None, "several_of_each", n, 5)[1], a))
        got = batch.host_rollout(h_act, h_rew, h_done, chunk=chunk, ring=1)
        torch.cuda.synchronize()
        assert np.array_equal(h_rew.numpy(), want["reward"].cpu().numpy(), equal_nan=True)
        assert np.array_equal(h_done.numpy(), want["done"].cpu().numpy())
        assert torch.equal(got["obs_ring"], want["obs_ring"])
        assert torch.equal(got["flags"], want["flags"]) and bool(want["flags"].any())
        assert torch.equal(batch.fstate, ref.fstate) and torch.equal(batch.istate, ref.istate)
        assert torch.equal(batch.step_counter, ref.step_counter)
    with pytest.raises(ValueError):
        batch.host_rollout(h_act[:3], h_rew, h_done)
