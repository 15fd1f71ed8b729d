"""Shared bodies of the composed-microgrid parity tests: run once on the CPU against the host build of the C source
(tests/test_compose_host.py) and once on the GPU against the CUDA build (tests/test_zz_gpu_compose.py).  `lib` is the
host-build library handle, or None for the product path (CUDA)."""
import numpy as np
import pytest
import torch

from oracle.compose import ComposedOracle
from pymgrid_b200.compose import ComposedBatch, ComposedMicrogrid
from tests.compose_cases import BALANCE_COLS, load_cases
from tests.hostsim import select as hostsim_select

CASES = load_cases()


def widths(mg):
    return [(name, [s.n_act for s in slots]) for name, slots in mg.composition.controllable()]


def listing_flat(obs, mg):
    return np.concatenate([np.asarray(obs[s.name][s.index]).ravel() for s in mg.composition.slots] + [np.zeros(0)])


def host(t):
    return t.detach().cpu().numpy()


def check_microgrid_reproduces_reference(case, order, lib):
    hostsim_select(lib)
    mg = ComposedMicrogrid(case.modules(), obs_order=order, **case.microgrid_kwargs, **case.callable_kwargs)
    assert [(s.name, s.index) for s in mg.composition.slots] == [(n, j) for n, j, _ in case.names]
    assert {k: len(v) for k, v in mg.get_empty_action().items()} == case.json("empty_action")
    reset = mg.reset()
    assert list(reset.keys()) == case.json("reset_keys")
    assert np.array_equal(listing_flat(reset, mg), case["obs_reset"])
    n = len(case["rewards"])
    reset_at, n_resets = list(case.spec.get("reset_at", [])), 0
    for k in range(n):
        if k in reset_at:
            assert np.array_equal(listing_flat(mg.reset(), mg), case["reset_obs_rows"][n_resets]), k
            n_resets += 1
        obs, reward, done, info = mg.run(case.control(k, widths(mg)), normalized=bool(case["normalized"][k]))
        assert mg.current_step == int(case["steps_after"][k]), k
        assert list(obs.keys()) == case.json("run_keys")
        assert reward == case["rewards"][k] and done == bool(case["dones"][k]), k
        assert np.array_equal(listing_flat(obs, mg), case["obs"][k]), k
        for i, s in enumerate(mg.composition.slots):
            inf, want = info[s.name][s.index], case["info"][k, i]
            assert inf.get("provided_energy", 0.0) == want[0] and inf.get("absorbed_energy", 0.0) == want[1], (k, s.name)
            assert inf.get("co2_production", inf.get("curtailment", 0.0)) == want[2]
            assert ("absorbed_energy" in inf) == bool(want[4])
        t, f, i_ = mg._state()
        state = []
        for s in mg.composition.slots:
            if s.kind == "battery":
                state += list(f[s.fstate_off:s.fstate_off + 2])
            elif s.kind == "genset":
                state += list(i_[s.istate_off:s.istate_off + 4])
        assert np.array_equal(np.array(state, dtype=np.float64), case["states"][k]), k
    if int(case["raised_at"]) >= 0:
        ctrl = {name: [np.array([0.5, 0.5]) if w == 2 else 0.5 for w in ws] for name, ws in widths(mg)}
        exc = {"RuntimeError": RuntimeError, "IndexError": IndexError}[str(case["raised_type"])]
        with pytest.raises(exc):
            mg.run(ctrl)
    assert mg.current_step == int(case["current_step"])
    if str(case["log_raises"]):
        # the reference's own get_log() fails under a trajectory_func (tests/golden/make_compose.py); ours indexes the rows
        # since the last reset
        assert len(mg.get_log()) == n - max(reset_at)
        return
    log = mg.get_log()
    assert [list(c) for c in log.columns] == case.json("log_columns")
    assert np.array_equal(log.to_numpy(dtype=np.float64), case["log_values"], equal_nan=True)
    assert np.array_equal(np.array(log.index), case["log_index"])
    assert mg.current_step == int(case["current_step"])
    ss = mg.state_series()          # recorded at the end of the run
    assert [[str(x) for x in i] for i in ss.index] == case.json("state_series_index")
    assert np.array_equal(ss.to_numpy(), case["state_series_values"])
    if n:
        assert np.array_equal(log["balance"][0][list(BALANCE_COLS)].to_numpy()[:n], case["balance"][:n])


def _batch_case(lib, label, n_envs, seed):
    """a batch whose envs are re-parameterised copies of one golden composition, and the oracle twin of every env"""
    hostsim_select(lib)
    case = next(c for c in CASES if c.label == label)
    rng = np.random.default_rng(seed)
    configs = []
    for _ in range(3):
        mods = case.modules()
        for m in mods:
            m = m[1] if isinstance(m, tuple) else m
            if hasattr(m, "time_series") and m.module_type[0] != "grid":
                m.time_series = m.time_series * rng.uniform(0.5, 1.5)
            if m.module_type[0] == "battery":
                m.init_charge = m.min_capacity + rng.random() * (m.max_capacity - m.min_capacity)
                m.init_soc = m.init_charge / m.max_capacity
        configs.append(mods)
    env_config = rng.integers(0, 3, n_envs)
    batch = ComposedBatch(configs, env_config, obs_order="container", with_info=True, microgrid_kwargs=case.microgrid_kwargs)
    oracles = [ComposedOracle(configs[c], **case.microgrid_kwargs) for c in env_config]
    return case, batch, oracles


def check_batch_matches_oracle_and_rollout_matches_steps(label, lib, n_envs=131, T=9):
    hostsim_select(lib)
    case, batch, oracles = _batch_case(lib, label, n_envs, 7)
    comp = batch.comp
    rng = np.random.default_rng(3)
    actions = rng.random((T, n_envs, comp.n_act))
    ctl = comp.controllable()
    want_r, want_obs = np.zeros((T, n_envs)), None
    for e, orc in enumerate(oracles):
        for k in range(T):
            control = {name: [actions[k, e, s.act_col:s.act_col + s.n_act] if s.n_act == 2 else actions[k, e, s.act_col]
                              for s in slots] for name, slots in ctl}
            obs, r, d, info = orc.run(control, normalized=True)
            want_r[k, e] = r
        if want_obs is None:
            want_obs = np.zeros((n_envs, comp.obs_dim))
        want_obs[e] = np.concatenate([np.asarray(obs[m.name][m.index]).ravel() for m in orc.listing])
    # one launch for the whole rollout ...
    out = batch.rollout(torch.from_numpy(actions).to(batch.device), ring=2)
    assert np.array_equal(out["reward"].cpu().numpy(), want_r)
    assert np.array_equal(out["obs_ring"][(T - 1) % 2].cpu().numpy(), want_obs)
    assert batch.launch_count == 1
    # ... equals T single steps on a fresh batch
    _, batch2, _ = _batch_case(lib, label, n_envs, 7)
    for k in range(T):
        obs, r, d, info = batch2.step(torch.from_numpy(actions[k]).to(batch2.device))
        assert np.array_equal(r.cpu().numpy(), want_r[k])
    assert np.array_equal(obs.cpu().numpy(), want_obs)
    assert np.array_equal(batch2.fstate.cpu().numpy(), batch.fstate.cpu().numpy()) and np.array_equal(batch2.istate.cpu().numpy(), batch.istate.cpu().numpy())
    # the same rows when the kernel normalises the series itself instead of gathering from the pre-normalised pool
    batch3 = ComposedBatch([c for c in batch.compositions], batch.env_config, obs_order="container", prenormalised=False)
    for a in ("fstate", "istate"):
        getattr(batch3, a).copy_(getattr(_batch_case(lib, label, n_envs, 7)[1], a))
    out3 = batch3.rollout(torch.from_numpy(actions).to(batch3.device), ring=2)
    assert np.array_equal(out3["reward"].cpu().numpy(), want_r) and np.array_equal(out3["obs_ring"][(T - 1) % 2].cpu().numpy(), want_obs)
    # checkpoint / resume: restoring the state tensors and replaying the tail gives the same tail
    _, batch4, _ = _batch_case(lib, label, n_envs, 7)
    head = batch4.rollout(torch.from_numpy(actions[:4]).to(batch4.device), obs=False)
    saved = batch4.state_dict()
    tail1 = batch4.rollout(torch.from_numpy(actions[4:]).to(batch4.device), ring=1)
    batch4.load_state_dict(saved)
    tail2 = batch4.rollout(torch.from_numpy(actions[4:]).to(batch4.device), ring=1)
    assert np.array_equal(head["reward"].cpu().numpy(), want_r[:4]) and np.array_equal(tail1["reward"].cpu().numpy(), want_r[4:])
    assert torch.equal(tail1["reward"], tail2["reward"]) and torch.equal(tail1["obs_ring"], tail2["obs_ring"])
    # masked reset: only the step counter of the masked envs moves (microgrid.py:205-225)
    mask = (np.arange(n_envs) % 3 == 0).astype(np.uint8)
    before = batch.fstate.clone()
    batch.reset(mask)
    assert np.array_equal(batch.step_counter.cpu().numpy(), np.where(mask, comp.initial_step, T + comp.initial_step))
    assert torch.equal(before, batch.fstate)


# ---- priority lists: DiscreteMicrogridEnv / RuleBasedControl on composed microgrids ------------------------------------
DISCRETE_CASES = load_cases("compose_discrete.npz")


def check_discrete_env_and_rbc(case, lib):
    hostsim_select(lib)
    from pymgrid_b200.compose import ComposedDiscreteEnv, ComposedRuleBasedControl
    kw = {}
    rows = lambda pls: [[[el.module[0], el.module[1], el.module_actions, el.action] for el in pl] for pl in pls]     # noqa: E731
    # the action tables, with and without the redundant genset lists
    for flag in (0, 1):
        env = ComposedDiscreteEnv(case.modules(), obs_order="container", remove_redundant_gensets=bool(flag), **kw, **case.microgrid_kwargs)
        assert rows(env.actions_list) == case.json(f"table_{flag}")
        assert env.action_space.n == len(case.json(f"table_{flag}"))
    env = ComposedDiscreteEnv(case.modules(), obs_order="container", remove_redundant_gensets=False, **kw, **case.microgrid_kwargs)
    assert env.observation_space.shape == (case["obs"].shape[1],)
    env.reset()
    mg = env._mg
    for k, a in enumerate(case["actions"]):
        obs, reward, done, info = env.step(int(a))
        assert reward == case["rewards"][k] and done == bool(case["dones"][k]), k
        assert np.array_equal(obs, case["obs"][k]), k          # container order = the recorded listing order
        t, f, i_ = mg._state()
        state = []
        for s in mg.composition.slots:
            if s.kind == "battery":
                state += list(f[s.fstate_off:s.fstate_off + 2])
            elif s.kind == "genset":
                state += list(i_[s.istate_off:s.istate_off + 4])
        assert np.array_equal(np.array(state, dtype=np.float64), case["states"][k]), k
    with pytest.raises(ValueError):
        env.step(env.action_space.n)
    # remove_action (envs/discrete/discrete.py:90-105): the remaining actions are renumbered and keep their own lists
    if env.action_space.n > 2:
        e1 = ComposedDiscreteEnv(case.modules(), obs_order="container", remove_redundant_gensets=False, **kw, **case.microgrid_kwargs)
        e2 = ComposedDiscreteEnv(case.modules(), obs_order="container", remove_redundant_gensets=False, **kw, **case.microgrid_kwargs)
        e2.remove_action(1)
        assert e2.action_space.n == e1.action_space.n - 1 and e2.actions_list[1] == e1.actions_list[2]
        for old, new in ((0, 0), (2, 1), (e1.action_space.n - 1, e2.action_space.n - 1)):
            (o1, r1, d1, _), (o2, r2, d2, _) = e1.step(old), e2.step(new)
            assert r1 == r2 and np.array_equal(o1, o2)
    # the same action sequence for a batch of replicas, one launch
    benv = ComposedDiscreteEnv(case.modules(), obs_order="container", remove_redundant_gensets=False, batch=130, **kw,
                               **case.microgrid_kwargs)
    benv.reset()
    acts = torch.from_numpy(np.repeat(case["actions"][:, None], 130, axis=1).astype(np.int32)).to(benv.batch.device)
    out = benv.batch.rollout_discrete(acts, ring=1)
    assert np.array_equal(host(out["reward"]), np.repeat(case["rewards"][:, None], 130, axis=1))
    assert np.array_equal(host(out["obs_ring"][0]), np.repeat(case["obs"][-1][None], 130, axis=0))
    obs, reward, done, _ = benv.step(torch.full((130,), -1, dtype=torch.int32))       # not in the action space
    assert bool(torch.isnan(reward).all()) and bool((benv.batch.flags & 64).all())
    # rule-based control
    import pymgrid_b200
    mg = ComposedMicrogrid(case.modules(), obs_order="container", **kw, **case.microgrid_kwargs)
    rbc = pymgrid_b200.algos.RuleBasedControl(mg)
    assert isinstance(rbc, ComposedRuleBasedControl)
    assert rows([rbc.priority_list])[0] == case.json("rbc_list")
    log = rbc.run(max_steps=len(case["rbc_rewards"]))
    assert [list(c) for c in log.columns] == case.json("rbc_log_columns")
    assert np.array_equal(log.to_numpy(dtype=np.float64), case["rbc_log_values"], equal_nan=True)
    assert np.array_equal(log[("balance", 0, "reward")].to_numpy(), case["rbc_rewards"])
    assert mg.current_step == mg.initial_step           # the controller ran on a copy (rbc.py:28-30)
    # one launch for the whole run, B replicas
    batch = ComposedBatch([case.modules()], np.zeros(5, dtype=np.int64), microgrid_kwargs=case.microgrid_kwargs, **kw)
    T = len(case["rbc_rewards"])
    out = batch.rollout_discrete(np.full(5, rbc._index, dtype=np.int32), n_steps=T, obs=False)
    assert np.array_equal(host(out["reward"]), np.repeat(case["rbc_rewards"][:, None], 5, axis=1))


# ---- the reference's quick-start notebook, cell by cell ----------------------------------------------------------------
def check_quickstart_notebook(lib):
    """notebooks/quick-start.ipynb against pymgrid_b200: same cells (tests/golden/make_quickstart.py: notebook()), same
    displays, same log -- including the ten `sample_action(strict_bound=True)` steps under np.random.seed(0)"""
    hostsim_select(lib)
    import json
    import os
    import pymgrid_b200
    from pymgrid_b200.modules import BatteryModule, GridModule, LoadModule, RenewableModule
    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "golden", "make_quickstart.py")).read()
    start, end = src.index("def notebook("), src.index("def main():")
    ns = {"np": np}
    exec(compile(src[start:end], "quickstart_notebook", "exec"), ns)      # the notebook cells only, not the reference loader
    kw = {}
    got = ns["notebook"](pymgrid_b200.Microgrid, BatteryModule, LoadModule, RenewableModule, GridModule, obs_order="container", **kw)
    want = np.load(os.path.join(here, "golden", "quickstart.npz"))
    for key, value in got.items():
        ref = want[key]
        if isinstance(value, np.ndarray):
            assert np.array_equal(value, ref, equal_nan=True), key
        elif isinstance(value, float):
            assert value == float(json.loads(str(ref))), key
        else:
            assert json.loads(json.dumps(value)) == json.loads(str(ref)), key


def check_batch_trajectory_windows(lib):
    """per-env episode windows on a composed batch: reset puts every env at its own initial step, done fires at its own
    final_step - 1 (base_timeseries_module.py:124-125), and every env equals the oracle run with that window"""
    hostsim_select(lib)
    from pymgrid_b200.compose import ComposedContinuousEnv
    case = next(c for c in CASES if c.label == "trajectory_window")
    kw = {}
    n, T = 9, 12
    rng = np.random.default_rng(21)
    windows = [(int(a), int(a + b)) for a, b in zip(rng.integers(0, 20, n), rng.integers(3, 30, n))]
    it = iter([windows[0]] + windows)          # the constructor validates the function with one call (microgrid.py:167-199)
    env = ComposedContinuousEnv(case.modules(), obs_order="container", batch=n, trajectory_func=lambda lo, hi: next(it), **kw,
                                **case.microgrid_kwargs)
    obs0 = host(env.reset()).copy()
    assert np.array_equal(host(env.current_step), np.array([w[0] for w in windows]))
    actions = rng.random((T, n, env.composition.n_act))
    out = env.batch.rollout(torch.from_numpy(actions).to(env.batch.device), ring=1)
    done = host(out["done"])
    for e, (lo, hi) in enumerate(windows):
        orc = ComposedOracle(case.modules(), trajectory_func=lambda a, b, w=(lo, hi): w, **case.microgrid_kwargs)
        flat = lambda obs: np.concatenate([np.asarray(obs[m.name][m.index]).ravel() for m in orc.listing])      # noqa: E731
        assert np.array_equal(flat(orc.reset()), obs0[e])
        for k in range(T):
            control = {name: [actions[k, e, s.act_col] for s in slots] for name, slots in env.composition.controllable()}
            o, r, d, _ = orc.run(control)
            assert r == host(out["reward"])[k, e] and d == bool(done[k, e]), (e, k)
            assert d == (lo + k >= hi - 1)
        assert np.array_equal(flat(o), host(out["obs_ring"][0])[e])
    # masked reset: the envs that finished get a new window, the others keep theirs and their step
    finished = done[-1].astype(bool)
    new = [(1, 9)] * int(finished.sum())
    it2 = iter(new + [(0, 5)] * n)
    env.trajectory_func = lambda lo, hi: next(it2)
    before = host(env.current_step).copy()
    env.reset(mask=torch.from_numpy(finished.astype(np.uint8)).to(env.batch.device))
    after = host(env.current_step)
    assert np.array_equal(after[~finished], before[~finished])
    assert np.array_equal(host(env.batch.env_final_step)[~finished], np.array([w[1] for w in windows])[~finished])
    with pytest.raises(ValueError):
        env.batch.set_trajectories(np.zeros(n), np.full(n, 10 ** 6))


def check_set_forecaster(lib, golden_dir=None):
    """Microgrid.set_forecaster on the quick-start notebook's two-battery grid (tests/golden/make_set_forecaster.py)"""
    hostsim_select(lib)
    import os
    import pymgrid_b200
    from pymgrid_b200 import modules as M
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "set_forecaster.npz"))
    rng = np.random.default_rng(0)
    load, pv = 100 + 100 * rng.random(200), 200 * rng.random(200)
    mods = [M.BatteryModule(10, 100, 50, 50, 0.9, init_soc=0.2), M.BatteryModule(10, 1000, 10, 10, 0.7, init_soc=0.2),
            ("pv", M.RenewableModule(time_series=pv)), M.LoadModule(time_series=load),
            M.GridModule(100, 100, [0.2, 0.1, 0.5] * np.ones((200, 3)))]
    kw = {}
    mg = pymgrid_b200.Microgrid(mods, **kw)
    _set_forecaster_flow(mg, z, "quick", ["load"])
    with pytest.raises(AttributeError):
        mg.set_module_attr("blah", "blah")                  # tests/microgrid/test_microgrid.py:144-147
    mg.set_module_attr("forecast_horizon", 50)              # :135-142
    assert {m.forecast_horizon for m in mg.modules.iterlist() if hasattr(m, "forecast_horizon")} == {50}

PHASES = (("keep", 4), ("oracle24", 4), ("dict", 2), ("none", 3))


def _set_forecaster_flow(m, z, prefix, names):
    """the recorded flow of tests/golden/make_set_forecaster.py against microgrid `m`"""
    import json
    for phase, n in PHASES:
        if phase == "oracle24":
            m.set_forecaster("oracle", forecast_horizon=24)             # BASELINE config 4's call
        elif phase == "dict":
            m.set_forecaster({names[0]: None}, forecast_horizon=7)      # a silent no-op in the reference, mirrored
            with pytest.raises(NameError):
                m.set_forecaster({"no_such_module": None})
        elif phase == "none":
            m.set_forecaster(None)
        for k in range(n):
            flat, col, control = z[f"{prefix}_{phase}_actions"][k], 0, {}
            for name, mods in m.controllable.items():
                vals = []
                for _ in mods:
                    w = 2 if name == "genset" else 1
                    vals.append(np.array(flat[col:col + w]) if w == 2 else float(flat[col]))
                    col += w
                control[name] = vals
            obs, reward, done, info = m.run(control)
            assert reward == z[f"{prefix}_{phase}_rewards"][k], (phase, k)
            row = np.concatenate([np.asarray(x).ravel() for name in sorted(obs) for x in obs[name]])
            np.testing.assert_array_equal(row, z[f"{prefix}_{phase}_obs"][k])
        if phase == "dict":
            log = m.get_log()
            assert [list(c) for c in log.columns] == json.loads(str(z[f"{prefix}_log_columns"]))
            assert np.array_equal(log.to_numpy(dtype=np.float64), z[f"{prefix}_log_values"], equal_nan=True)
    # a shorter horizon: the reference's own get_log() raises a length mismatch here; ours keeps every column, NaN-padded
    assert str(z[f"{prefix}_final_log_raises"]) == "ValueError"
    assert len(m.get_log()) == sum(n for _, n in PHASES)


def check_observation_keys(lib):
    """observation_keys on a composed env: a single microgrid picks the selected elements out of its row, a batch has the
    kernel write only those (MgcLayout.obs_select); both equal the live reference's recorded observations"""
    hostsim_select(lib)
    import importlib.util
    import os
    from pymgrid_b200 import modules as M
    from pymgrid_b200.envs import DiscreteMicrogridEnv
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_observation_keys", os.path.join(here, "golden", "make_observation_keys.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    z = np.load(os.path.join(here, "golden", "observation_keys.npz"))
    kw = {}
    env = DiscreteMicrogridEnv(mk.composed_modules(M), observation_keys=mk.COMPOSED_KEYS, **kw)
    assert env.observation_space.shape == (z["composed_obs"].shape[1],)
    rows, rewards = mk.flow(env, mk.COMPOSED_ACTIONS)
    assert np.array_equal(rows, z["composed_obs"]) and np.array_equal(rewards, z["composed_rewards"])
    benv = DiscreteMicrogridEnv(mk.composed_modules(M), observation_keys=mk.COMPOSED_KEYS, batch=130, **kw)
    assert benv.batch.obs.shape == (130, z["composed_obs"].shape[1])
    assert np.array_equal(host(benv.reset()), np.repeat(z["composed_obs"][:1], 130, axis=0))
    for k, a in enumerate(mk.COMPOSED_ACTIONS):
        obs, reward, _, _ = benv.step(torch.full((130,), a, dtype=torch.int32))
        assert np.array_equal(host(obs), np.repeat(z["composed_obs"][k + 1][None], 130, axis=0)), k
        assert np.array_equal(host(reward), np.full(130, z["composed_rewards"][k]))
    with pytest.raises(NameError):
        DiscreteMicrogridEnv(mk.composed_modules(M), observation_keys=["no_such_field"], **kw)
    # the log a single env keeps: the microgrid's columns + the action it was given (envs/discrete/discrete.py:141)
    import json
    env = DiscreteMicrogridEnv(mk.composed_modules(M), **kw)
    for a in z["envlog_composed_actions"]:
        env.step(int(a))
    log = env.log
    assert [list(c) for c in log.columns] == json.loads(str(z["envlog_composed_columns"]))
    assert np.array_equal(log.to_numpy(dtype=np.float64), z["envlog_composed_values"], equal_nan=True)


# ---- modules on their own: BaseMicrogridModule.step / reset / state (the reference's operator API) ----------------------
def check_standalone_module_steps(lib):
    """module.step(action, normalized) on pymgrid_b200.modules' objects without a Microgrid, against the live reference's
    recorded outputs (tests/golden/make_module_steps.py): observation, reward, done, info, state, exception type"""
    hostsim_select(lib)
    import importlib.util
    import os
    import pymgrid_b200.compose as cp
    from pymgrid_b200 import modules as M
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_module_steps", os.path.join(here, "golden", "make_module_steps.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    z = np.load(os.path.join(here, "golden", "module_steps.npz"))
    try:
        rng = np.random.default_rng(5)
        for label, cls, kwargs in mk.specs():
            acts = mk.actions_for(label, cls, kwargs, rng)
            got = mk.run(getattr(M, cls)(**kwargs), cls, acts)
            for key, value in got.items():
                want = z[f"{label}_{key}"]
                if value.dtype.kind in "US":
                    assert str(value) == str(want), (label, key)
                else:
                    assert np.array_equal(value, want, equal_nan=True), (label, key)
        # the reference's genset known answers (tests/microgrid/modules/module_tests/test_genset_module.py:64-163)
        genset = M.GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5)
        obs, reward, done, info = genset.step(np.array([1.0, 20.0]), normalized=False)
        assert reward == -10.0 and not done and info["provided_energy"] == 20.0 and list(obs) == [1, 1, 0, 0]
        obs, reward, done, info = genset.step(np.array([1.0, 5.0]), normalized=False)         # below running_min: clamped up
        assert info["provided_energy"] == 10.0 and reward == -5.0
        obs, reward, done, info = genset.step(np.array([0.0, 30.0]), normalized=False)        # switched off: nothing produced
        assert info["provided_energy"] == 0.0 and reward == 0.0 and genset.current_status == 0
        assert genset.log_dict()["genset_production"] == [20.0, 10.0, 0.0]
        # what the reference's time-series module tests pin (module_tests/timeseries_modules.py:10-100, test_load_module.py,
        # test_renewable_module.py): on the series 2 - cos(pi t / 2), with and without a 24-step forecast, the state is the
        # window of the series, observations stay in [0, 1] until done, bounds are (min(0, ts), max(0, ts))
        wave = 2 - np.cos(np.pi * np.arange(100) / 2)
        for horizon in (0, 24):
            kw = dict(forecaster="oracle", forecast_horizon=horizon) if horizon else {}
            for cls, sign, given in ((M.LoadModule, -1, wave), (M.LoadModule, -1, -wave), (M.RenewableModule, 1, wave), (M.RenewableModule, 1, -wave)):
                module = cls(time_series=given, **kw)
                assert np.array_equal(module.state, sign * wave[:1 + horizon]) and len(module.state_dict()) == 1 + horizon
                assert module.min_obs.tolist() == [min(0, (sign * wave).min())] * (1 + horizon)
                assert module.max_obs.tolist() == [max(0, (sign * wave).max())] * (1 + horizon)
                done, steps = False, 0
                while not done:
                    act = np.array([]) if cls is M.LoadModule else float(np.random.default_rng(steps).uniform(0, 3))
                    obs, reward, done, info = module.step(act, normalized=False)
                    assert ((0 <= obs) & (obs <= 1)).all() and reward == 0.0 and obs.shape == (1 + horizon,)
                    if cls is M.LoadModule:
                        assert info == {"absorbed_energy": wave[steps]}
                    else:
                        assert info["provided_energy"] == min(act, wave[steps]) and info["curtailment"] == wave[steps] - info["provided_energy"]
                    steps += 1
                assert steps == 100 == module.current_step      # done is reported by the step taken AT final_step - 1 (base_timeseries_module.py:124-125)
        # the reference's exhaustive look-ahead check (test_genset_long_status_changes.py:217-262): the status next_status()
        # predicts for a goal is the status after stepping with that goal, for every start-up / wind-down time, both initial
        # states and every goal sequence
        import itertools
        for U, D, init in itertools.product(range(3), range(3), (True, False)):
            for goals in itertools.product((0, 1), repeat=4):
                g = M.GensetModule(10, 50, 0.5, start_up_time=U, wind_down_time=D, init_start_up=init)
                for goal in goals:
                    predicted = g.next_status(goal)
                    assert g.next_max_production(goal) == predicted * 50 and g.next_min_production(goal) == predicted * 10
                    g.step(np.array([goal, 0.0]), normalized=False)
                    assert g.current_status == predicted, (U, D, init, goals)
    finally:
        pass


def check_microgrid_helpers(lib):
    """get_forecast_horizon, to_normalized / from_normalized (round trip, module spaces), the grid's price / status columns"""
    hostsim_select(lib)
    from pymgrid_b200 import modules as M
    kw = {}
    rng = np.random.default_rng(3)
    load, pv = 100 + 100 * rng.random(60), 200 * rng.random(60)
    g = np.stack([rng.uniform(0.05, 0.9, 60), rng.uniform(0, 0.4, 60), rng.uniform(0, 0.6, 60)], axis=1)
    ts = dict(forecaster="oracle", forecast_horizon=4)
    mg = ComposedMicrogrid([M.BatteryModule(10, 100, 50, 50, 0.9, init_soc=0.2), M.BatteryModule(10, 1000, 10, 10, 0.7, init_soc=0.3),
                            M.GensetModule(5, 40, 0.3), ("pv", M.RenewableModule(time_series=pv, **ts)),
                            M.LoadModule(time_series=load, **ts), M.GridModule(100, 100, g, **ts)], **kw)
    assert mg.get_forecast_horizon() == 4
    act = {"battery": [12.0, -3.0], "genset": [np.array([1.0, 22.0])], "grid": [40.0]}
    nrm = mg.to_normalized(act, act=True)
    assert nrm["battery"][0] == (12.0 - (-50 / 0.9)) / (50 * 0.9 + 50 / 0.9) and nrm["grid"][0] == 140.0 / 200.0     # battery_module.py:332-338
    assert np.array_equal(nrm["genset"][0], [1.0, 22.0 / 40.0])
    back = mg.from_normalized(nrm, act=True)
    assert np.allclose(back["battery"], act["battery"], rtol=0, atol=1e-12) and np.allclose(back["genset"][0], act["genset"][0], rtol=0, atol=1e-12)
    grid = mg.modules.grid[0]
    assert np.array_equal(grid.import_price, g[:5, 0]) and np.array_equal(grid.export_price, g[:5, 1])
    assert np.array_equal(grid.co2_per_kwh, g[:5, 2]) and np.array_equal(grid.grid_status, np.ones(5))
    obs = mg.reset()
    state = mg.from_normalized({k: v for k, v in obs.items() if k not in ("balance", "other")}, obs=True)
    assert np.allclose(state["load"][0], -load[:5], rtol=0, atol=1e-9) and np.allclose(state["pv"][0], pv[:5], rtol=0, atol=1e-9)
    mixed = ComposedMicrogrid([M.LoadModule(time_series=load), M.RenewableModule(time_series=pv, **ts)], **kw)
    with pytest.raises(ValueError):
        mixed.get_forecast_horizon()


def check_batch_log_recorder(lib):
    """ComposedBatch.recorder(env_ids): the reference-format log of selected envs of a batch == the get_log() of a single
    microgrid stepped with the same actions (continuous and discrete steps, a masked reset in between)"""
    hostsim_select(lib)
    kw = {}
    case, batch, _ = _batch_case(lib, "several_of_each", 9, 13)
    rec = batch.recorder([0, 4, 8])
    singles = {e: ComposedMicrogrid(batch.compositions[int(batch.env_config[e])].records_named(), add_unbalanced_module=False,
                                    obs_order="container", **kw) for e in (0, 4, 8)}
    for e, mg in singles.items():        # same live state as the batch env
        for a in ("fstate", "istate"):
            getattr(mg._batch, a)[0].copy_(getattr(batch, a)[e])
    rng = np.random.default_rng(2)
    comp = batch.comp
    n_lists = len(batch.action_lists)
    for k in range(9):
        if k == 5:
            mask = np.zeros(9, dtype=np.uint8)
            mask[4] = 1
            rec.reset(torch.from_numpy(mask).to(batch.device))
            singles[4].reset()
        if k % 3 == 2:
            d = rng.integers(0, n_lists, 9).astype(np.int32)
            rec.step_discrete(torch.from_numpy(d).to(batch.device))
            for e, mg in singles.items():
                mg.run_priority_list(int(d[e]), 1)
        else:
            a = rng.random((9, comp.n_act))
            rec.step(torch.from_numpy(a).to(batch.device))
            for e, mg in singles.items():
                mg.run({name: [a[e, s.act_col:s.act_col + s.n_act] if s.n_act == 2 else a[e, s.act_col] for s in slots]
                        for name, slots in comp.controllable()})
    for e, mg in singles.items():
        got, want = rec.get_log(e), mg.get_log()
        assert list(got.columns) == list(want.columns) and list(got.index) == list(want.index), e
        assert np.array_equal(got.to_numpy(dtype=np.float64), want.to_numpy(dtype=np.float64), equal_nan=True), e
    assert len(rec.get_log(4)) == 4 and len(rec.get_log(0)) == 9


def check_forecast_noise(lib):
    """Gaussian-noise forecasters on a composed batch (mgc_forecast_noise): the per-element standard deviations equal the
    reference-pinned restatement of GaussianNoiseForecaster (oracle/forecast_noise.py), and every noisy observation equals
    clip(clean + z * sigma * scale) with z from the numpy restatement of the kernel's Philox / Box-Muller stream."""
    hostsim_select(lib)
    from oracle.forecast_noise import NoisyModule
    from pymgrid_b200 import modules as M
    from tests.helpers import engine_noise_normals
    kw = {}
    rng = np.random.default_rng(8)
    T = 40
    load, pv = 100 + 100 * rng.random(T), 200 * rng.random(T)
    g = np.stack([rng.uniform(0.05, 0.9, T), np.zeros(T), rng.uniform(0, 0.6, T), (rng.random(T) > 0.3).astype(float)], axis=1)

    def mods(noisy):
        f = (lambda std: std) if noisy else (lambda std: "oracle")
        return [M.LoadModule(time_series=load, forecaster=f(12.5), forecast_horizon=6, forecaster_relative_noise=False),
                M.LoadModule(time_series=0.5 * load, forecaster="oracle", forecast_horizon=6),
                M.RenewableModule(time_series=pv, forecaster=f(0.3), forecast_horizon=9, forecaster_increase_uncertainty=True,
                                  forecaster_relative_noise=True),
                M.BatteryModule(10, 100, 50, 50, 0.9, init_soc=0.5),
                M.GridModule(100, 100, g, forecaster=f(0.02), forecast_horizon=4, forecaster_relative_noise=True)]
    n = 7
    noisy = ComposedBatch([mods(True)], np.zeros(n, dtype=np.int64), obs_order="container", **kw)
    clean = ComposedBatch([mods(False)], np.zeros(n, dtype=np.int64), obs_order="container", **kw)
    comp = noisy.comp
    # sigma table against the reference-pinned restatement
    rows = host(noisy._noise_rows)[0]
    sigma, inc = rows[:comp.obs_dim], rows[comp.obs_dim:]
    want = {("load", 0): NoisyModule(-load, 6, True, 12.5, False, False, 0, T), ("renewable", 0): NoisyModule(pv, 9, True, 0.3, True, True, 0, T),
            ("grid", 0): NoisyModule(g, 4, False, 0.02, False, True, 0, T)}
    for s in comp.slots:
        block = sigma[s.obs_off:s.obs_off + s.obs_len]
        if (s.name, s.index) in want:
            tab = want[(s.name, s.index)].sigma_normalised()          # [H, C], incl. the 1 + log(1 + k) growth
            C_ = tab.shape[1]
            assert np.array_equal(block[:C_], np.zeros(C_))            # the current value carries no noise
            base = block[C_:].reshape(-1, C_)
            grow = (1 + np.log(1 + np.arange(len(base))))[:, None] if inc[s.obs_off + C_] else 1.0
            np.testing.assert_allclose(base * grow, tab, rtol=1e-15, atol=0)
        else:
            assert not block.any()
    assert sigma[comp.slots[-1].obs_off + 4 + 1] == 0.0                # export price is constant: the clip pins it
    # every element against the generator restatement, across the end of the series (padding rows carry no noise)
    noisy.set_forecast_noise(seed=77, env_offset=1000)
    for b in (noisy, clean):
        b.step_counter.fill_(T - 7)
    acts = rng.random((6, n, comp.n_act))
    call = 0
    for k in range(6):
        o_noisy = host(noisy.step(torch.from_numpy(acts[k]).to(noisy.device))[0]).copy()
        o_clean = host(clean.step(torch.from_numpy(acts[k]).to(clean.device))[0]).copy()
        call += 1
        t = T - 7 + k + 1
        for e in range(n):
            z = engine_noise_normals(1000 + e, t, (comp.obs_dim + 1) // 2, 77, call)
            expect = o_clean[e].copy()
            for s in comp.slots:
                if s.kind not in ("load", "renewable", "grid"):
                    continue
                C_ = 4 if s.kind == "grid" else 1
                for kk in range(C_, s.obs_len):
                    j, row = s.obs_off + kk, kk // C_ - 1
                    if sigma[j] == 0 or t + 1 + row >= T:
                        continue
                    scale = 1 + np.log(1 + row) if inc[j] else 1.0
                    expect[j] = min(max(o_clean[e, j] + z[j] * sigma[j] * scale, 0.0), 1.0)
            np.testing.assert_allclose(o_noisy[e], expect, rtol=0, atol=1e-12)
        assert not np.array_equal(o_noisy, o_clean) or t + 1 >= T
    assert np.array_equal(host(noisy.reward), host(clean.reward))      # noise never touches the physics
    # reproducible per seed, different per env and per call; rollouts keep the oracle forecast
    a1 = host(noisy.observe()).copy()
    noisy.set_forecast_noise(seed=77, env_offset=1000)
    noisy._noise_calls = call
    assert np.array_equal(host(noisy.observe()), a1)
    noisy.clear_forecast_noise()
    assert np.array_equal(host(noisy.observe()), host(clean.observe()))


def check_modules_step_batch(lib):
    """mgc_modules_step on a batch spanning several tiles: every env equals a batch of one stepped with the same actions"""
    hostsim_select(lib)
    case = next(c for c in CASES if c.label == "several_of_each")
    kw = {}
    make = lambda k: ComposedBatch([case.modules()], np.zeros(k, dtype=np.int64), obs_order="container", with_info=True,      # noqa: E731
                                   microgrid_kwargs=case.microgrid_kwargs, **kw)
    n = 300
    batch = make(n)
    W = sum(2 if s.kind == "genset" else 0 if s.kind == "load" else 1 for s in batch.comp.slots)
    rng = np.random.default_rng(0)
    a1, a2 = rng.random((n, W)), rng.random((n, W)) * 10
    a2[:, 0] = rng.random(n)            # the genset goals stay in [0, 1]
    goal_cols = []
    col = 0
    for s in batch.comp.dispatch:
        if s.kind == "genset":
            goal_cols.append(col)
        col += 2 if s.kind == "genset" else 0 if s.kind == "load" else 1
    a2[:, goal_cols] = rng.random((n, len(goal_cols)))
    batch.modules_step(torch.from_numpy(a1).to(batch.device), normalized=True)
    obs, reward, _, info = batch.modules_step(torch.from_numpy(a2).to(batch.device), normalized=False)
    obs, reward, info = host(obs).copy(), host(reward).copy(), host(info).copy()
    for e in (0, 127, 128, 299):
        one = make(1)
        one.modules_step(torch.from_numpy(a1[e:e + 1]).to(one.device), normalized=True)
        o, r, _, i = one.modules_step(torch.from_numpy(a2[e:e + 1]).to(one.device), normalized=False)
        assert np.array_equal(host(o)[0], obs[e], equal_nan=True) and np.array_equal(host(r)[0], reward[e], equal_nan=True)
        assert np.array_equal(host(i)[0], info[e], equal_nan=True)


def check_control_dict_conventions(lib):
    """Microgrid.run's control conventions (microgrid.py:262-284): a bare scalar stands for [scalar], unknown keys warn,
    a missing controllable module raises ValueError"""
    hostsim_select(lib)
    import warnings
    from pymgrid_b200 import modules as M
    kw = {}
    rng = np.random.default_rng(1)

    def build():
        r = np.random.default_rng(1)
        return ComposedMicrogrid([M.LoadModule(time_series=50 + 50 * r.random(20)), M.LoadModule(time_series=20 * r.random(20)),
                                  M.RenewableModule(time_series=80 * r.random(20)), M.BatteryModule(10, 100, 50, 50, 0.9, init_soc=0.5),
                                  M.GridModule(100, 50, np.ones((20, 3)) * [0.2, 0.1, 0.3])], **kw)
    a, b = build(), build()
    for _ in range(3):
        x, y = float(rng.random()), float(rng.random())
        assert a.run({"battery": x, "grid": y})[1] == b.run({"battery": [x], "grid": [y]})[1]
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        a.run({"battery": [0.3], "grid": [0.2], "extra": [1.0]})
    assert any("Ignoring the following keys" in str(w.message) for w in caught)
    with pytest.raises(ValueError, match='Control for module "grid" not found'):
        a.run({"battery": [0.3]})
