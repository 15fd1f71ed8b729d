"""GPU parity: the CUDA path (through the C-ABI) against the golden vectors recorded from the reference and against
the C oracle on the same seeded inputs.  Bar: bit-exact for integers (step, done, genset status) AND for floats --
the kernels are compiled with -fmad=false so every f64 operation matches the reference's un-fused arithmetic;
the assertions below therefore use exact equality (stricter than the 1e-6 relative the spec allows)."""
import copy

import numpy as np
import pytest
import torch

from oracle.oracle import OracleBatch
from pymgrid_b200.scenario import load_pymgrid25
from tests.helpers import custom_params, fuzz_params, jump_to, overfull_params

pytestmark = pytest.mark.gpu
CONTAINER = ("genset", "battery", "grid")


def engine(configs, env_config, **kw):
    from pymgrid_b200.engine import BatchedMicrogrid
    kw.setdefault("action_order", CONTAINER)
    kw.setdefault("with_info", True)
    return BatchedMicrogrid(configs, env_config, device="cuda:0", **kw)


def group_actions(bm, per_env_actions):
    """per_env_actions: list (env order) of 1-D arrays in container order -> one device tensor per group."""
    out = []
    for g in bm.groups:
        a = np.stack([per_env_actions[e] for e in g.env_ids])
        out.append(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    return out


def gather(bm, per_group):
    """per-group device tensors -> list in env order of numpy arrays."""
    out = [None] * bm.n_envs
    for g, x in zip(bm.groups, per_group):
        x = x.cpu().numpy()
        for slot, e in enumerate(g.env_ids):
            out[e] = x[slot]
    return out


def state_rows(bm):
    rows = [None] * bm.n_envs
    for gi, g in enumerate(bm.groups):
        step, charge = g.step.cpu().numpy(), g.charge.cpu().numpy()
        gen = bm.genset_status(gi).cpu().numpy() if g.genset is not None else np.zeros((g.n_envs, 4), dtype=np.int64)
        for slot, e in enumerate(g.env_ids):
            rows[e] = np.array([step[slot], charge[slot], *gen[slot]], dtype=np.float64)
    return rows


def as_lists(res):
    obs, reward, done, info = res
    if isinstance(obs, torch.Tensor):
        return [obs], [reward], [done], [info]
    return obs, reward, done, info


def check_against_golden(bm, seq, normalized=True):
    """seq: per env dict(a, r, d, o, i, s) of golden arrays; all envs have the same number of steps."""
    n_steps = len(seq[0]["a"])
    for k in range(n_steps):
        obs, reward, done, info = as_lists(bm.step(group_actions(bm, [s["a"][k] for s in seq]), normalized=normalized))
        o, r, d, inf, st = gather(bm, obs), gather(bm, reward), gather(bm, done), gather(bm, info), state_rows(bm)
        for e, s in enumerate(seq):
            assert r[e] == s["r"][k], (e, k, r[e], s["r"][k])
            assert bool(d[e]) == bool(s["d"][k]), (e, k)
            np.testing.assert_array_equal(o[e], s["o"][k], err_msg=f"obs env {e} step {k}")
            np.testing.assert_array_equal(inf[e], s["i"][k], err_msg=f"info env {e} step {k}")
            np.testing.assert_array_equal(st[e], s["s"][k], err_msg=f"state env {e} step {k}")


def test_pymgrid25_golden_steps(golden):
    """All 25 scenarios side by side (three architecture groups in ONE launch): normalised steps from t=0, then
    unnormalised steps, bit-exact against the reference's recorded outputs."""
    z = golden["pymgrid25_steps"]
    configs = [load_pymgrid25(n) for n in range(25)]
    bm = engine(configs, np.arange(25))
    assert len(bm.groups) == 3
    seq = [dict(a=z[f"s{n}_a0"], r=z[f"s{n}_r0"], d=z[f"s{n}_d0"], o=z[f"s{n}_o0"], i=z[f"s{n}_i0"], s=z[f"s{n}_s0"]) for n in range(25)]
    check_against_golden(bm, seq)
    seq = [dict(a=z[f"s{n}_au"], r=z[f"s{n}_ru"], d=z[f"s{n}_du"], o=z[f"s{n}_ou"], i=z[f"s{n}_iu"], s=z[f"s{n}_su"]) for n in range(25)]
    check_against_golden(bm, seq, normalized=False)
    flags = torch.cat([g.flags for g in bm.groups]).cpu().numpy()
    assert (flags & 0x7f == 0).all()


def test_pymgrid25_golden_end_of_series(golden):
    """Across the end of the series: forecast padding, done at final_step-1, last valid step, then the past-the-end flag."""
    z = golden["pymgrid25_steps"]
    end_from = int(z["end_from"])
    configs = [jump_to(load_pymgrid25(n), end_from) for n in range(25)]
    bm = engine(configs, np.arange(25))
    reset_obs = gather(bm, bm.reset())
    for n in range(25):
        np.testing.assert_array_equal(reset_obs[n], z[f"s{n}_reset_obs"])
    seq = [dict(a=z[f"s{n}_a1"], r=z[f"s{n}_r1"], d=z[f"s{n}_d1"], o=z[f"s{n}_o1"], i=z[f"s{n}_i1"], s=z[f"s{n}_s1"]) for n in range(25)]
    check_against_golden(bm, seq)
    # one more step: the reference raises IndexError; the engine flags it and leaves the state untouched
    before = state_rows(bm)
    obs, reward, done, info = as_lists(bm.step(group_actions(bm, [s["a"][0] for s in seq])))
    after = state_rows(bm)
    for e in range(25):
        np.testing.assert_array_equal(before[e], after[e])
    for g, r, d in zip(bm.groups, reward, done):
        assert torch.isnan(r).all() and (d == 1).all() and ((g.flags & (1 << 5)) != 0).all()


@pytest.mark.parametrize("i", range(6))
def test_custom_grids_golden(golden, i):
    """Slow gensets (start-up / wind-down, abortion on and off), weak grid, short series, no forecaster."""
    z = golden["custom"]
    p = custom_params(z, i)
    bm = engine([p], np.zeros(3, dtype=np.int64))        # three replicas fed the same actions
    ro = bm.reset()
    for e in range(3):
        np.testing.assert_array_equal(ro[e].cpu().numpy(), z[f"c{i}_reset_obs"])
    seq = [dict(a=z[f"c{i}_a"], r=z[f"c{i}_r"], d=z[f"c{i}_d"], o=z[f"c{i}_o"], i=z[f"c{i}_i"], s=z[f"c{i}_s"])] * 3
    check_against_golden(bm, seq)
    charge = bm.groups[0].charge.clone()
    ro = bm.reset()
    np.testing.assert_array_equal(ro[0].cpu().numpy(), z[f"c{i}_after_reset_obs"])
    assert torch.equal(charge, bm.groups[0].charge) and (bm.groups[0].step == 0).all()


@pytest.mark.parametrize("i", range(40))
def test_fuzz_grids_golden(golden, i):
    """Randomised constructor arguments recorded from the live reference (tests/golden/make_fuzz.py): every architecture
    (with and without genset / grid), horizons 0 .. 40, initial_step > 0, final_step inside the series, min_capacity 0,
    running_min_production 0 or == max, max_export 0, init_soc that does not survive * max_capacity / max_capacity.
    Continuous steps (normalised, then unnormalised beyond every limit), the discrete env with slow gensets, and
    rule-based control as one persistent rollout -- all bit for bit."""
    z = golden["fuzz"]
    g = lambda k: z[f"f{i}_{k}"]  # noqa: E731
    p = fuzz_params(z, i)
    bm = engine([p], np.zeros(2, dtype=np.int64), with_flags=True)       # two replicas fed the same actions
    ro = bm.reset()
    for e in range(2):
        np.testing.assert_array_equal(ro[e].cpu().numpy(), g("reset_obs"))
    np.testing.assert_array_equal(bm.observe()[1].cpu().numpy(), g("reset_obs"))
    for seg, normalized in (("n", True), ("u", False)):
        n = len(g(f"{seg}_r"))
        seq = [dict(a=g(f"{seg}_a")[:n], r=g(f"{seg}_r"), d=g(f"{seg}_d"), o=g(f"{seg}_o"), i=g(f"{seg}_i"), s=g(f"{seg}_s"))] * 2
        check_against_golden(bm, seq, normalized=normalized)
        err = int(g(f"{seg}_err"))
        if err >= 0:     # the reference raised AssertionError (base_module.py:272) on this action
            assert err == n
            bm.step(group_actions(bm, [g(f"{seg}_a")[err]] * 2), normalized=normalized)
            assert ((bm.groups[0].flags.cpu().numpy() & (1 << 4)) != 0).all()
            break
    # DiscreteMicrogridEnv: host-side action table == the reference's, expansion + step on the device
    bm = engine([fuzz_params(z, i)], np.zeros(2, dtype=np.int64))
    np.testing.assert_array_equal(bm.reset()[0].cpu().numpy(), g("d_reset_obs"))
    mod, act = g("d_table_mod"), g("d_table_act")
    table = bm.action_tables[0]
    assert len(table) == len(mod)
    for row, pl in enumerate(table):
        assert [tuple(x) for x in pl] == [(int(m), int(a)) for m, a in zip(mod[row], act[row]) if m >= 0]
    for k, a in enumerate(g("d_actions")):
        obs, reward, done, _ = bm.step_discrete(torch.full((2,), int(a), dtype=torch.int32, device="cuda"))
        assert reward[1].item() == g("d_rewards")[k] and bool(done[1].item()) == bool(g("d_dones")[k]), k
        np.testing.assert_array_equal(obs[1].cpu().numpy(), g("d_obs")[k], err_msg=f"discrete step {k}")
    np.testing.assert_array_equal(state_rows(bm)[1], g("d_state"))
    # RuleBasedControl: the sorted list chosen on the host, the whole run in one persistent kernel
    bm = engine([fuzz_params(z, i)], np.zeros(2, dtype=np.int64))
    want = [(int(m), int(a)) for m, a in zip(g("rbc_list_mod"), g("rbc_list_act"))]
    assert [tuple(x) for x in bm.action_tables[0][int(bm.rbc_actions()[0][0].item())]] == want
    out = bm.rollout_rbc(len(g("rbc_rewards")), keep_obs=False)
    np.testing.assert_array_equal(out["reward"][:, 1].cpu().numpy(), g("rbc_rewards"))
    assert out["done"][-1].all() and not out["done"][:-1].any()
    np.testing.assert_array_equal(state_rows(bm)[0], g("rbc_final_state"))


def test_overfull_battery_is_flagged_where_the_reference_asserts(golden):
    """charge one ulp above max_capacity: MG_FLAG_NEGATIVE_ABSORB where the reference raises AssertionError
    (base_module.py:272 on a continuous charge request, priority_list.py:124 on a list that makes the battery absorb);
    discharging and the list that lets the grid take the surplus first run normally, bit for bit."""
    z = golden["fuzz"]
    act = lambda a, b: torch.tensor([[a, b]] * 3, dtype=torch.float64, device="cuda")   # noqa: E731  container order: battery, grid
    bm = engine([overfull_params(z)], np.zeros(3, dtype=np.int64), with_flags=True)
    bm.step(act(-10.0, 0.0), normalized=False)
    assert ((bm.groups[0].flags & (1 << 4)) != 0).all()
    bm = engine([overfull_params(z)], np.zeros(3, dtype=np.int64), with_flags=True)
    obs, reward, _, _ = bm.step(act(5.0, -9.0), normalized=False)
    assert ((bm.groups[0].flags & 0x7f) == 0).all() and reward[2].item() == float(z["over_discharge_reward"])
    np.testing.assert_array_equal(obs[2].cpu().numpy(), z["over_discharge_obs"])
    np.testing.assert_array_equal(state_rows(bm)[2], z["over_discharge_state"])
    for a, raised in enumerate(z["over_discrete_raised"]):
        bm = engine([overfull_params(z)], np.zeros(3, dtype=np.int64), with_flags=True)
        _, reward, _, _ = bm.step_discrete(torch.full((3,), a, dtype=torch.int32, device="cuda"))
        flagged = (bm.groups[0].flags & (1 << 4)) != 0
        if str(raised):
            assert flagged.all()
        else:
            assert not flagged.any() and reward[1].item() == z["over_discrete_reward"][a]
    # the same through the persistent kernel: flags are OR-ed over the rollout
    bm = engine([overfull_params(z)], np.zeros(3, dtype=np.int64), with_flags=True)
    bm.rollout(torch.zeros((2, 3), dtype=torch.int32, device="cuda"), discrete=True, keep_obs=False)
    assert ((bm.groups[0].flags & (1 << 4)) != 0).all()


def _discrete_cases(z):
    return sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_actions")})


@pytest.mark.parametrize("horizon", (23, 24))
def test_discrete_env_golden(golden, horizon):
    """DiscreteMicrogridEnv: action tables, rewards and flat observations of the reference env (H=23: all 25
    scenarios; H=24 + grid: BASELINE config 4)."""
    z = golden["discrete"]
    scen = [int(t.split("_")[1][1:]) for t in _discrete_cases(z) if t.startswith(f"h{horizon}_")]
    configs = []
    for n in scen:
        p = load_pymgrid25(n)
        p.forecast_horizon = horizon
        configs.append(p)
    bm = engine(configs, np.arange(len(scen)))
    for k, n in enumerate(scen):     # host-side action table == the reference's actions_list
        tag = f"h{horizon}_s{n}"
        mod, act = z[f"{tag}_table_mod"], z[f"{tag}_table_act"]
        table = bm.action_tables[k]
        assert len(table) == len(mod)
        for row, pl in enumerate(table):
            assert [tuple(x) for x in pl] == [(int(m), int(a)) for m, a in zip(mod[row], act[row]) if m >= 0]
    ro = gather(bm, bm.reset() if not bm.single_group else [bm.reset()])
    for k, n in enumerate(scen):
        np.testing.assert_array_equal(ro[k], z[f"h{horizon}_s{n}_reset_obs"])
    n_steps = len(z[f"h{horizon}_s{scen[0]}_actions"])
    for s in range(n_steps):
        acts = []
        for g in bm.groups:
            acts.append(torch.tensor([z[f"h{horizon}_s{scen[e]}_actions"][s] for e in g.env_ids], dtype=torch.int32, device="cuda"))
        obs, reward, done, info = as_lists(bm.step_discrete(acts))
        o, r, d = gather(bm, obs), gather(bm, reward), gather(bm, done)
        for k, n in enumerate(scen):
            tag = f"h{horizon}_s{n}"
            assert r[k] == z[f"{tag}_rewards"][s], (tag, s)
            assert bool(d[k]) == bool(z[f"{tag}_dones"][s])
            np.testing.assert_array_equal(o[k], z[f"{tag}_obs"][s], err_msg=f"{tag} step {s}")


def randomise_state(bm, rng, configs, env_config, max_t):
    """Give every env its own step / charge / genset status so the batch is not in lock-step."""
    plist = []
    for e, c in enumerate(env_config):
        p = copy.copy(configs[c])
        p.battery = copy.copy(p.battery)
        p.current_step = int(rng.integers(0, max_t))
        p.battery.current_charge = float(rng.uniform(p.battery.min_capacity, p.battery.max_capacity))
        if p.genset is not None:
            p.genset = copy.copy(p.genset)
            cs = int(rng.integers(0, 2))
            p.genset.current_status = p.genset.goal_status = cs
            p.genset.steps_until_up, p.genset.steps_until_down = (0, p.genset.wind_down_time) if cs else (p.genset.start_up_time, 0)
        plist.append(p)
    for gi, g in enumerate(bm.groups):
        g.step.copy_(torch.tensor([plist[e].current_step for e in g.env_ids], dtype=torch.int32))
        g.charge.copy_(torch.tensor([plist[e].battery.current_charge for e in g.env_ids], dtype=torch.float64))
        if g.genset is not None:
            g.genset.copy_(torch.tensor([plist[e].genset.current_status * 0x101 | (plist[e].genset.steps_until_up << 16)
                                         | (plist[e].genset.steps_until_down << 24) for e in g.env_ids], dtype=torch.int32))
    return plist


@pytest.mark.parametrize("order", ("gym_sorted", "container"))
def test_batch_vs_oracle_ragged_state(order):
    """4099 envs (ragged last tile) over all 25 scenarios, every env at its own step (some inside the last 30
    steps of the year), own charge, own genset status; 12 normalised steps + 6 unnormalised, engine == oracle."""
    rng = np.random.default_rng(11)
    configs = [load_pymgrid25(n) for n in range(25)]
    B = 4099
    env_config = rng.integers(0, 25, B)
    bm = engine(configs, env_config, obs_order=order)
    plist = randomise_state(bm, rng, configs, env_config, 8740)
    for e in range(0, B, 7):          # a slice of envs starts close to the end of the series
        plist[e].current_step = int(rng.integers(8725, 8742))
    for gi, g in enumerate(bm.groups):
        g.step.copy_(torch.tensor([plist[e].current_step for e in g.env_ids], dtype=torch.int32))
    ob = OracleBatch(plist, order=0 if order == "gym_sorted" else 1)
    for normalized, n_steps in ((True, 12), (False, 6)):
        acts = []
        for e in range(B):
            p = plist[e]
            a = rng.random((n_steps, p.n_act))
            if not normalized:
                col = 0
                if p.genset is not None:
                    a[:, 0] = rng.integers(0, 2, n_steps)
                    a[:, 1] = rng.uniform(0, 1.3, n_steps) * p.genset.running_max_production
                    col = 2
                a[:, col] = rng.uniform(-1.5, 1.5, n_steps) * p.battery.max_charge
                if p.grid is not None:
                    a[:, col + 1] = rng.uniform(-1.2, 1.2, n_steps) * p.grid.max_import
            acts.append(a)
        padded = np.zeros((n_steps, B, 4))
        for e in range(B):
            padded[:, e, :acts[e].shape[1]] = acts[e]
        for k in range(n_steps):
            obs, reward, done, info = as_lists(bm.step(group_actions(bm, [a[k] for a in acts]), normalized=normalized))
            o_rew, o_done, o_obs = ob.rollout(padded[k:k + 1], normalized=normalized, n_threads=4)
            r, d, o = gather(bm, reward), gather(bm, done), gather(bm, obs)
            np.testing.assert_array_equal(np.array(r), o_rew[0])
            np.testing.assert_array_equal(np.array(d), o_done[0])
            for e in range(B):
                np.testing.assert_array_equal(o[e], o_obs[e, :len(o[e])], err_msg=f"env {e} step {k}")
        t, charge, gen = ob.state()
        st = state_rows(bm)
        np.testing.assert_array_equal(np.array([s[0] for s in st]), t)
        np.testing.assert_array_equal(np.array([s[1] for s in st]), charge)
        np.testing.assert_array_equal(np.array([s[2:] for s in st]), gen)


EMITTER_VARIANTS = [("lsu", 0, True), ("lsu", 0, False)] + [("image", shape, ws) for shape in range(6) for ws in (True, False)]


def set_emitter(bm, emit, shape, specialised):
    bm.set_emit_image(emit == "image")
    bm.set_image_shape(shape)
    bm.set_rollout_specialised(specialised)


@pytest.mark.parametrize("emit,shape,specialised", EMITTER_VARIANTS)
def test_row_emitters_agree_on_ragged_batches(emit, shape, specialised):
    """Every row emitter (per-lane 16-byte stores; shared-memory images + TMA bulk stores in each instantiated shape, with and
    without the owner / emitter warp split) writes the same bytes: 3 001 envs (a partial last tile in every group) at
    unrelated steps, some inside the last steps of the year; the persistent kernel cut into two launches (ring 1 and ring
    5), single steps, observe and masked reset, against the C oracle."""
    rng = np.random.default_rng(23)
    configs = [load_pymgrid25(n) for n in range(25)]
    B, n_steps = 3001, 11
    env_config = rng.integers(0, 25, B)
    bm = engine(configs, env_config)
    set_emitter(bm, emit, shape, specialised)
    plist = randomise_state(bm, rng, configs, env_config, 8740)
    for e in range(0, B, 5):          # (n_steps more steps end at the last valid step of the year at the latest)
        plist[e].current_step = int(rng.integers(8738, 8749))
    for g in bm.groups:
        g.step.copy_(torch.tensor([plist[e].current_step for e in g.env_ids], dtype=torch.int32))
    ob = OracleBatch(plist)
    padded = np.zeros((n_steps, B, 4))
    for e in range(B):
        padded[:, e, :plist[e].n_act] = rng.random((n_steps, plist[e].n_act))
    acts = [torch.from_numpy(np.ascontiguousarray(padded[:, g.env_ids, :g.n_act])).cuda() for g in bm.groups]
    o_rew, o_done, o_obs = ob.rollout(padded[:4], n_threads=4)
    out = bm.rollout([a[:4].contiguous() for a in acts], ring=1)
    for g, r in zip(bm.groups, out):
        np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:, g.env_ids])
        np.testing.assert_array_equal(r["done"].cpu().numpy(), o_done[:, g.env_ids])
        np.testing.assert_array_equal(r["obs_ring"][0].cpu().numpy(), o_obs[g.env_ids][:, :g.obs_dim])
    o_rew, o_done, o_obs = ob.rollout(padded[4:10], n_threads=4)
    out = bm.rollout([a[4:10].contiguous() for a in acts], ring=5)
    for g, r in zip(bm.groups, out):
        np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:, g.env_ids])
        np.testing.assert_array_equal(r["obs_ring"][5 % 5].cpu().numpy(), o_obs[g.env_ids][:, :g.obs_dim])
    o_rew, o_done, o_obs = ob.rollout(padded[10:11], n_threads=4)
    obs, reward, done, _ = as_lists(bm.step([a[10].contiguous() for a in acts]))
    for g, o, r in zip(bm.groups, obs, reward):
        np.testing.assert_array_equal(r.cpu().numpy(), o_rew[0, g.env_ids])
        np.testing.assert_array_equal(o.cpu().numpy(), o_obs[g.env_ids][:, :g.obs_dim])
    stepped = [o.clone() for o in obs]
    for o in obs:
        o.zero_()
    for o, want in zip(bm.observe(), stepped):
        assert torch.equal(o, want)
    # masked reset: the selected envs go back to their initial step, every row is rewritten
    masks = [torch.from_numpy((rng.random(g.n_envs) < 0.5).astype(np.uint8)).cuda() for g in bm.groups]
    before = [g.step.clone() for g in bm.groups]
    rows = bm.reset(mask=masks)
    ref = engine(configs, env_config)
    ref.set_emit_image(False)
    ref.load_state_dict(bm.state_dict())
    for g, gr, m, b0, row in zip(bm.groups, ref.groups, masks, before, rows):
        assert torch.equal(g.step, torch.where(m.bool(), torch.zeros_like(b0), b0))
        assert torch.equal(row, ref.observe()[ref.groups.index(gr)])


@pytest.mark.parametrize("emit,shape,specialised", EMITTER_VARIANTS)
def test_row_emitters_agree_on_generator_grids(emit, shape, specialised):
    """The same for heterogeneous MicrogridGenerator grids (per-env series in shared-memory rings, per-env grid-status bits):
    2 500 grids started at unrelated steps incl. the end of the year, persistent kernel and single steps against the oracle."""
    from pymgrid_b200 import generator
    from tests.test_generator import pv_first
    gb = generator.sample(2500, seed=5)
    bm = generator.engine_from_batch(gb, device="cuda:0", action_order=CONTAINER)
    set_emitter(bm, emit, shape, specialised)
    rng = np.random.default_rng(9)
    plist = [gb.to_params(i) for i in range(gb.n)]
    starts = rng.integers(0, 8700, gb.n)
    starts[::4] = rng.integers(8740, 8752, len(starts[::4]))
    for p, s0 in zip(plist, starts):
        p.current_step = int(s0)
    for g in bm.groups:
        g.step.copy_(torch.from_numpy(starts[g.env_ids].astype(np.int32)))
    ob = OracleBatch(plist)
    n_steps = 9
    padded = np.zeros((n_steps, gb.n, 4))
    for e, p in enumerate(plist):
        padded[:, e, :p.n_act] = rng.random((n_steps, p.n_act))
    acts = [torch.from_numpy(np.ascontiguousarray(padded[:, g.env_ids, :g.n_act])).cuda() for g in bm.groups]
    o_rew, o_done, o_obs = ob.rollout(padded[:8], n_threads=4)
    out = bm.rollout([a[:8].contiguous() for a in acts], ring=3)
    for g, r in zip(bm.groups, out):
        np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:, g.env_ids])
        np.testing.assert_array_equal(r["done"].cpu().numpy(), o_done[:, g.env_ids])
        got = r["obs_ring"][7 % 3].cpu().numpy()
        for slot, e in enumerate(g.env_ids):
            np.testing.assert_array_equal(got[slot], pv_first(o_obs[e, :g.obs_dim], plist[e]), err_msg=f"env {e}")
    o_rew, o_done, o_obs = ob.rollout(padded[8:9], n_threads=4)
    obs, reward, done, _ = as_lists(bm.step([a[8].contiguous() for a in acts]))
    for g, o, r in zip(bm.groups, obs, reward):
        np.testing.assert_array_equal(r.cpu().numpy(), o_rew[0, g.env_ids])
        got = o.cpu().numpy()
        for slot, e in enumerate(g.env_ids):
            np.testing.assert_array_equal(got[slot], pv_first(o_obs[e, :g.obs_dim], plist[e]), err_msg=f"env {e}")


@pytest.mark.parametrize("horizon", (1, 3, 8, 24, 31, 40))
@pytest.mark.parametrize("order", ("gym_sorted", "container"))
def test_row_emitters_over_horizons_and_row_layouts(horizon, order):
    """Both emitter families over the row layouts they must handle: forecast horizons from 1 (windows much shorter than a warp)
    through 24 (odd number of forecast rows) and 31 (a full warp of window elements) to 40 (no image layout: the library falls
    back to the per-lane emitters on its own), both observation orders, the three architectures in one batch with partial
    last tiles, envs at unrelated steps -- against the C oracle, bit for bit."""
    rng = np.random.default_rng(100 + horizon)
    configs = []
    for n in (0, 1, 2, 5, 9):
        p = copy.copy(load_pymgrid25(n))
        p.forecast_horizon = horizon
        configs.append(p)
    B, n_steps = 777, 6
    env_config = rng.integers(0, len(configs), B)
    outs = {}
    for emit in ("lsu", "image", "auto"):
        bm = engine(configs, env_config, obs_order=order)
        if emit != "auto":
            bm.set_emit_image(emit == "image")
        plist = randomise_state(bm, rng if emit == "lsu" else np.random.default_rng(5), configs, env_config, 8700)
        if emit == "lsu":
            state = bm.state_dict()
            ob = OracleBatch(plist, order=0 if order == "gym_sorted" else 1)
            padded = np.zeros((n_steps, B, 4))
            for e in range(B):
                padded[:, e, :plist[e].n_act] = rng.random((n_steps, plist[e].n_act))
            o_rew, o_done, o_obs = ob.rollout(padded, n_threads=4)
        else:
            bm.load_state_dict(state)
        acts = [torch.from_numpy(np.ascontiguousarray(padded[:, g.env_ids, :g.n_act])).cuda() for g in bm.groups]
        out = bm.rollout([a[:n_steps - 1].contiguous() for a in acts], ring=2)
        obs, reward, done, _ = as_lists(bm.step([a[n_steps - 1].contiguous() for a in acts]))
        for g, r, o, rw in zip(bm.groups, out, obs, reward):
            np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:n_steps - 1, g.env_ids], err_msg=emit)
            np.testing.assert_array_equal(rw.cpu().numpy(), o_rew[n_steps - 1, g.env_ids], err_msg=emit)
            np.testing.assert_array_equal(o.cpu().numpy(), o_obs[g.env_ids][:, :g.obs_dim], err_msg=emit)
        outs[emit] = [r["obs_ring"].clone() for r in out]
        if emit == "image" and horizon <= 31:
            assert bm.last_kernel == "mg_step_img_kernel"
        if horizon > 31:
            assert bm.last_kernel == "mg_step_kernel"          # no image layout for these rows
    for a, b, c in zip(outs["lsu"], outs["image"], outs["auto"]):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_automatic_emitter_choice():
    """MG_OPT_EMIT_IMAGE = 2 (the default): the library picks the emitters per launch and says which kernel ran
    (mg_last_kernel).  Large table-backed batches in lock-step keep the per-lane store emitters, the ragged hint (set by a
    masked reset) and small batches switch to the image emitters; results are the same either way (tests above)."""
    configs = [load_pymgrid25(n) for n in range(25)]
    big = engine(configs, np.arange(20000) % 25, with_info=False)          # 313+ tiles > 2 per SM
    acts = [torch.rand((2, g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in big.groups]
    big.rollout(acts, ring=1)
    assert big.last_kernel == "mg_rollout_ws_kernel"
    big.step([a[0].contiguous() for a in acts])
    assert big.last_kernel == "mg_step_kernel"
    big.reset(mask=[torch.ones(g.n_envs, dtype=torch.uint8, device="cuda") for g in big.groups])      # leaves lock-step
    big.rollout(acts, ring=1)
    assert big.last_kernel == "mg_rollout_img_kernel (owner / emitter warps)"
    big.step([a[0].contiguous() for a in acts])
    assert big.last_kernel == "mg_step_img_kernel"
    small = engine(configs, np.arange(3000) % 25, with_info=False)
    acts = [torch.rand((2, g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in small.groups]
    small.rollout(acts, ring=1)
    assert small.last_kernel.startswith("mg_rollout_img_kernel")
    small.set_emit_image(False)
    small.rollout(acts, ring=1)
    assert small.last_kernel == "mg_rollout_ws_kernel"


@pytest.mark.parametrize("specialised", [True, False])
def test_rollout_kernel_equals_repeated_steps(specialised):
    """mg_rollout (persistent kernel, state in registers; warp-specialised or plain, MG_OPT_ROLLOUT_SPECIALISED)
    == n_steps x mg_step, bit for bit, incl. the obs ring."""
    rng = np.random.default_rng(5)
    configs = [load_pymgrid25(n) for n in range(25)]
    B, n_steps, ring = 1000, 17, 4
    env_config = np.arange(B) % 25
    a = engine(configs, env_config)
    a.set_rollout_specialised(specialised)
    b = engine(configs, env_config)
    randomise_state(a, rng, configs, env_config, 8700)
    b.load_state_dict(a.state_dict())
    acts = [torch.rand((n_steps, g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in a.groups]
    out = a.rollout(acts, ring=ring)
    obs_hist = []
    for k in range(n_steps):
        obs, reward, done, _ = as_lists(b.step([x[k].contiguous() for x in acts]))
        obs_hist.append([o.clone() for o in obs])
        for gi in range(len(a.groups)):
            assert torch.equal(out[gi]["reward"][k], reward[gi])
            assert torch.equal(out[gi]["done"][k], done[gi])
    for gi in range(len(a.groups)):
        for k in range(n_steps - ring, n_steps):
            assert torch.equal(out[gi]["obs_ring"][k % ring], obs_hist[k][gi])
        assert torch.equal(a.groups[gi].step, b.groups[gi].step) and torch.equal(a.groups[gi].charge, b.groups[gi].charge)
        if a.groups[gi].genset is not None:
            assert torch.equal(a.groups[gi].genset, b.groups[gi].genset)


def test_discrete_rollout_vs_oracle():
    """Config 4 shape at small size: discrete actions, H=24, 15 grid scenarios, engine rollout == oracle rollout."""
    rng = np.random.default_rng(9)
    scen = [0, 4, 6, 11, 12, 14, 16, 1, 8, 9, 10, 13, 18, 22, 24]
    configs = []
    for n in scen:
        p = load_pymgrid25(n)
        p.forecast_horizon = 24
        configs.append(p)
    B, n_steps = 600, 30
    env_config = np.arange(B) % len(scen)
    bm = engine(configs, env_config)
    plist = randomise_state(bm, rng, configs, env_config, 8700)
    acts_env = np.stack([rng.integers(0, len(bm.action_tables[c]), n_steps) for c in env_config], axis=1).astype(np.int32)
    width = 3
    lut_mod, lut_act, offsets, off = [], [], {}, 0
    for c, table in enumerate(bm.action_tables):
        offsets[c] = off
        for pl in table:
            lut_mod.append([pl[j][0] if j < len(pl) else -1 for j in range(width)])
            lut_act.append([pl[j][1] if j < len(pl) else 0 for j in range(width)])
        off += len(table)
    ob = OracleBatch(plist)
    o_rew, o_done, o_obs = ob.rollout_discrete(acts_env, np.array(lut_mod), np.array(lut_act),
                                               np.array([offsets[c] for c in env_config]), n_threads=4)
    acts = [torch.from_numpy(np.ascontiguousarray(acts_env[:, g.env_ids])).cuda() for g in bm.groups]
    out = bm.rollout(acts, discrete=True, ring=1)
    out = [out] if isinstance(out, dict) else out
    for g, r in zip(bm.groups, out):
        np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:, g.env_ids])
        np.testing.assert_array_equal(r["done"].cpu().numpy(), o_done[:, g.env_ids])
        np.testing.assert_array_equal(r["obs_ring"][0].cpu().numpy(), o_obs[g.env_ids][:, :g.obs_dim])


@pytest.mark.parametrize("n", (0, 1, 2))
def test_full_year_golden(golden, n):
    """BASELINE config 1 on the GPU: the 8760-step year of scenarios 0 / 1 / 2 against the reference's rewards."""
    z = golden["pymgrid25_year"]
    p = load_pymgrid25(n)
    bm = engine([p], np.zeros(2, dtype=np.int64))
    a = np.random.default_rng(0).random((8760, p.n_act))
    acts = torch.from_numpy(np.ascontiguousarray(np.repeat(a[:, None, :], 2, axis=1))).cuda()
    out = bm.rollout(acts, ring=1, reward_sum=True)
    rewards = out["reward"].cpu().numpy()
    np.testing.assert_array_equal(rewards[:, 0], z[f"s{n}_rewards"])
    np.testing.assert_array_equal(rewards[:, 1], z[f"s{n}_rewards"])
    assert int(np.argmax(out["done"][:, 0].cpu().numpy())) == int(z[f"s{n}_first_done"])
    st = state_rows(bm)[0]
    np.testing.assert_array_equal(st, z[f"s{n}_final_state"])
    # sequential in-kernel sum == sequential host sum
    assert out["reward_sum"][0].item() == float(np.add.accumulate(z[f"s{n}_rewards"])[-1])


def test_full_size_properties():
    """BASELINE config 3 size (65 536 envs, 25 scenarios tiled): size-independent properties.
    (1) replicas of one scenario fed identical actions stay identical; (2) the energy balance closes:
    load_met + charge + export + overgeneration == pv_used + discharge + import + genset + loss_load;
    (3) sampled envs (one per scenario) equal the oracle; (4) obs in [0, 1]; (5) step counters advance by one."""
    B, n_steps = 65536, 6
    configs = [load_pymgrid25(n) for n in range(25)]
    env_config = np.arange(B) % 25
    bm = engine(configs, env_config)
    rng = np.random.default_rng(21)
    base = {n: rng.random((n_steps, configs[n].n_act)) for n in range(25)}
    sample = list(range(25))
    ob = OracleBatch([configs[env_config[e]] for e in sample])
    for k in range(n_steps):
        acts = []
        for g in bm.groups:
            cfg = env_config[g.env_ids]
            acts.append(torch.from_numpy(np.stack([base[c][k] for c in cfg])).cuda())
        obs, reward, done, info = as_lists(bm.step(acts))
        padded = np.zeros((1, len(sample), 4))
        for j, e in enumerate(sample):
            padded[0, j, :configs[env_config[e]].n_act] = base[env_config[e]][k]
        o_rew, o_done, o_obs = ob.rollout(padded)
        for g, o, r, inf in zip(bm.groups, obs, reward, info):
            cfg = torch.from_numpy(env_config[g.env_ids]).cuda()
            for c in torch.unique(cfg).tolist():
                rows = (cfg == c).nonzero()[:, 0]
                assert (r[rows] == r[rows[0]]).all() and (o[rows] == o[rows[0]]).all()
            lhs = inf[:, 0] + inf[:, 8] + inf[:, 10] + inf[:, 4]
            rhs = inf[:, 1] + inf[:, 7] + inf[:, 9] + inf[:, 5] + inf[:, 3]
            assert torch.allclose(lhs, rhs, rtol=1e-12, atol=1e-6)
            assert (o >= 0).all() and (o <= 1).all()
            assert (g.step == k + 1).all()
            for j, e in enumerate(sample):
                if bm.env_group[e] == bm.groups.index(g):
                    slot = bm.env_slot[e]
                    assert r[slot].item() == o_rew[0, j]
                    np.testing.assert_array_equal(o[slot].cpu().numpy(), o_obs[j, :g.obs_dim])
    assert bm.launch_count == 1 + n_steps


def test_full_size_generator_shard_properties():
    """BASELINE config 5, one GPU's shard at full size (131 072 heterogeneous MicrogridGenerator grids): size-independent
    properties of a 24-step persistent rollout through the kernel the library picks (images + bulk stores, rings, role split):
    (1) the same batch through the round-1 kernel family (per-lane stores) gives the same rewards, flags and rows, bit for bit;
    (2) rewards are finite, observations lie in [0, 1], step counters advance by the number of steps;
    (3) 192 sampled grids (every architecture, weak grids included) equal the C oracle on their explicit form."""
    from pymgrid_b200 import generator
    from tests.test_generator import pv_first
    B, n_steps = 131072, 24
    gb = generator.sample(B, seed=77)
    gen = torch.Generator(device="cuda")
    outs, engines = {}, {}
    for emit in ("auto", "lsu"):
        bm = generator.engine_from_batch(gb, device="cuda:0", action_order=CONTAINER, with_flags=True)
        if emit == "lsu":
            bm.set_emit_image(False)
        gen.manual_seed(12)
        acts = [torch.rand((n_steps, g.n_envs, g.n_act), dtype=torch.float64, device="cuda", generator=gen) for g in bm.groups]
        outs[emit] = bm.rollout(acts, ring=2)
        engines[emit] = (bm, acts)
        assert bm.last_kernel.startswith("mg_rollout_img_kernel" if emit == "auto" else "mg_rollout_kernel")
    bm, acts = engines["auto"]
    for g, a, b in zip(bm.groups, outs["auto"], outs["lsu"]):
        assert torch.equal(a["reward"], b["reward"]) and torch.equal(a["done"], b["done"]) and torch.equal(a["obs_ring"], b["obs_ring"])
        assert bool(torch.isfinite(a["reward"]).all())
        assert bool((a["obs_ring"] >= 0).all()) and bool((a["obs_ring"] <= 1).all())
        assert bool((g.step == n_steps).all())
    rng = np.random.default_rng(3)
    sample = np.sort(rng.choice(B, 192, replace=False))
    plist = [gb.to_params(int(i)) for i in sample]
    ob = OracleBatch(plist)
    padded = np.zeros((n_steps, len(sample), 4))
    for gi, g in enumerate(bm.groups):
        a = acts[gi].cpu().numpy()
        for j, e in enumerate(sample):
            if bm.env_group[e] == gi:
                padded[:, j, :g.n_act] = a[:, bm.env_slot[e]]
    o_rew, o_done, o_obs = ob.rollout(padded, n_threads=4)
    for gi, (g, r) in enumerate(zip(bm.groups, outs["auto"])):
        rew, rows = r["reward"].cpu().numpy(), r["obs_ring"][(n_steps - 1) % 2].cpu().numpy()
        for j, e in enumerate(sample):
            if bm.env_group[e] == gi:
                slot = bm.env_slot[e]
                np.testing.assert_array_equal(rew[:, slot], o_rew[:, j], err_msg=f"grid {e}")
                np.testing.assert_array_equal(rows[slot], pv_first(o_obs[j, :g.obs_dim], plist[j]), err_msg=f"grid {e}")


def test_trajectory_windows_and_masked_reset():
    """Per-env episode windows (reference: tests/envs/test_trajectory.py:134-156: episode length == final - initial)."""
    p = load_pymgrid25(0)
    B = 300
    bm = engine([p], np.zeros(B, dtype=np.int64))
    rng = np.random.default_rng(2)
    initial = rng.integers(0, 8000, B).astype(np.int32)
    length = rng.integers(2, 12, B).astype(np.int32)
    bm.set_trajectories(initial, initial + length)
    bm.reset()
    assert torch.equal(bm.groups[0].step.cpu(), torch.from_numpy(initial))
    steps_to_done = np.zeros(B, dtype=np.int64)
    alive = np.ones(B, dtype=bool)
    for k in range(12):
        _, _, done, _ = bm.step(torch.rand((B, 2), dtype=torch.float64, device="cuda"))
        d = done.cpu().numpy().astype(bool)
        steps_to_done[alive & d] = k + 1
        alive &= ~d
    np.testing.assert_array_equal(steps_to_done, length)   # done fires on the step that runs at t = final_step - 1
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    mask[::2] = 1
    before = bm.groups[0].step.clone()
    bm.reset(mask=mask)
    after = bm.groups[0].step.cpu().numpy()
    np.testing.assert_array_equal(after[::2], initial[::2])
    np.testing.assert_array_equal(after[1::2], before.cpu().numpy()[1::2])


def test_set_trajectories_keeps_the_handle_and_bound_launchers():
    """mg_set_trajectories swaps the per-env window arrays inside the handle: a launcher bound BEFORE the windows change (what
    HostIO / HostRollout / prepare_step keep) stays valid and honours the new windows, options survive, and the launch
    counter keeps counting (the handle used to be destroyed and re-created here: a use-after-free for bound launchers)."""
    p = load_pymgrid25(1)
    B = 200
    bm = engine([p], np.zeros(B, dtype=np.int64))
    bm.set_image_shape(1)
    acts = torch.rand((B, 4), dtype=torch.float64, device="cuda")
    launch = bm.prepare_step(acts)                       # bound to the handle as it is now
    hio = bm.host_io()
    handle_before, launches_before = bm._handle.value, bm.launch_count
    launch()
    rng = np.random.default_rng(3)
    initial = rng.integers(100, 8000, B).astype(np.int32)
    length = rng.integers(2, 6, B).astype(np.int32)
    bm.set_trajectories(initial, initial + length)
    assert bm._handle.value == handle_before and bm.launch_count == launches_before + 1
    assert bm._options[4] == 1                            # MG_OPT_IMAGE_SHAPE kept
    bm.reset()
    assert torch.equal(bm.groups[0].step.cpu(), torch.from_numpy(initial))
    steps_to_done, alive = np.zeros(B, dtype=np.int64), np.ones(B, dtype=bool)
    for k in range(6):
        launch()                                          # the OLD launcher
        d = bm.groups[0].done.cpu().numpy().astype(bool)
        steps_to_done[alive & d] = k + 1
        alive &= ~d
    np.testing.assert_array_equal(steps_to_done, length)
    hio.actions[0].uniform_(0, 1)
    hio.step()
    hio.sync()
    assert bm.last_kernel == "mg_step_img_kernel"         # windows installed: the ragged hint switched the emitters
    # a second change of windows, then back to the configs' own window
    bm.set_trajectories(np.zeros(B, dtype=np.int32), np.full(B, 3, dtype=np.int32))
    bm.reset()
    for k in range(3):
        launch()
    assert bool(bm.groups[0].done.all()) and int(bm.groups[0].step.max()) == 3


def test_bad_discrete_action_is_flagged():
    p = load_pymgrid25(0)
    bm = engine([p], np.zeros(4, dtype=np.int64))
    a = torch.tensor([0, 1, 2, -1], dtype=torch.int32, device="cuda")
    _, reward, _, _ = bm.step_discrete(a)
    f = bm.groups[0].flags.cpu().numpy()
    assert (f[:2] & (1 << 6) == 0).all() and (f[2:] & (1 << 6) != 0).all()
    assert torch.isnan(reward[2:]).all() and not torch.isnan(reward[:2]).any()
    assert bm.groups[0].step.cpu().tolist() == [1, 1, 0, 0]


@pytest.mark.parametrize("mode,graph", [(1, False), (2, False), (1, True), (2, True)])
def test_overlapped_launches_equal_serial(mode, graph):
    """MG_OPT_STEP_OVERLAP: consecutive step launches chained with programmatic dependent launch -- the next launch's
    physics starts while the previous one is still streaming rows (mode 1), and the row streams overlap too when the
    observation buffers rotate (mode 2) -- issued back to back with nothing in between, directly and replayed from a CUDA
    graph.  Same state, rewards and rows as fully serialised launches."""
    configs = [load_pymgrid25(n) for n in range(25)]
    B, n_steps, R = 65536, 24, 4
    env_config = np.arange(B) % 25
    a = engine(configs, env_config, with_info=False)
    b = engine(configs, env_config, with_info=False)
    a.set_step_overlap(mode)
    b.set_step_overlap(0)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    acts = [torch.rand((n_steps, g.n_envs, g.n_act), dtype=torch.float64, device="cuda", generator=gen) for g in a.groups]
    rings = [torch.empty((R, g.n_envs, g.obs_dim), dtype=torch.float64, device="cuda") for g in a.groups]
    launchers = [a.prepare_step([x[k] for x in acts], obs=[ring[k % R] for ring in rings]) for k in range(n_steps)]
    state0 = a.state_dict()
    if graph:
        for f in launchers[:2]:
            f()
        a.load_state_dict(state0)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            for f in launchers:
                f()
        a.load_state_dict(state0)
        g_.replay()
    else:
        for f in launchers:       # back to back: the chain is never broken
            f()
    torch.cuda.synchronize()
    assert a.last_kernel == "mg_step_kernel"
    for k in range(n_steps):       # one fixed buffer, a clone between the steps -> fully serialised launches
        obs_b, r, _, _ = as_lists(b.step([x[k] for x in acts]))
        if k >= n_steps - R:
            for gi in range(len(a.groups)):
                assert torch.equal(rings[gi][k % R], obs_b[gi]), (k, gi)
    for ga, gb in zip(a.groups, b.groups):
        assert torch.equal(ga.reward, gb.reward) and torch.equal(ga.done, gb.done)
        assert torch.equal(ga.step, gb.step) and torch.equal(ga.charge, gb.charge)
        if ga.genset is not None:
            assert torch.equal(ga.genset, gb.genset)


def test_generator_grids_scaled_series_and_weak_grid_bits(golden):
    """BASELINE config 5 machinery: real MicrogridGenerator grids in (profile, scale) form with per-env grid-status bits:
    engine == the reference's recorded outputs, bit for bit (PV-first observation order)."""
    from tests.helpers import generator_params
    z = golden["generator"]
    n = int(z["n"])
    configs = [generator_params(z, i) for i in range(n)]
    bm = engine(configs, np.arange(n))
    assert bm.obs_order == "gym_sorted_pv_first"
    for k in range(60):
        obs, reward, done, info = as_lists(bm.step(group_actions(bm, [z[f"g{i}_a"][k] for i in range(n)])))
        o, r, d = gather(bm, obs), gather(bm, reward), gather(bm, done)
        for i in range(n):
            assert r[i] == z[f"g{i}_r"][k] and bool(d[i]) == bool(z[f"g{i}_d"][k]), (i, k)
            np.testing.assert_array_equal(o[i], z[f"g{i}_o"][k], err_msg=f"grid {i} step {k}")
    st = state_rows(bm)
    for i in range(n):
        np.testing.assert_array_equal(st[i], z[f"g{i}_s"][-1])


def test_generator_grids_rollout_keeps_sliding_windows(golden):
    """The persistent kernel for per-env series keeps each env's normalised load / pv windows in shared memory and appends
    one value per step (MG_OPT_ROLLOUT_RING): every step's observation row of the real MicrogridGenerator grids equals
    the reference's recorded row, bit for bit -- also when the rollout is cut into launches (the windows are rebuilt at the
    start of each) -- and equals what the window-per-row kernel writes."""
    from tests.helpers import generator_params
    z = golden["generator"]
    n = int(z["n"])
    configs = [generator_params(z, i) for i in range(n)]
    env_config = np.arange(3 * n) % n            # 72 envs: more than one tile's worth of rows per warp pass
    outs = {}
    for ring_kernel in (True, False):
        bm = engine(configs, env_config)
        bm.set_rollout_ring(ring_kernel)
        acts = [torch.from_numpy(np.ascontiguousarray(np.stack([z[f"g{env_config[e]}_a"][:60] for e in g.env_ids], axis=1))).cuda()
                for g in bm.groups]
        first = bm.rollout([a[:23].contiguous() for a in acts], ring=23)
        rest = bm.rollout([a[23:].contiguous() for a in acts], ring=37)
        first, rest = ([x] if isinstance(x, dict) else x for x in (first, rest))
        outs[ring_kernel] = (first, rest)
        for g, f, r in zip(bm.groups, first, rest):
            rows = torch.cat([f["obs_ring"], r["obs_ring"]]).cpu().numpy()          # [60, n_g, D]
            rew = torch.cat([f["reward"], r["reward"]]).cpu().numpy()
            for slot, e in enumerate(g.env_ids):
                i = env_config[e]
                np.testing.assert_array_equal(rew[:, slot], z[f"g{i}_r"][:60])
                np.testing.assert_array_equal(rows[:, slot], z[f"g{i}_o"][:60], err_msg=f"ring={ring_kernel} grid {i}")
    for a, b in zip(outs[True], outs[False]):
        for x, y in zip(a, b):
            assert torch.equal(x["obs_ring"], y["obs_ring"]) and torch.equal(x["done"], y["done"])


@pytest.mark.parametrize("ring_kernel", [True, False])
def test_vectorised_generator_batch_vs_oracle(ring_kernel):
    """5 000 heterogeneous grids built in array form (one config record per env) == the oracle on each grid's explicit
    form, over a rollout that crosses the end of the series for envs started late (both persistent kernels for per-env
    series: sliding windows in shared memory, and whole windows per row)."""
    from pymgrid_b200 import generator
    gb = generator.sample(5000, seed=11)
    bm = generator.engine_from_batch(gb, device="cuda:0", action_order=CONTAINER)
    bm.set_rollout_ring(ring_kernel)
    rng = np.random.default_rng(8)
    plist = [gb.to_params(i) for i in range(gb.n)]
    starts = rng.integers(0, 8700, gb.n)
    starts[::5] = rng.integers(8735, 8745, len(starts[::5]))
    for p, s in zip(plist, starts):
        p.current_step = int(s)
    for g in bm.groups:
        g.step.copy_(torch.from_numpy(starts[g.env_ids].astype(np.int32)))
    ob = OracleBatch(plist)
    n_steps = 14
    padded = np.zeros((n_steps, gb.n, 4))
    for e, p in enumerate(plist):
        padded[:, e, :p.n_act] = rng.random((n_steps, p.n_act))
    o_rew, o_done, o_obs = ob.rollout(padded, n_threads=4)
    acts = [torch.from_numpy(np.ascontiguousarray(padded[:, g.env_ids, :g.n_act])).cuda() for g in bm.groups]
    out = bm.rollout(acts, ring=1)
    out = [out] if isinstance(out, dict) else out
    from tests.test_generator import pv_first
    for g, r in zip(bm.groups, out):
        np.testing.assert_array_equal(r["reward"].cpu().numpy(), o_rew[:, g.env_ids])
        np.testing.assert_array_equal(r["done"].cpu().numpy(), o_done[:, g.env_ids])
        got = r["obs_ring"][0].cpu().numpy()
        for slot, e in enumerate(g.env_ids):
            np.testing.assert_array_equal(got[slot], pv_first(o_obs[e, :g.obs_dim], plist[e]), err_msg=f"env {e}")
    t, charge, gen = ob.state()
    st = state_rows(bm)
    np.testing.assert_array_equal(np.array([s[1] for s in st]), charge)


def test_aggregate_reward_shuffle_reduction():
    """Optional logging aggregate: the kernel's warp-shuffle + atomicAdd total == sum of the per-env rewards
    (floating-point summation order differs: tolerance 1e-12 relative)."""
    configs = [load_pymgrid25(n) for n in range(25)]
    B = 10007
    bm = engine(configs, np.arange(B) % 25, with_info=False)
    total = torch.zeros(1, dtype=torch.float64, device="cuda")
    acts = [torch.rand((g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in bm.groups]
    _, reward, _, _ = as_lists(bm.step(acts, reward_total=total))
    want = sum(float(r.sum()) for r in reward)
    assert abs(total.item() - want) <= 1e-12 * abs(want)
    n_steps = 9
    per_step = torch.zeros(n_steps, dtype=torch.float64, device="cuda")
    racts = [torch.rand((n_steps, g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in bm.groups]
    out = bm.rollout(racts, keep_obs=False, reward_total=per_step)
    want = sum(r["reward"].sum(dim=1) for r in out)
    assert torch.allclose(per_step, want, rtol=1e-12, atol=0)


def test_float32_observation_output_is_the_rounded_f64_observation():
    """Secondary mode (MG_LAYOUT_OBS_F32): obs rows written as float32 == the bit-exact f64 observation rounded to
    nearest; rewards / state are untouched.  Step kernel, persistent kernel, ragged steps, generator grids."""
    from pymgrid_b200 import generator
    rng = np.random.default_rng(3)
    configs = [load_pymgrid25(n) for n in range(25)]
    B = 3001
    env_config = np.arange(B) % 25
    a = engine(configs, env_config, with_info=False)
    b = engine(configs, env_config, with_info=False, obs_dtype=torch.float32)
    randomise_state(a, rng, configs, env_config, 8700)
    b.load_state_dict(a.state_dict())
    for k in range(5):
        acts = [torch.rand((g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in a.groups]
        oa, ra, _, _ = as_lists(a.step(acts))
        ob, rb, _, _ = as_lists(b.step(acts))
        for x, y, r1, r2 in zip(oa, ob, ra, rb):
            assert y.dtype == torch.float32 and torch.equal(x.to(torch.float32), y) and torch.equal(r1, r2)
    racts = [torch.rand((7, g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in a.groups]
    out_a, out_b = a.rollout(racts, ring=2), b.rollout(racts, ring=2)
    for x, y in zip(out_a, out_b):
        assert torch.equal(x["obs_ring"].to(torch.float32), y["obs_ring"]) and torch.equal(x["reward"], y["reward"])
    gb = generator.sample(700, seed=5)
    ga = generator.engine_from_batch(gb, device="cuda:0")
    gf = generator.engine_from_batch(gb, device="cuda:0", obs_dtype=torch.float32)
    for k in range(4):
        acts = [torch.rand((g.n_envs, g.n_act), dtype=torch.float64, device="cuda") for g in ga.groups]
        oa, _, _, _ = as_lists(ga.step(acts))
        ob, _, _, _ = as_lists(gf.step(acts))
        for x, y in zip(oa, ob):
            assert torch.equal(x.to(torch.float32), y)


def test_reward_shapers_golden_and_oracle(golden):
    """Microgrid(reward_shaping_func=...): the shaped reward replaces the step reward (utils/step.py:38-46).
    (1) the reference's recorded BatteryDischargeShaper runs under priority-list dispatch, bit for bit;
    (2) a mixed batch (no shaper / PVCurtailmentShaper / BatteryDischargeShaper side by side in one launch) against the
        oracle, including the flag that stands for the shaper's assert."""
    from oracle.oracle import OracleGrid
    z = golden["shaped"]
    scen = (0, 1, 2, 13)
    configs = []
    for n in scen:
        p = jump_to(load_pymgrid25(n), int(z[f"s{n}_batd_t0"]))
        p.reward_shaper = "battery_discharge"
        configs.append(p)
    bm = engine(configs, np.arange(len(scen)))
    for k in range(150):
        acts = group_actions(bm, [z[f"s{n}_batd_controls"][k] for n in scen])
        _, reward, _, info = as_lists(bm.step(acts, normalized=False))
        r, inf = gather(bm, reward), gather(bm, info)
        for e, n in enumerate(scen):
            assert r[e] == z[f"s{n}_batd_r"][k], (n, k)
            unshaped = 0.0
            for col in (12, 13, 14, 15):
                unshaped += inf[e][col]
            assert unshaped == z[f"s{n}_batd_log_reward"][k]
    assert all(((g.flags & (1 << 7)) == 0).all() for g in bm.groups)

    rng = np.random.default_rng(77)
    configs = []
    for n in scen:
        for shaper in (None, "pv_curtailment", "battery_discharge"):
            p = jump_to(load_pymgrid25(n), int(rng.integers(0, 8000)))
            p.reward_shaper = shaper
            configs.append(p)
    B, n_steps = 96, 25
    env_config = np.arange(B) % len(configs)
    bm = engine(configs, env_config)
    oracles = [OracleGrid(configs[c]) for c in env_config]
    n_flagged = 0
    for k in range(n_steps):
        a = [rng.random(o.n_act) for o in oracles]
        _, reward, _, _ = as_lists(bm.step(group_actions(bm, a)))
        r = gather(bm, reward)
        flags = gather(bm, [g.flags for g in bm.groups])
        for e, o in enumerate(oracles):
            _, ro, _, _, err = o.run(a[e])
            assert (r[e] == ro) or (np.isnan(r[e]) and np.isnan(ro)), (e, k)
            assert (int(flags[e]) & 0xffff) == (err & 0xffff), (e, k)
            n_flagged += bool(err & (1 << 7))
    assert n_flagged > 20


@pytest.mark.parametrize("emit", ["lsu", "image"])
def test_float32_actions_equal_their_float64_widening(emit):
    """set_action_dtype(torch.float32): a policy's float32 actions are widened exactly on the device, so steps, rollouts,
    HostIO and HostRollout give bit for bit what the float64 path gives for `actions.double()` -- all three action widths
    (2, 3 and 4 columns: vector and scalar loads), normalised and not, partial tiles."""
    rng = np.random.default_rng(31)
    configs = [load_pymgrid25(n) for n in range(25)]
    B, T = 1733, 9
    env_config = rng.integers(0, 25, B)
    ref, bm = engine(configs, env_config), engine(configs, env_config)
    for e in (ref, bm):
        e.set_emit_image(emit == "image")
    bm.set_action_dtype(torch.float32)
    assert bm.action_dtype == torch.float32 and ref.action_dtype == torch.float64
    a32 = [torch.from_numpy(rng.random((T, g.n_envs, g.n_act)).astype(np.float32)).cuda() for g in bm.groups]
    a64 = [a.double() for a in a32]
    assert sorted({g.n_act for g in bm.groups}) == [2, 3, 4]
    # persistent rollout
    want, got = ref.rollout([a[:5].contiguous() for a in a64], ring=2), bm.rollout([a[:5].contiguous() for a in a32], ring=2)
    for w, g in zip(want, got):
        for k in ("reward", "done", "obs_ring"):
            assert torch.equal(w[k], g[k]), k
    # single steps, normalised and unnormalised
    for k, normalized in ((5, True), (6, False)):
        w = as_lists(ref.step([a[k].contiguous() for a in a64], normalized=normalized))
        g = as_lists(bm.step([a[k].contiguous() for a in a32], normalized=normalized))
        for wo, go, wr, gr in zip(w[0], g[0], w[1], g[1]):
            assert torch.equal(wo, go) and torch.equal(wr, gr)
    # host-resident actions: one step, then a pipelined rollout in chunks of two
    hio_w, hio_g = ref.host_io(), bm.host_io()
    for hw, hg, aw, ag in zip(hio_w.actions, hio_g.actions, a64, a32):
        hw.copy_(aw[7].cpu())
        hg.copy_(ag[7].cpu())
    hio_w.step(); hio_g.step()
    hio_w.sync(); hio_g.sync()
    assert torch.equal(hio_w.reward, hio_g.reward) and torch.equal(hio_w.done, hio_g.done)
    hr_w, hr_g = ref.host_rollout(2, chunk=1), bm.host_rollout(2, chunk=1)
    assert hr_g.h2d_bytes_per_step * 2 == hr_w.h2d_bytes_per_step
    for hw, hg, aw, ag in zip(hr_w.actions, hr_g.actions, a64, a32):
        hw.copy_(aw[7:9].cpu())
        hg.copy_(ag[7:9].cpu())
    hr_w.run(); hr_g.run()
    hr_w.sync(); hr_g.sync()
    for rw, rg, dw, dg in zip(hr_w.reward, hr_g.reward, hr_w.done, hr_g.done):
        assert torch.equal(rw, rg) and torch.equal(dw, dg)
    for gw, gg in zip(ref.groups, bm.groups):
        assert torch.equal(gw.charge, gg.charge) and torch.equal(gw.step, gg.step)
        assert gw.genset is None or torch.equal(gw.genset, gg.genset)
    # the wrong element type is refused
    with pytest.raises(ValueError):
        bm.step([a[0].contiguous() for a in a64])
    bm.set_action_dtype(torch.float64)
    bm.step([a[0].contiguous() for a in a64])
