"""CPU stand-in for `pymgrid_b200.engine.BatchedMicrogrid`, backed by the C oracle -- TEST INFRASTRUCTURE ONLY.

The drop-in host layer (`pymgrid_b200.microgrid.Microgrid`, `pymgrid_b200.envs`, `pymgrid_b200.algos`) is plain Python on
top of the engine's tensors: dict / DataFrame conversions, module views, trajectory windows, priority-list bookkeeping.
None of it needs a GPU to be WRONG, so the CPU suite runs it here against the same golden vectors the GPU suite uses
(tests/test_dropin_host.py), with this class monkeypatched in for the engine.  It implements only what that layer calls,
for ONE architecture group, stepping every env through `oracle.OracleGrid`.  It is not importable from the package and
nothing in `pymgrid_b200/` refers to it: the product path still fails loudly without CUDA (tests/test_cabi.py).
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from oracle.oracle import ORDER_CONTAINER, ORDER_GYM_SORTED, OracleGrid
from pymgrid_b200 import priority_list as PL
from pymgrid_b200 import views

CONTAINER = ("genset", "battery", "grid")
FLAG_BAD_ACTION = 1 << 6


class OracleBackedEngine:
    def __init__(self, configs, env_config, device=None, obs_order="gym_sorted", with_info=False, with_flags=False,
                 action_order=None, remove_redundant_gensets=True, **_):
        self.configs, self.env_config = list(configs), np.asarray(env_config, dtype=np.int64)
        if len({c.arch for c in self.configs}) != 1:
            raise NotImplementedError("the CPU stand-in holds one architecture group")
        if any(c.renewable_name < "battery" for c in self.configs) and obs_order == "gym_sorted":
            raise NotImplementedError("the oracle has no 'PV'-first observation order")
        self.device = torch.device("cpu")
        self.obs_order = obs_order
        self._order = ORDER_GYM_SORTED if obs_order == "gym_sorted" else ORDER_CONTAINER
        self.n_envs = len(self.env_config)
        self.grids = [OracleGrid(self.configs[c], order=self._order) for c in self.env_config]
        p = self.configs[0]
        names = [m for m, present in zip(CONTAINER, (p.has_genset, True, p.has_grid)) if present]
        act_names = (sorted(names) if obs_order.startswith("gym_sorted") else names) if action_order is None \
            else [m for m in action_order if m in names]
        cols, col = {}, 0
        for m in act_names:
            cols[m] = col
            col += 2 if m == "genset" else 1
        n = self.n_envs
        self.action_tables = [PL.priority_lists(c.has_genset, c.has_grid,
                                                c.genset.running_min_production if c.has_genset else None,
                                                remove_redundant_gensets) for c in self.configs]
        self._g = SimpleNamespace(
            arch=p.arch, env_ids=np.arange(n), n_act=p.n_act, obs_dim=p.obs_dim, act_cols=cols, n_envs=n,
            step=torch.zeros(n, dtype=torch.int32), charge=torch.zeros(n, dtype=torch.float64),
            genset=torch.zeros(n, dtype=torch.int32) if p.has_genset else None,
            obs=torch.zeros((n, p.obs_dim), dtype=torch.float64), reward=torch.zeros(n, dtype=torch.float64),
            done=torch.zeros(n, dtype=torch.uint8), info=torch.zeros((n, 16), dtype=torch.float64),
            flags=torch.zeros(n, dtype=torch.int32), env_initial_step=None, env_final_step=None,
            n_actions=max(len(t) for t in self.action_tables))
        self.groups = [self._g]
        self.single_group = True
        self._soc_pristine = True
        self.launch_count = 0
        self._pull()

    # ---- state mirrors: the host layer reads `groups[0].step / charge / genset` like the engine's device tensors ---------
    def _pull(self):
        g = self._g
        for e, o in enumerate(self.grids):
            s = o.state
            g.step[e], g.charge[e] = s["t"], s["charge"]
            if g.genset is not None:
                cs, gs, up, dn = s["genset"]
                g.genset[e] = cs | (gs << 8) | (up << 16) | (dn << 24)

    def genset_status(self, gi=0):
        w = self._g.genset
        return torch.stack([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, (w >> 24) & 0xff], dim=1)

    def _container_row(self, row):
        g, out = self._g, []
        for m in CONTAINER:
            if m in g.act_cols:
                out += list(row[g.act_cols[m]:g.act_cols[m] + (2 if m == "genset" else 1)])
        return np.array(out, dtype=np.float64)

    def _run(self, e, ctrl, normalized, extra_flags=0):
        g, o = self._g, self.grids[e]
        obs, r, d, info, flags = o.run(ctrl, normalized=normalized)
        flags |= extra_flags
        if flags & (1 << 5):            # stepping past the end: nothing is written but nan / done / the flag
            g.reward[e], g.done[e], g.flags[e] = float("nan"), 1, int(flags)
            return
        g.obs[e], g.reward[e], g.done[e] = torch.from_numpy(obs), r, int(d)
        g.info[e], g.flags[e] = torch.from_numpy(info), int(np.int32(np.uint32(flags)))

    def _result(self, obs):
        g = self._g
        return (g.obs if obs else None), g.reward, g.done, g.info

    def step(self, actions, normalized=True, obs=True, reward_total=None):
        a = actions[0] if isinstance(actions, (list, tuple)) else actions
        a = a.numpy()
        for e in range(self.n_envs):
            self._run(e, self._container_row(a[e]), normalized)
        self._soc_pristine = False
        self.launch_count += 1
        self._pull()
        return self._result(obs)

    def step_discrete(self, actions, obs=True):
        a = (actions[0] if isinstance(actions, (list, tuple)) else actions).numpy()
        g = self._g
        for e in range(self.n_envs):
            table = self.action_tables[self.env_config[e]]
            if not 0 <= int(a[e]) < len(table):
                g.reward[e], g.done[e], g.flags[e] = float("nan"), 0, FLAG_BAD_ACTION
                continue
            o = self.grids[e]
            ctrl = o.priority_control(table[int(a[e])])
            self._run(e, ctrl, False, extra_flags=o.list_flags)
        self._soc_pristine = False
        self.launch_count += 1
        self._pull()
        return self._result(obs)

    def reset(self, mask=None, obs=True):
        m = None if mask is None else np.asarray(mask, dtype=bool).reshape(-1)
        for e, o in enumerate(self.grids):
            self._g.obs[e] = torch.from_numpy(o.reset() if (m is None or m[e]) else o.observe())
        self._pull()
        return self._g.obs

    def observe(self, obs=True):
        for e, o in enumerate(self.grids):
            self._g.obs[e] = torch.from_numpy(o.observe())
        return self._g.obs

    def set_trajectories(self, initial_step, final_step):
        initial_step, final_step = np.asarray(initial_step, dtype=np.int32), np.asarray(final_step, dtype=np.int32)
        for e, o in enumerate(self.grids):
            o.g.initial_step, o.g.final_step = int(initial_step[e]), int(final_step[e])
        self._g.env_initial_step = torch.from_numpy(initial_step.copy())
        self._g.env_final_step = torch.from_numpy(final_step.copy())

    def rbc_actions(self):
        idx = [self.action_tables[k].index(PL.rbc_priority_list(p)) for k, p in enumerate(self.configs)]
        return [torch.from_numpy(np.asarray(idx, dtype=np.int32)[self.env_config])]


def install(monkeypatch):
    """Route the drop-in host layer to the stand-in for the duration of one test."""
    import pymgrid_b200.envs as envs
    import pymgrid_b200.microgrid as microgrid
    monkeypatch.setattr(microgrid, "BatchedMicrogrid", OracleBackedEngine)
    monkeypatch.setattr(envs, "BatchedMicrogrid", OracleBackedEngine)


__all__ = ["OracleBackedEngine", "install", "C", "views"]
