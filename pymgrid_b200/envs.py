"""Gym-style environments on the batched engine -- the reference's `pymgrid.envs` surface, for B envs at once.

reference: src/pymgrid/envs/base/base.py (BaseMicrogridEnv: `step`, `reset`, flat `observation_space`),
src/pymgrid/envs/discrete/discrete.py (DiscreteMicrogridEnv: action = priority list), and
src/pymgrid/envs/continuous/continuous.py (ContinuousMicrogridEnv; broken at the reference commit -- the intended
semantics are implemented: flat Box action in [0,1]^n_act = the controllable modules' normalised actions,
`Microgrid.run(normalized=True)`, flattened normalised observation).

`step(actions)` takes / returns torch DEVICE tensors with a leading batch dimension and follows the old 4-tuple Gym API
like the reference: `(obs [B, D], reward [B], done [B], info)`.  With `batch=None` the env is a single microgrid and
`step` takes / returns the reference's host types (int / np.ndarray action, np.ndarray observation, float, bool, dict).
gym itself is not a dependency: `Box` / `Discrete` below carry the attributes callers read (`shape`, `low`, `high`, `n`,
`sample`, `contains`).
"""
import numpy as np
import torch

from . import views
from .engine import BatchedMicrogrid
from .scenario import load_pymgrid25


class Box:
    def __init__(self, low, high, shape, dtype=np.float64):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    __contains__ = contains

    def __repr__(self):
        return f"Discrete({self.n})"


def _composed_route(cls, configs, kw):
    """module lists outside the fused kernels' scope go to the composed path (compose.py), same surface"""
    if isinstance(configs, (list, tuple)) and configs and not any(hasattr(c, "arch") for c in configs):
        from .compose import ComposedContinuousEnv, ComposedDiscreteEnv, in_fused_scope
        if not in_fused_scope(configs, kw.get("add_unbalanced_module", True)):
            return ComposedDiscreteEnv if cls is DiscreteMicrogridEnv else ComposedContinuousEnv
    return None


class _BaseEnv:
    """Shared plumbing: one architecture (all envs share obs / action layout), B replicas or B heterogeneous configs."""

    def __new__(cls, configs=None, env_config=None, batch=None, **kw):
        target = _composed_route(cls, configs, kw) if cls in (DiscreteMicrogridEnv, ContinuousMicrogridEnv) and env_config is None else None
        if target is not None:
            return target(configs, batch=batch, **kw)
        return super().__new__(cls)

    def __init__(self, configs, env_config=None, batch=None, device=None, obs_order="gym_sorted", with_info=False,
                 add_unbalanced_module=True, loss_load_cost=10., overgeneration_cost=2., reward_shaping_func=None,
                 trajectory_func=None, flat_spaces=True, observation_keys=(), remove_redundant_gensets=True):
        """`configs`: one MicrogridParams, a list of them, or -- the reference's call, BaseMicrogridEnv(modules,
        add_unbalanced_module, loss_load_cost, overgeneration_cost, reward_shaping_func, trajectory_func)
        (envs/base/base.py:84-110) -- a list of `pymgrid_b200.modules` objects / (name, module) tuples."""
        from .params import MicrogridParams
        if isinstance(configs, MicrogridParams):
            configs = [configs]
        else:
            configs = list(configs)
            if not all(isinstance(c, MicrogridParams) for c in configs):
                from .modules import params_from_modules
                configs = [params_from_modules(configs, add_unbalanced_module, loss_load_cost, overgeneration_cost)]
        if reward_shaping_func is not None:
            import dataclasses
            name = reward_shaping_func if isinstance(reward_shaping_func, str) else type(reward_shaping_func).__name__
            configs = [dataclasses.replace(c, reward_shaper=name) for c in configs]
        self.single = batch is None and env_config is None
        if env_config is None:
            env_config = np.arange(1 if batch is None else batch) % len(configs)
        if len({c.arch for c in configs}) != 1:
            raise ValueError("an env batch holds one architecture (use BatchedMicrogrid directly for mixed batches)")
        self.engine = BatchedMicrogrid(configs, env_config, device=device, obs_order=obs_order,
                                       with_info=with_info or self.single, with_flags=True,
                                       remove_redundant_gensets=remove_redundant_gensets,
                                       action_order=None if obs_order == "gym_sorted" else views.CONTROL_ORDER)
        self.group = self.engine.groups[0]
        self.params = configs[0]
        self.n_envs = self.engine.n_envs
        self._obs_order = obs_order
        if not flat_spaces:
            raise NotImplementedError("flat_spaces=False (nested gym spaces) is not part of the batched surface")
        # observation_keys (base.py:109-163, 211-218): the observation is state_series(normalized=True).loc[:, :, keys] -- for
        # every key in the order given, the modules that have such a field in listing order.  The fused kernels write full
        # rows; the env gathers the selected columns (composed batches write only the selected elements, compose.py)
        observation_keys = observation_keys or ()      # (None is the reference DiscreteMicrogridEnv's own default)
        self.observation_keys = [observation_keys] if isinstance(observation_keys, str) else list(observation_keys)
        self._take = self._take_dev = None
        if self.observation_keys:
            sl = views.obs_slices(self.params, obs_order)
            sd = views.state_dict(self.params, self.params.current_step, 0.0, (0, 0, 0, 0))
            fields = {name: list(sd[name].keys()) for name in sl}
            bad = [k for k in self.observation_keys if not any(k in f for f in fields.values())]
            if bad:
                raise NameError(f'Keys {bad} not found in state.')
            listing = [n for n in ("load", "pv", "genset", "battery", "grid") if n in sl]
            self._take = np.array([sl[n].start + fields[n].index(k) for k in self.observation_keys for n in listing if k in fields[n]],
                                  dtype=np.int64)
            self._take_dev = torch.from_numpy(self._take).to(self.engine.device)
        self.observation_space = Box(0.0, 1.0, (len(self._take) if self._take is not None else self.group.obs_dim,))     # base.py:161-163
        self._log_rows = []         # single microgrid: the per-step log the reference's env keeps (it IS a Microgrid)
        self.trajectory_func = self._check_trajectory_func(trajectory_func)

    def _check_trajectory_func(self, trajectory_func):
        """reference: Microgrid._check_trajectory_func (microgrid.py:167-199): see trajectory.validated"""
        from .trajectory import validated
        return validated(trajectory_func, self.params.initial_step, self.params.final_step, integer_types=(int, np.integer))

    @classmethod
    def from_scenario(cls, microgrid_number=0, batch=None, **kw):
        """reference: BaseMicrogridEnv.from_scenario (envs/base/base.py:292-299)."""
        return cls(load_pymgrid25(microgrid_number), batch=batch, **kw)

    @classmethod
    def from_microgrid(cls, microgrid, batch=None, **kw):
        """reference: BaseMicrogridEnv.from_microgrid (envs/base/base.py:270-290): an env over a copy of a (possibly
        running) microgrid, state included.  Accepts a pymgrid_b200.Microgrid or a MicrogridParams."""
        from .compose import ComposedContinuousEnv, ComposedDiscreteEnv, ComposedMicrogrid
        if isinstance(microgrid, ComposedMicrogrid):       # any module list: the composed path
            target = ComposedDiscreteEnv if issubclass(cls, DiscreteMicrogridEnv) else ComposedContinuousEnv
            return target.from_microgrid(microgrid, batch=batch, **kw)
        params = microgrid.export_params() if hasattr(microgrid, "export_params") else microgrid
        return cls(params, batch=batch, **kw)

    @property
    def modules(self):
        """module views of the (first) microgrid, as notebooks read them (`env.modules`)"""
        from .microgrid import ModuleContainerView, ModuleView
        names = ["load", "pv", "unbalanced_energy"] + (["genset"] if self.params.has_genset else []) + ["battery"] + \
                (["grid"] if self.params.has_grid else [])
        nm = lambda n: {"pv": self.params.renewable_name, "unbalanced_energy": self.params.unbalanced_name}.get(n, n)   # noqa: E731
        return ModuleContainerView((nm(n), [ModuleView(self, n, nm(n))]) for n in names)

    def _state(self):     # live state of env 0, for the module views
        g = self.group
        gen = tuple(int(x) for x in self.engine.genset_status(0)[0].tolist()) if g.genset is not None else (0, 0, 0, 0)
        charge, b = float(g.charge[0].item()), self.params.battery
        soc = b.soc if (self.engine._soc_pristine and b.soc is not None) else charge / b.max_capacity   # battery_module.py:89, 130
        return dict(t=int(g.step[0].item()), charge=charge, genset=gen, soc=soc)

    @property
    def initial_step(self):
        return self.params.initial_step

    @property
    def final_step(self):
        return self.params.final_step

    @property
    def current_step(self):
        return self.group.step if not self.single else int(self.group.step[0].item())

    def reset(self, mask=None):
        """reference: BaseMicrogridEnv.reset (base.py:165-167): flat observation after Microgrid.reset."""
        if self.trajectory_func is not None:      # microgrid.py:221-225: a new episode window per reset
            self._draw_windows(mask)
        obs = self.engine.reset(mask=mask)
        self._log_rows = []
        return self._select(obs[0].cpu().numpy() if self.single else obs)

    # ---- the log of a single microgrid (reference: the env inherits Microgrid.get_log / .log, microgrid.py:434-475) ----
    def _log_step(self, pre, info_row, reward, action=None):
        p = self.params
        row = views.log_row(p, views.state_dict(p, pre["t"], pre["charge"], pre["genset"], pre["soc"]), info_row, reward,
                            self._state()["genset"])
        row = views.caller_names(row, p)
        if action is not None:
            row[("action", 0, "")] = action      # DiscreteMicrogridEnv logs the action it was given (discrete.py:141)
        self._log_rows.append(row)

    def get_log(self, as_frame=True, drop_singleton_key=False):
        if not self.single:
            raise NotImplementedError("batched envs keep no per-step log (808 GB per year at 65 536 envs); use "
                                      "BatchedMicrogrid.recorder(env_ids) for selected envs")
        df = views.log_frame(self._log_rows, int(self.group.step[0].item()), drop_singleton_key)
        return df if as_frame else df.to_dict()

    @property
    def log(self):
        return self.get_log()

    def _select(self, obs):
        if self._take is None:
            return obs
        return obs[self._take] if self.single else obs.index_select(1, self._take_dev)

    def _draw_windows(self, mask):
        """One (initial_step, final_step) pair per env that is being reset, from `trajectory_func(initial, final)`; the
        vectorised trajectory classes of pymgrid_b200.trajectory draw all of them in one call (`n=`)."""
        from .trajectory import draw
        lo, hi, n = self.params.initial_step, self.params.final_step, self.n_envs
        initial, final = draw(self.trajectory_func, lo, hi, n)
        if (initial < lo).any() or (final > hi).any() or (initial >= final).any():
            raise ValueError(f"trajectory_func returned a window outside [{lo}, {hi}] or an empty one")   # microgrid.py:184-197
        if mask is not None and getattr(self, "_windows", None) is not None:     # envs that keep running keep their window
            keep = ~np.asarray(mask.cpu() if hasattr(mask, "cpu") else mask, dtype=bool).reshape(n)
            initial[keep], final[keep] = self._windows[0][keep], self._windows[1][keep]
        self._windows = (initial, final)
        self._module_window = (int(initial[0]), int(final[0]))      # what the module views of env 0 report
        self.engine.set_trajectories(initial, final)

    def _finish(self, res, pre=None, action=None):
        obs, reward, done, info = res
        if not self.single:
            return self._select(obs), reward, done, ({} if info is None else {"info_block": info, "flags": self.group.flags})
        flags = int(self.group.flags[0].item()) & 0xffffffff
        # one microgrid: the reference raises where the engine flags (a batch keeps the flags and a NaN reward per env instead);
        # like Microgrid.run, the exception comes after the step has been applied to the state
        from .microgrid import raise_for_flags
        raise_for_flags(flags, self.params, False, step=None if pre is None else pre["t"])
        info_row, r = info[0].cpu().numpy(), float(reward[0].item())
        if pre is not None:
            self._log_step(pre, info_row, r, action)
        return (self._select(obs[0].cpu().numpy()), r, bool(done[0].item()),
                views.caller_names(views.info_row_to_dict(info_row, flags, self.params), self.params))

    def __len__(self):
        return len(self.params)


class DiscreteMicrogridEnv(_BaseEnv):
    """Action = index of a priority list (reference: envs/discrete/discrete.py:60-143)."""

    def __init__(self, configs, env_config=None, batch=None, remove_redundant_gensets=True, **kw):
        # (the flag decides which priority lists exist, i.e. what an action index means: discrete.py:60-80, priority_list.py:15-67)
        super().__init__(configs, env_config, batch, remove_redundant_gensets=remove_redundant_gensets, **kw)
        from .algos import _elements
        # the reference's form (envs/discrete/discrete.py:60-80): one tuple of PriorityListElement per action
        self.actions_list = [tuple(_elements(self.params, pl)) for pl in self.engine.action_tables[0]]
        self.action_space = Discrete(len(self.actions_list))
        self._a = torch.zeros(self.n_envs, dtype=torch.int32, device=self.engine.device)
        self._index = None          # env action -> row of the engine's table, once remove_action() has been used

    def remove_action(self, action_number):
        """reference: DiscreteMicrogridEnv.remove_action (envs/discrete/discrete.py:90-105): drop one priority list from the
        action space; the remaining actions are renumbered"""
        if action_number not in self.action_space:
            raise ValueError('Cannot remove action that is not in the action space!')
        if self._index is None:
            self._index = list(range(len(self.actions_list)))
        self.actions_list.pop(action_number)
        self._index.pop(action_number)
        self.action_space = Discrete(self.action_space.n - 1)
        self._index_dev = torch.tensor(self._index, dtype=torch.int64, device=self.engine.device)

    def step(self, action):
        if self.single:
            if action not in self.action_space:
                raise ValueError(f" Action {action} not in action space {self.action_space}")   # discrete.py:84
            self._a[0] = int(action) if self._index is None else self._index[int(action)]
            pre = self._state()
            return self._finish(self.engine.step_discrete(self._a), pre, int(action))
        if self._index is not None:     # renumbered action space: map to the engine's rows, out-of-space actions stay invalid
            a = torch.as_tensor(action, device=self.engine.device).to(torch.int64)
            bad = (a < 0) | (a >= len(self._index))
            action = torch.where(bad, torch.full_like(a, -1), self._index_dev[a.clamp(0, len(self._index) - 1)]).to(torch.int32)
        return self._finish(self.engine.step_discrete(action))

    def sample_action(self):
        if self.single:
            return self.action_space.sample()
        return torch.randint(0, self.action_space.n, (self.n_envs,), dtype=torch.int32, device=self.engine.device)


class ContinuousMicrogridEnv(_BaseEnv):
    """Action = flat vector in [0,1]^n_act, the controllable modules' normalised actions concatenated in the env's
    module order (`action_layout`); see the module docstring for the relation to the reference's class."""

    def __init__(self, configs, env_config=None, batch=None, **kw):
        super().__init__(configs, env_config, batch, **kw)
        self.action_space = Box(0.0, 1.0, (self.group.n_act,))
        self.action_layout = dict(self.group.act_cols)       # module name -> first column
        self._a = torch.zeros((self.n_envs, self.group.n_act), dtype=torch.float64, device=self.engine.device)

    def step(self, action, normalized=True):
        if self.single:
            self._a.copy_(torch.as_tensor(np.asarray(action, dtype=np.float64)).reshape(1, -1))
            pre = self._state()
            return self._finish(self.engine.step(self._a, normalized=normalized), pre)
        return self._finish(self.engine.step(action, normalized=normalized))

    def sample_action(self):
        if self.single:
            return self.action_space.sample()
        return torch.rand((self.n_envs, self.group.n_act), dtype=torch.float64, device=self.engine.device)
