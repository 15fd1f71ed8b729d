"""Episode windows, vectorised over a batch (reference: src/pymgrid/microgrid/trajectory/{deterministic,stochastic}.py).

The reference calls `trajectory_func(initial_step, final_step)` once per `reset()` of one microgrid
(microgrid.py:221-225); here the same three rules draw one `(initial, final)` pair per env and the pairs go to the
engine with `BatchedMicrogrid.set_trajectories` (the kernel's `done` and `reset` honour them per env).
"""
import numpy as np


class DeterministicTrajectory:
    """deterministic.py:4-12: every episode is [initial_step, final_step)."""

    def __init__(self, initial_step, final_step):
        self.initial_step, self.final_step = initial_step, final_step

    def __call__(self, initial_step, final_step, n=None, rng=None):
        if n is None:
            return self.initial_step, self.final_step
        return np.full(n, self.initial_step, dtype=np.int32), np.full(n, self.final_step, dtype=np.int32)


class StochasticTrajectory:
    """stochastic.py:6-13: initial ~ U{initial_step .. final_step-3}, final ~ U{initial .. final_step-1}."""

    def __call__(self, initial_step, final_step, n=None, rng=None):
        rng = np.random.default_rng() if rng is None else rng
        size = 1 if n is None else n
        initial = rng.integers(initial_step, final_step - 2, size)
        final = rng.integers(initial, final_step, size)          # note: the reference allows final == initial
        final = np.maximum(final, initial + 1)                    # an empty window is rejected by Microgrid (microgrid.py:199-201)
        if n is None:
            return int(initial[0]), int(final[0])
        return initial.astype(np.int32), final.astype(np.int32)


class FixedLengthStochasticTrajectory:
    """stochastic.py:16-30: a window of `trajectory_length` steps starting uniformly inside [initial_step, final_step)."""

    def __init__(self, trajectory_length):
        self.trajectory_length = trajectory_length

    def __call__(self, initial_step, final_step, n=None, rng=None):
        if final_step - initial_step < self.trajectory_length:
            raise ValueError(f'Cannot create a trajectory of length {self.trajectory_length}'
                             f'between initial_step ({initial_step}) and final_step ({final_step})')
        rng = np.random.default_rng() if rng is None else rng
        size = 1 if n is None else n
        initial = rng.integers(initial_step, max(final_step - self.trajectory_length, initial_step + 1), size)
        final = initial + self.trajectory_length
        if n is None:
            return int(initial[0]), int(final[0])
        return initial.astype(np.int32), final.astype(np.int32)


def apply(bm, trajectory, rng=None):
    """Draw one window per env of `bm` and install them (then call `bm.reset()` to start the episodes)."""
    lo = min(p.initial_step for p in bm.configs) if bm.configs else 0
    hi = min(p.final_step for p in bm.configs) if bm.configs else bm.series_len
    initial, final = trajectory(lo, hi, n=bm.n_envs, rng=rng)
    bm.set_trajectories(initial, final)
    return initial, final
