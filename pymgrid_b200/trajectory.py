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


def validated(func, first, last, integer_types=(int,)):
    """Accept or refuse a `trajectory_func` the way the reference's constructors do (Microgrid._check_trajectory_func,
    microgrid/microgrid.py:167-199): None passes, anything else is called ONCE on the microgrid's own [first, last) window and
    must answer with two integers that lie inside it and span at least one step.  Same exception types and messages as the
    reference, so callers' error handling carries over.  Returns `func`."""
    if func is None:
        return None
    if not callable(func):
        raise TypeError('trajectory_func must be callable.')
    answer = func(first, last)
    pair = None
    try:
        a, b = answer
        if isinstance(a, integer_types) and isinstance(b, integer_types):
            pair = (a, b)
    except (TypeError, ValueError):
        pass
    if pair is None:
        raise TypeError(f'trajectory func must return two integer values, not {answer}')
    start, stop = pair
    problems = ((start < first, f'trajectory_func returned initial_step value ({start}) less than env\'s initial step: ({first})'),
                (stop > last, f'trajectory_func returned final_step value ({stop}) greater than env\'s final step: ({last})'),
                (start >= stop, f'trajectory_func returned values ({start}, {stop}) such that initial_step'
                                f'was greater than or equal to final_step.'))
    for bad, message in problems:
        if bad:
            raise ValueError(message)
    return func


def takes_batch_size(func):
    """True when `func` can draw many windows in one call (`func(first, last, n=...)`, like the classes above); a plain
    reference-style callable is called once per env instead.  Decided from the signature, so that a TypeError raised INSIDE
    a user's function is never mistaken for a missing parameter."""
    import inspect
    try:
        params = inspect.signature(func).parameters
    except (TypeError, ValueError):
        return False
    return "n" in params or any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())


def draw(func, first, last, n):
    """n windows from `func` as two int32 arrays"""
    if takes_batch_size(func):
        initial, final = func(first, last, n=n)
    else:
        pairs = [func(first, last) for _ in range(n)]
        initial, final = [p[0] for p in pairs], [p[1] for p in pairs]
    return np.asarray(initial, dtype=np.int32).reshape(n).copy(), np.asarray(final, dtype=np.int32).reshape(n).copy()


def apply(bm, trajectory, rng=None):
    """Draw one window per env of `bm` and install them (then call `bm.reset()` to start the episodes).  The windows lie
    inside EVERY config's own window: from the latest initial_step to the earliest final_step."""
    lo = max(p.initial_step for p in bm.configs) if bm.configs else 0
    hi = min(p.final_step for p in bm.configs) if bm.configs else bm.series_len
    initial, final = trajectory(lo, hi, n=bm.n_envs, rng=rng)
    bm.set_trajectories(initial, final)
    return initial, final
