"""Minimal stand-in for the `gym` package (test infrastructure only).

The upstream reference imports `gym` (un-vendored, un-pinned: requirements.txt:3 `gym>=0.15.7`)
for its space containers.  gym is not installable in the build container (no network), so this
shim restates just the container semantics the reference touches: Box / Dict / Tuple / Discrete,
`flatten_space` and `flatten`.  `Dict` sorts plain-mapping keys like real gym (0.21-0.26) does,
which is what decides the flat observation order.

Nothing in the shipped package imports this.  It exists so that `oracle/ref_loader.py` can import the
unmodified reference from /root/reference/src inside the build container to pin the oracle and to
generate the golden fixtures under tests/golden/.
"""
__version__ = "0.26.2+shim"

from . import spaces  # noqa: E402,F401


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self
