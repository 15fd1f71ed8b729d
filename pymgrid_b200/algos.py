"""Rule-based control on the engine -- the surface of `pymgrid.algos.RuleBasedControl` (reference: algos/rbc/rbc.py:7-140,
algos/priority_list/priority_list.py:14-67, priority_list_element.py:6-80) for code like `notebooks/rbc-example.ipynb`:

    rbc = RuleBasedControl(microgrid)
    log = rbc.run()                      # DataFrame, same columns as the reference's
    log.loc[:, pd.IndexSlice[:, :, 'reward']].sum()

The reference expands the priority list into controls in Python every step and calls `Microgrid.run`; here the list is
an index into the engine's action table and the expansion (`priority_control()` in csrc/mg_engine.cu) is fused in front of
the step on the device.  `BatchedMicrogrid.rollout_rbc` is the throughput path (one persistent kernel for a whole batch);
this class keeps the per-step log.
"""
from typing import NamedTuple, Optional, Tuple

from . import priority_list as _pl
from .microgrid import Microgrid


class PriorityListElement(NamedTuple):
    """One position of a deployment order (reference: priority_list_element.py:6-80).  Ordered by marginal cost, ties
    towards the higher action number; elements without a cost do not compare."""
    module: Tuple[str, int]
    module_actions: int
    action: int
    marginal_cost: Optional[float] = None

    def _key(self):
        return (self.marginal_cost, -self.action)

    def __lt__(self, other):
        if not isinstance(other, PriorityListElement) or self.marginal_cost is None or other.marginal_cost is None:
            return NotImplemented
        return self._key() < other._key()

    def __gt__(self, other):
        if not isinstance(other, PriorityListElement) or self.marginal_cost is None or other.marginal_cost is None:
            return NotImplemented
        return self._key() > other._key()

    def __le__(self, other):
        return self == other or self < other

    def __ge__(self, other):
        return self == other or self > other


def _elements(params, pl):
    """engine form ((module id, action), ...) -> the reference's PriorityListElement objects"""
    cost = _pl.marginal_costs(params)
    return [PriorityListElement(module=(_pl.MODULE_NAMES[m], 0), module_actions=2 if m == _pl.GENSET else 1, action=a,
                                marginal_cost=cost[m]) for m, a in pl]


def _engine_form(elements):
    ids = {name: k for k, name in _pl.MODULE_NAMES.items()}
    try:
        return tuple((ids[el.module[0]], int(el.action)) for el in elements)
    except (AttributeError, KeyError, TypeError, IndexError):
        return None


class RuleBasedControl:
    def __new__(cls, microgrid=None, *args, **kw):
        from .compose import ComposedMicrogrid, ComposedRuleBasedControl
        if cls is RuleBasedControl and isinstance(microgrid, ComposedMicrogrid):      # any module list: the composed path
            return ComposedRuleBasedControl(microgrid, *args, **kw)
        return super().__new__(cls)

    def __init__(self, microgrid, priority_list=None, remove_redundant_gensets=True):
        """`microgrid`: a pymgrid_b200.Microgrid; like the reference (rbc.py:28-30) the controller works on a COPY that
        carries the microgrid's current state.  `priority_list`: None (ordered by marginal cost, rbc.py:31-44) or one of
        `get_priority_lists()`."""
        if not isinstance(microgrid, Microgrid):
            raise TypeError("RuleBasedControl needs a pymgrid_b200.Microgrid")
        self._remove_redundant_gensets = remove_redundant_gensets
        params = microgrid.export_params()
        self._microgrid = Microgrid(params, device=microgrid._engine.device, obs_order=microgrid._obs_order)
        self._microgrid.trajectory_func = microgrid.trajectory_func
        self._microgrid.raise_errors = microgrid.raise_errors
        self._table = self._engine_table(remove_redundant_gensets)
        # the engine's own table is built with remove_redundant_gensets=True; indices below refer to IT
        self._engine_lists = list(self._microgrid._engine.action_tables[0])
        if priority_list is None:
            chosen = _pl.rbc_priority_list(params, remove_redundant_gensets)
        else:
            chosen = _engine_form(priority_list)
            if chosen is None or chosen not in self._table:
                raise ValueError('Invalid priority list. Use RuleBasedControl.get_priority_lists to view all '
                                 'valid priority lists.')
        if chosen not in self._engine_lists:
            raise NotImplementedError("this priority list switches off a genset whose running_min_production is 0; the "
                                      "engine's action table leaves such lists out (remove_redundant_gensets)")
        self._index = self._engine_lists.index(chosen)
        self._priority_list = _elements(params, chosen)

    def _engine_table(self, remove_redundant_gensets):
        p = self._microgrid.params
        return _pl.priority_lists(p.has_genset, p.has_grid, p.genset.running_min_production if p.genset is not None else None,
                                  remove_redundant_gensets)

    def get_priority_lists(self, remove_redundant_gensets=None):
        """reference: PriorityListAlgo.get_priority_lists (priority_list.py:15-67)"""
        if remove_redundant_gensets is None:
            remove_redundant_gensets = self._remove_redundant_gensets
        return [_elements(self._microgrid.params, pl) for pl in self._engine_table(remove_redundant_gensets)]

    def reset(self):
        return self._microgrid.reset()

    def run(self, max_steps=None, verbose=False):
        """reference: RuleBasedControl.run (rbc.py:64-93): reset, then deploy the list every step until `max_steps` or
        until the microgrid reports done; returns the microgrid's log."""
        if max_steps is None:
            max_steps = len(self._microgrid)
        self.reset()
        self._microgrid.run_priority_list(self._index, max_steps)
        return self._microgrid.get_log(as_frame=True)

    def get_empty_action(self):
        return self._microgrid.get_empty_action()

    @property
    def microgrid(self):
        return self._microgrid

    @property
    def fixed(self):
        return self._microgrid.fixed

    @property
    def flex(self):
        return self._microgrid.flex

    @property
    def modules(self):
        return self._microgrid.modules

    @property
    def priority_list(self):
        return self._priority_list
