"""Host-side enumeration of the discrete action table (priority lists).

Restates `PriorityListAlgo.get_priority_lists` (reference: algos/priority_list/priority_list.py:15-67) for the
module set on the batched path: one element per action-space dimension of every controllable source (the
genset: goal 0, goal 1), then one per controllable source-and-sink (battery, grid); all permutations; within a
permutation later elements of an already-listed module are dropped; duplicates removed in first-seen order;
optionally lists that switch a genset with running_min_production == 0 off are removed (:53-67).
The table is tiny (<= 12 x 3) and is computed once per configuration; the per-step EXPANSION of a list into
controls (:69-116) runs on the device inside mg_step_discrete.
"""
from itertools import permutations

GENSET, BATTERY, GRID = 0, 1, 2
MODULE_NAMES = {GENSET: "genset", BATTERY: "battery", GRID: "grid"}


def priority_lists(has_genset, has_grid, genset_running_min=None, remove_redundant_gensets=True):
    """Returns a list of priority lists; each is a tuple of (module, action) pairs in deployment order."""
    elements = []
    if has_genset:
        elements += [(GENSET, 0), (GENSET, 1)]          # controllable.sources, one element per action dim
    elements.append((BATTERY, 0))                        # controllable.source_and_sinks, insertion order
    if has_grid:
        elements.append((GRID, 0))
    seen, out = set(), []
    for perm in permutations(elements):
        listed, pl = set(), []
        for el in perm:
            if el[0] not in listed:
                listed.add(el[0])
                pl.append(el)
        pl = tuple(pl)
        if pl not in seen:
            seen.add(pl)
            out.append(pl)
    if remove_redundant_gensets and has_genset and genset_running_min == 0:
        out = [pl for pl in out if (GENSET, 0) not in pl]
    return out


def marginal_costs(p):
    """`module.marginal_cost` of the controllable modules as the reference evaluates them when the controller is built:
    battery_module.py:340-346, genset_module.py:519-521 (get_cost(1.0)), grid_module.py:322-324 (current import price)."""
    out = {BATTERY: p.battery.battery_cost_cycle}
    if p.genset is not None:
        g = p.genset
        out[GENSET] = g.genset_cost * 1.0 + g.cost_per_unit_co2 * (g.co2_per_unit * 1.0)
    if p.grid is not None:
        out[GRID] = float(p.grid.time_series[p.current_step, 0])
    return out


def rbc_priority_list(p, remove_redundant_gensets=True):
    """RuleBasedControl's automatic list (algos/rbc/rbc.py:31-44): the FIRST priority list sorted by marginal cost,
    ties broken towards the higher action number (priority_list_element.py:73-80).  Note the reference quirk kept
    here: on genset grids the first list carries genset goal 0 (unless it was removed as redundant)."""
    first = priority_lists(p.has_genset, p.has_grid, p.genset.running_min_production if p.genset is not None else None,
                           remove_redundant_gensets)[0]
    cost = marginal_costs(p)
    return tuple(sorted(first, key=lambda el: (cost[el[0]], -el[1])))
