"""Loader for tests/golden/compose.npz (recorded from the live reference by tests/golden/make_compose.py): rebuilds each
case's module list with pymgrid_b200.modules' classes from the JSON spec stored beside the recorded outputs."""
import json
import os

import numpy as np

from pymgrid_b200 import modules as M

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INFO_SLOTS = 5
BALANCE_COLS = ("reward", "shaped_reward", "overall_provided_to_microgrid", "overall_absorbed_from_microgrid",
                "controllable_provided_to_microgrid", "controllable_absorbed_from_microgrid",
                "fixed_provided_to_microgrid", "fixed_absorbed_from_microgrid")


def marginal_cost_total(energy_info, cost_info):
    """the reward shaper of the reference's TestMicrogridRewardShaping (tests/microgrid/test_microgrid.py:436-455), with the
    module index taken from the list position: energy x marginal cost over all modules"""
    total = 0
    for module_name, info_list in energy_info.items():
        for module_n, module_info in enumerate(info_list):
            for energy_type, energy_amount in module_info.items():
                if energy_type == 'absorbed_energy':
                    marginal_cost = cost_info[module_name][module_n]['absorption_marginal_cost']
                elif energy_type == 'provided_energy':
                    marginal_cost = cost_info[module_name][module_n]['production_marginal_cost']
                else:
                    continue
                total += energy_amount * marginal_cost
    return total


def window_trajectory(lo_offset, hi_offset):
    """a deterministic trajectory_func (microgrid/trajectory/deterministic.py): the window [initial + lo, final + hi]"""
    return lambda initial_step, final_step: (initial_step + lo_offset, final_step + hi_offset)


def callable_kwargs(spec):
    """the Python callables a case's Microgrid takes, rebuilt from their names in the JSON spec"""
    kw = {}
    if spec.get("shaper") == "marginal_cost_total":
        kw["reward_shaping_func"] = marginal_cost_total
    if spec.get("trajectory"):
        kw["trajectory_func"] = window_trajectory(*spec["trajectory"])
    return kw


class ComposeCase:
    def __init__(self, data, i):
        self.i = i
        self._d, self._p = data, f"c{i}_"
        self.spec = json.loads(str(data[self._p + "spec"]))
        self.label = self.spec["label"]
        self.names = [tuple(x) for x in json.loads(str(data[self._p + "names"]))]      # (name, index, class) listing order

    def __getitem__(self, key):
        return self._d[self._p + key]

    def json(self, key):
        return json.loads(str(self._d[self._p + key]))

    def modules(self):
        out = []
        for e in self.spec["modules"]:
            kw = dict(e["kwargs"])
            if e["ts"] is not None:
                kw["time_series"] = self._d[f"{self._p}ts{e['ts']}"]
            m = getattr(M, e["cls"])(**kw)
            out.append((e["name"], m) if e["name"] is not None else m)
        return out

    @property
    def microgrid_kwargs(self):
        return dict(self.spec["microgrid_kwargs"])

    @property
    def callable_kwargs(self):
        return callable_kwargs(self.spec)

    def control(self, k, controllable):
        """the recorded action row of step k -> {name: [action per module]}; `controllable`: [(name, [n_act per module])]"""
        row, col, out = self["actions"][k], 0, {}
        for name, widths in controllable:
            vals = []
            for w in widths:
                vals.append(np.array(row[col:col + w]) if w > 1 else float(row[col]))
                col += w
            out[name] = vals
        return out


def load_cases(file="compose.npz"):
    data = np.load(os.path.join(GOLDEN_DIR, file))
    return [ComposeCase(data, i) for i in range(int(data["n_cases"]))]
