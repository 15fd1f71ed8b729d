"""CPU restatement of `Microgrid.run` for ANY module composition -- TEST INFRASTRUCTURE, not product.

Pure-Python scalar IEEE-f64 arithmetic in the reference's operation order (small cases only: tens of modules, hundreds
of steps).  Pinned by tests/test_compose_oracle.py against tests/golden/compose.npz, which tests/golden/make_compose.py
recorded from the live, unmodified reference.  The shipped CUDA path (pymgrid_b200/csrc/mg_compose.cu) never imports
this file; only tests/ may.  Citations are relative to /root/reference/src/pymgrid/.

Modules are the parameter records of pymgrid_b200.modules (the reference's constructor arguments); nothing else of the
package is used here: ordering, physics, observations and the log are restated independently of the host layer under
test (pymgrid_b200/compose.py).
"""
import math
from collections import OrderedDict

import numpy as np

FIXED, FLEX, CONTROLLABLE = "fixed", "flex", "controllable"


def np_sum(values):
    """numpy.sum of a python list of floats as numpy 2.x computes it (pairwise_sum in loops_utils.h): sequential from
    0.0 for fewer than 8 items, eight running sums combined pairwise up to 128 items.  microgrid/utils/step.py:33-36
    sums the provided / absorbed energy lists this way."""
    n = len(values)
    if n < 8:
        res = 0.0
        for v in values:
            res += v
        return res
    assert n <= 128
    r = list(values[:8])
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += values[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += values[i]
        i += 1
    return res


def isclose(a, b, rtol=1e-5, atol=1e-8):
    return abs(a - b) <= atol + rtol * abs(b)


def spread(low, high):
    s = high - low
    return 1.0 if s == 0.0 else s          # utils/space.py:204-205


def normalize(v, low, high):
    return (v - low) / spread(low, high)   # utils/space.py:207-218


def denormalize(x, low, high):
    return low + spread(low, high) * x     # utils/space.py:220-231


class _Mod:
    """live state of one module + its static description"""

    def __init__(self, name, index, rec):
        self.name, self.index, self.rec = name, index, rec
        self.kind, self.dispatch = rec.module_type
        self.is_source = self.kind != "load"
        self.is_sink = self.kind in ("load", "battery", "grid", "balancing")
        self.t = rec.initial_step
        if self.kind == "battery":
            self.charge, self.soc = float(rec.init_charge), float(rec.init_soc)      # battery_module.py:89, 96-106
        if self.kind == "genset":                                                    # genset_module.py:91-92
            self.cs = self.gs = int(rec.init_start_up)
            self.up, self.dn = (0, int(rec.wind_down_time)) if self.cs else (int(rec.start_up_time), 0)
        if self.kind in ("load", "renewable", "grid"):
            ts = rec.time_series
            self.ts, self.T, self.C, self.H = ts, ts.shape[0], ts.shape[1], int(rec.forecast_horizon)
            self.final_step = rec.final_step if rec.final_step > 0 else self.T        # base_timeseries_module.py:317-330
            if self.kind == "grid":        # grid_module.py:125-132: per column, no pull towards zero
                self.low = [float(ts[:, c].min()) for c in range(self.C)]
                self.high = [float(ts[:, c].max()) for c in range(self.C)]
            else:                          # base_timeseries_module.py:81-88
                mn, mx = float(ts.min()), float(ts.max())
                if mn > 0:
                    mn = 0.0
                elif mx < 0:
                    mx = 0.0
                self.low, self.high = [mn], [mx]

    # ---- observation / state ----
    def series_state(self):
        """unnormalised [current(C), forecast rows...]: base_timeseries_module.py:103-140, forecaster.py:95,120-149"""
        out = []
        for k in range(self.H + 1):
            idx = self.t + k
            for c in range(self.C):
                if idx < self.T and self.t < self.T:
                    v = float(self.ts[idx, c])
                    if k > 0:
                        v = min(max(v, self.low[c]), self.high[c])
                else:
                    v = (self.high[c] + self.low[c]) / 2
                    if k > 0 and self.t < self.T:
                        v = min(max(v, self.low[c]), self.high[c])
                out.append(v)
        return out

    def state(self):
        k = self.kind
        if k in ("load", "renewable", "grid"):
            return self.series_state()
        if k == "battery":
            return [self.soc, self.charge]
        if k == "genset":
            return [float(self.cs), float(self.gs), float(self.up), float(self.dn)]
        return []

    def obs_bounds(self):
        k, r = self.kind, self.rec
        if k in ("load", "renewable", "grid"):
            return self.low * (self.H + 1), self.high * (self.H + 1)
        if k == "battery":                  # battery_module.py:323-330
            return [r.min_capacity / r.max_capacity, r.min_capacity], [1.0, r.max_capacity]
        if k == "genset":                   # genset_module.py:503-509
            return [0.0] * 4, [1.0, 1.0, float(r.start_up_time), float(r.wind_down_time)]
        return [], []

    def obs(self):
        lo, hi = self.obs_bounds()
        return np.array([normalize(v, a, b) for v, a, b in zip(self.state(), lo, hi)], dtype=np.float64)

    def state_keys(self):
        k = self.kind
        if k in ("load", "renewable", "grid"):
            comps = {"load": ["load"], "renewable": ["renewable"],
                     "grid": ["import_price", "export_price", "co2_per_kwh", "grid_status"]}[k]
            return [f"{c}_current" for c in comps] + [f"{c}_forecast_{j}" for j in range(self.H) for c in comps]
        if k == "battery":
            return ["soc", "current_charge"]
        if k == "genset":
            return ["current_status", "goal_status", "steps_until_up", "steps_until_down"]
        return []

    # ---- limits ----
    def max_production(self):
        k, r = self.kind, self.rec
        if k == "battery":
            return min(r.max_discharge, self.charge - r.min_capacity) * r.efficiency     # battery_module.py:283-286
        if k == "genset":
            return self.cs * r.running_max_production                                     # genset_module.py:466-482
        if k == "grid":
            return r.max_import * float(self.ts[self.t, 3])                               # grid_module.py:314-316
        if k == "renewable":
            return float(self.ts[self.t, 0])                                              # renewable_module.py:95-110
        return math.inf

    def min_production(self):
        return self.cs * self.rec.running_min_production if self.kind == "genset" else 0.0

    def max_consumption(self):
        k, r = self.kind, self.rec
        if k == "battery":
            return min(r.max_charge, r.max_capacity - self.charge) / r.efficiency         # battery_module.py:288-291
        if k == "grid":
            return r.max_export * float(self.ts[self.t, 3])                               # grid_module.py:318-320
        return math.inf

    # ---- genset state machine: genset_module.py:216-346 ----
    def _reset_up_down(self):
        r = self.rec
        if self.cs:
            self.up, self.dn = 0, int(r.wind_down_time)
        else:
            self.up, self.dn = int(r.start_up_time), 0

    def update_status(self, goal_status):
        r = self.rec
        goal = round(goal_status)
        if goal == self.cs == self.gs:
            return
        instant_up = r.start_up_time == 0 and goal == 1
        instant_down = r.wind_down_time == 0 and goal == 0
        if goal != self.gs and (r.allow_abortion or instant_up or instant_down):
            self.gs = goal
        if self.up == 0 and self.gs == 1:
            self.cs = 1
            self._reset_up_down()
            return
        if self.dn == 0 and self.gs == 0:
            self.cs = 0
            self._reset_up_down()
            return
        if goal == self.cs and self.cs != self.gs and r.allow_abortion:
            self.gs = goal
            self._reset_up_down()
        elif self.cs == self.gs and self.gs != goal:
            self._reset_up_down()
            self.gs = goal
        if self.gs != self.cs:
            if self.gs == 0:
                self.dn -= 1
            else:
                self.up -= 1


class Raised(Exception):
    """the reference raised `kind` (an exception class name) inside this step"""

    def __init__(self, kind, message=""):
        super().__init__(f"{kind}: {message}")
        self.kind = kind


class ComposedOracle:
    def __init__(self, modules, add_unbalanced_module=True, loss_load_cost=10.0, overgeneration_cost=2.0,
                 reward_shaping_func=None, trajectory_func=None):
        """modules: list of pymgrid_b200.modules records or (name, record) tuples -- microgrid.py:100-165"""
        self.reward_shaping_func, self.trajectory_func = reward_shaping_func, trajectory_func
        from pymgrid_b200.modules import UnbalancedEnergyModule
        named = []
        for m in modules:
            name, rec = m if isinstance(m, tuple) else (None, m)
            named.append((name if name is not None else rec.module_type[0], rec))
        if add_unbalanced_module:
            named.append(("balancing", UnbalancedEnergyModule(False, loss_load_cost=loss_load_cost,
                                                              overgeneration_cost=overgeneration_cost)))
        # module_container.py:355-413: (fixed, flex, controllable) x (sources, sinks, source_and_sinks), names in
        # insertion order inside each cell, modules of one name in insertion order
        cells = OrderedDict(((d, s), OrderedDict()) for d in (FIXED, FLEX, CONTROLLABLE)
                            for s in ("sources", "sinks", "source_and_sinks"))
        for name, rec in named:
            probe = _Mod(name, 0, rec)
            s = "source_and_sinks" if probe.is_sink and probe.is_source else ("sources" if probe.is_source else "sinks")
            lst = cells[(rec.module_type[1], s)].setdefault(name, [])
            lst.append(_Mod(name, len(lst), rec))
        self.by_name = OrderedDict()
        for cell in cells.values():          # Container.to_dict: later cells overwrite equal names (none in practice)
            for name, lst in cell.items():
                self.by_name[name] = lst
        self.listing = [m for lst in self.by_name.values() for m in lst]
        self.log_rows = []
        self.initial_step = self.listing[0].t if self.listing else 0

    def _of(self, dispatch):
        return [(name, lst) for name, lst in self.by_name.items() if lst[0].dispatch == dispatch]

    @property
    def current_step(self):
        return self.listing[0].t

    def reset(self):
        """microgrid.py:205-225 + base_module.py:65-77: only the step counter moves"""
        initial = self.initial_step
        if self.trajectory_func is not None:      # microgrid.py:221-225, 652-684: the modules' window moves, the microgrid's stays
            ts = [m for m in self.listing if m.kind in ("load", "renewable", "grid")]
            final0 = ts[0].rec.final_step if ts[0].rec.final_step > 0 else ts[0].T
            initial, final = self.trajectory_func(self.initial_step, final0)
            for m in ts:
                m.final_step = final
        for m in self.listing:
            m.t = initial
        self.log_rows = []
        return OrderedDict((name, [m.obs() for m in lst]) for name, lst in self.by_name.items())

    # ---- one module step: base_module.py:95-274 + the module's update() ----
    def _module_step(self, m, action, normalized):
        k, r = m.kind, m.rec
        state_pre = None
        if k == "genset":                                        # genset_module.py:100-149
            goal = float(action[0])
            if not 0 <= goal <= 1:
                raise Raised("AssertionError", "genset goal outside [0, 1]")
            m.update_status(goal)
            a = denormalize(float(action[1]), 0.0, r.running_max_production) if normalized else float(action[1])
        elif k == "battery":
            a = float(action)
            if normalized:
                a = denormalize(a, -r.max_discharge / r.efficiency, r.max_charge * r.efficiency)
        elif k == "grid":
            a = float(action)
            if normalized:
                a = denormalize(a, -1 * r.max_export, r.max_import)
        else:
            a = float(action)
        state_pre = OrderedDict(zip(m.state_keys(), m.state()))   # base_module.py:152 (after the genset status update)
        if k in ("load", "renewable", "grid") and m.t >= m.T:
            raise Raised("IndexError", f"index {m.t} is out of bounds")
        info = OrderedDict()
        reward, done = 0.0, False
        as_source = a > 0 or (a == 0 and m.is_source)
        if as_source:
            if not m.is_source:
                raise Raised("AssertionError", "not a source")
            if m.dispatch == FIXED:
                energy = None
            else:
                mx, mn = m.max_production(), m.min_production()
                if a > mx:
                    energy, clipped = mx, True
                elif a < mn:
                    energy, clipped = mn, True
                else:
                    energy, clipped = a, False
                if clipped and r.raise_errors:
                    raise Raised("ValueError", "production outside the module's limits")
        else:
            e = -1.0 * a
            if m.dispatch == FIXED:
                energy = None
            else:
                if k in ("genset", "renewable"):      # max_consumption is NotImplemented: the comparison at base_module.py:265 fails
                    raise Raised("TypeError", "'>' not supported between instances of 'float' and 'NotImplementedType'")
                mc = m.max_consumption()
                if e > mc:
                    if r.raise_errors:
                        raise Raised("ValueError", "consumption outside the module's limits")
                    energy = mc
                else:
                    energy = e
                if not energy >= 0:
                    raise Raised("AssertionError", "absorbed_energy >= 0")
        # ---- update() ----
        if k == "load":                                           # load_module.py:86-91
            info["absorbed_energy"] = -1 * float(m.ts[m.t, 0])
            done = m.t >= m.final_step - 1
        elif k == "renewable":                                    # renewable_module.py:86-93
            cur = float(m.ts[m.t, 0])
            info["provided_energy"] = energy
            info["curtailment"] = cur - energy
            done = m.t >= m.final_step - 1
        elif k == "battery":                                      # battery_module.py:108-130, 244-278
            if as_source:
                info["provided_energy"] = energy
                internal = (-1.0 * energy) / r.efficiency
            else:
                info["absorbed_energy"] = energy
                internal = energy * r.efficiency
            m.charge += internal
            if m.charge < r.min_capacity:
                if not isclose(m.charge, r.min_capacity):
                    raise Raised("AssertionError", "battery below min_capacity")
                m.charge = r.min_capacity
            m.soc = m.charge / r.max_capacity
            reward = -1.0 * (abs(internal) * r.battery_cost_cycle)
        elif k == "genset":                                       # genset_module.py:151-214
            co2 = r.co2_per_unit * energy
            cost = r.genset_cost * energy + r.cost_per_unit_co2 * co2
            reward = -1.0 * cost
            info["provided_energy"] = energy
            info["co2_production"] = co2
        elif k == "grid":                                         # grid_module.py:134-228
            imp, exp_, co2k = (float(m.ts[m.t, c]) for c in range(3))
            if as_source:
                co2 = energy * co2k
                reward = -1 * imp * energy + (-1.0 * r.cost_per_unit_co2 * co2)
                info["provided_energy"] = energy
                info["co2_production"] = co2
            else:
                reward = exp_ * energy + (-1.0 * r.cost_per_unit_co2 * 0.0)
                info["absorbed_energy"] = energy
                info["co2_production"] = 0.0
            done = m.t >= m.final_step - 1
        elif k == "balancing":                                    # unbalanced_energy_module.py:28-70
            if as_source:
                reward = -1.0 * (r.loss_load_cost * energy)
                info["provided_energy"] = energy
            else:
                reward = -1.0 * (r.overgeneration_cost * energy)
                info["absorbed_energy"] = energy
        # ---- _log: base_module.py:276-290 ----
        names = {"load": (None, "load_met"), "renewable": ("renewable_used", None), "battery": ("discharge_amount", "charge_amount"),
                 "genset": ("genset_production", None), "grid": ("grid_import", "grid_export"),
                 "balancing": ("loss_load", "overgeneration")}[k]
        row = OrderedDict(reward=reward)
        for key, v in info.items():
            if key not in ("provided_energy", "absorbed_energy"):
                row[key] = v
        if names[0] is not None:
            row[names[0]] = info.get("provided_energy", 0.0)
        if names[1] is not None:
            row[names[1]] = info.get("absorbed_energy", 0.0)
        row.update(state_pre)
        m.t += 1
        return m.obs(), reward, done, info, row

    def run(self, control, normalized=True):
        """microgrid.py:227-325 -> (obs dict, reward, done, info dict); the log row is appended to self.log_rows"""
        obs, info_out = OrderedDict(), OrderedDict()
        cost_info = self.cost_info()              # microgrid.py:253, before any module steps
        shaped = lambda: (self.reward_shaping_func(OrderedDict(info_out), cost_info)      # noqa: E731  step.py:41-46
                          if self.reward_shaping_func is not None else reward_sum)
        reward_sum, done_any = 0.0, False
        provided, absorbed = [], []
        logs = {}

        def append(name, m, out):
            nonlocal reward_sum, done_any
            o, r, d, inf, row = out
            obs.setdefault(name, []).append(o)
            info_out.setdefault(name, []).append(inf)
            reward_sum += r
            done_any = done_any or d
            if "provided_energy" in inf:
                provided.append(inf["provided_energy"])
            if "absorbed_energy" in inf:
                absorbed.append(inf["absorbed_energy"])
            logs[(name, m.index)] = row

        for name, lst in self._of(FIXED):
            for m in lst:
                append(name, m, self._module_step(m, 0.0, False))
        fixed_p, fixed_a = np_sum(provided), np_sum(absorbed)
        shaped()                                  # balance() evaluates the shaper every time it is called (step.py:33-36)
        control = dict(control)
        for name, lst in self._of(CONTROLLABLE):
            if name not in control:
                raise Raised("ValueError", f'Control for module "{name}" not found')
            acts = control.pop(name)
            try:
                pairs = list(zip(lst, acts))
            except TypeError:
                pairs = list(zip(lst, [acts]))
            for m, a in pairs:
                append(name, m, self._module_step(m, a, normalized))
        p, a = np_sum(provided), np_sum(absorbed)
        shaped()
        difference = p - a
        ctrl_p, ctrl_a = p - fixed_p, a - fixed_a
        if difference > 0:
            excess = difference
            for name, lst in self._of(FLEX):
                for m in lst:
                    if not m.is_sink:
                        amt = 0.0
                    elif m.max_consumption() < excess:
                        amt = -1.0 * m.max_consumption()
                    else:
                        amt = -1.0 * excess
                    append(name, m, self._module_step(m, amt, False))
                    excess += amt
        else:
            needed = -difference
            for name, lst in self._of(FLEX):
                for m in lst:
                    if not m.is_source:
                        amt = 0.0
                    elif m.max_production() < needed:
                        amt = m.max_production()
                    else:
                        amt = needed
                    append(name, m, self._module_step(m, amt, False))
                    needed -= amt
        p, a = np_sum(provided), np_sum(absorbed)
        shaped_reward = shaped()
        row = OrderedDict()
        for m in self.listing:
            for key, v in logs[(m.name, m.index)].items():
                row[(m.name, m.index, key)] = v
        for key, v in (("reward", reward_sum), ("shaped_reward", shaped_reward), ("overall_provided_to_microgrid", p),
                       ("overall_absorbed_from_microgrid", a), ("controllable_provided_to_microgrid", ctrl_p),
                       ("controllable_absorbed_from_microgrid", ctrl_a), ("fixed_provided_to_microgrid", fixed_p),
                       ("fixed_absorbed_from_microgrid", fixed_a)):
            row[("balance", 0, key)] = v
        self.log_rows.append(row)
        if not isclose(p, a):
            raise Raised("RuntimeError", "Microgrid modules unable to balance energy production with consumption.")
        return obs, shaped(), done_any, info_out

    def cost_info(self):
        """Microgrid.get_cost_info (microgrid.py:334-335): each module's marginal costs at the CURRENT step"""
        out = OrderedDict()
        for name, lst in self.by_name.items():
            out[name] = []
            for m in lst:
                k, r = m.kind, m.rec
                if k == "battery":
                    pc = ac = r.battery_cost_cycle                                          # battery_module.py:340-346
                elif k == "genset":
                    pc, ac = r.genset_cost * 1.0 + r.cost_per_unit_co2 * (r.co2_per_unit * 1.0), 0.0      # genset_module.py:519-521
                elif k == "grid":
                    row = m.series_state()                                                  # grid_module.py:322-328: state[0], state[1]
                    pc, ac = row[0], row[1]
                elif k == "balancing":
                    pc, ac = r.loss_load_cost, r.overgeneration_cost
                else:
                    pc = ac = 0.0
                out[name].append(dict(production_marginal_cost=pc, absorption_marginal_cost=ac))
        return out


# ---- priority lists: algos/priority_list/priority_list.py:15-167, priority_list_element.py:6-80 ------------------------
def _marginal_cost(m):
    k, r = m.kind, m.rec
    if k == "battery":
        return r.battery_cost_cycle                                            # battery_module.py:340-346
    if k == "genset":
        return r.genset_cost * 1.0 + r.cost_per_unit_co2 * (r.co2_per_unit * 1.0)   # genset_module.py:519-521
    return float(m.ts[m.t, 0])                                                 # grid_module.py:322-324


def oracle_priority_lists(orc, remove_redundant_gensets=False):
    """every deployment order of the controllable modules as tuples of (name, index, module_actions, action)"""
    from itertools import permutations
    ctl = [m for _, lst in orc._of(CONTROLLABLE) for m in lst]
    ordered = [m for m in ctl if not m.is_sink] + [m for m in ctl if m.is_sink]      # sources, then source_and_sinks
    elements = []
    for m in ordered:
        n = 2 if m.kind == "genset" else 1
        elements += [(m.name, m.index, n, a) for a in range(n)]
    out, seen = [], set()
    for perm in permutations(elements):
        listed, pl = set(), []
        for el in perm:
            if el[:2] not in listed:
                listed.add(el[:2])
                pl.append(el)
        pl = tuple(pl)
        if pl not in seen:
            seen.add(pl)
            out.append(pl)
    if remove_redundant_gensets:
        redundant = [(m.name, m.index, 2, 0) for m in orc.listing if m.kind == "genset" and m.rec.running_min_production == 0]
        out = [pl for pl in out if not any(r in pl for r in redundant)]
    return out


def oracle_rbc_list(orc, remove_redundant_gensets=True):
    """RuleBasedControl's automatic list (algos/rbc/rbc.py:31-44): the first priority list sorted by marginal cost, ties
    towards the higher action number"""
    first = oracle_priority_lists(orc, remove_redundant_gensets)[0]
    cost = {(m.name, m.index): _marginal_cost(m) for m in orc.listing if m.dispatch == CONTROLLABLE}
    return sorted(first, key=lambda el: (cost[el[:2]], -el[3]))


def oracle_priority_control(orc, pl):
    """_populate_action (:69-116): priority list -> {name: [unnormalised action per module]}"""
    action = {name: [None] * len(lst) for name, lst in orc._of(CONTROLLABLE)}
    total_load = 0.0
    for m in orc.listing:
        if m.kind == "load":
            total_load += -1 * float(m.ts[m.t, 0])
    renewable = np_sum([float(m.ts[m.t, 0]) for m in orc.listing if m.kind == "renewable"])
    remaining = total_load - renewable
    for name, index, n_actions, act in pl:
        m = orc.by_name[name][index]
        if n_actions > 1:
            if action[name][index] is not None:
                continue
            action[name][index] = [act]
        if isclose(remaining, 0.0, atol=1e-4):
            energy = 0.0
        elif remaining > 0:
            if m.kind == "genset":
                nxt = (1 if (m.cs or m.up == 0) else 0) if act else (0 if (not m.cs or m.dn == 0) else 1)
                mx, mn = nxt * m.rec.running_max_production, nxt * m.rec.running_min_production
            else:
                mx, mn = m.max_production(), m.min_production()
            energy = remaining if mn <= remaining <= mx else (mn if remaining < mn else mx)
        else:
            if m.is_sink:
                mc = m.max_consumption()
                if not mc >= 0:
                    raise Raised("AssertionError", "module_max_consumption >= 0")
                energy = -1.0 * mc if -1 * remaining > mc else remaining
            else:
                energy = 0.0
        if n_actions > 1:
            action[name][index] = np.array(action[name][index] + [energy])
        else:
            action[name][index] = energy
        remaining -= energy
    return action
