"""What the reference's balance tests pin (tests/microgrid/test_microgrid.py:188-455, the TestMicrogridLoadPV family: a
microgrid of loads and renewables only -- equal series, PV surplus, load surplus, two loads, two PVs, two of each, three to
nine of each, and a Python reward shaper), written for this package.  The facts checked are the reference's; the code is
ours: one table of variants, one builder, plain check functions.  `checks(library)` binds them to a backend -- the host
build of the composed path's C source (CPU suite; a callable, so that nothing is compiled at collection time) or None = the
CUDA path (GPU suite).

Not covered: conversion to the deprecated non-modular stack (test_to_nonmodular), which is out of scope.
"""
import numpy as np
import pandas as pd
import pytest

import pymgrid_b200
from pymgrid_b200.modules import LoadModule, RenewableModule

STEPS = 100
# variant -> (number of loads, number of renewables (None = random 3..9), which side gets a surplus, reward shaper?)
VARIANTS = {
    "one_each": (1, 1, None, False), "pv_surplus": (1, 1, "pv", False), "load_surplus": (1, 1, "load", False),
    "two_loads": (2, 1, None, False), "two_pv": (1, 2, None, False), "two_each": (2, 2, None, False),
    "many_each": (None, None, None, False), "many_each_pv_surplus": (None, None, "pv", False),
    "many_each_load_surplus": (None, None, "load", False), "python_reward_shaper": (1, 1, None, True),
}
CLOSE = dict(rtol=1e-7, atol=1e-10)      # the tolerance of the reference's own assertEqual (tests/helpers/test_case.py:6-25)


def energy_times_marginal_cost(energy_info, cost_info):
    """the shaper of the reference's TestMicrogridRewardShaping: every module's energy times its marginal cost"""
    cost = 0
    for name, infos in energy_info.items():
        for info in infos:
            for position, (kind, amount) in enumerate(info.items()):
                if kind in ("absorbed_energy", "provided_energy"):
                    key = "absorption_marginal_cost" if kind == "absorbed_energy" else "production_marginal_cost"
                    cost += amount * cost_info[name][position][key]
    return cost


def split_positive(rng, total, parts):
    """`parts` positive series that add up to `total`"""
    out, rest = [], total.copy()
    for _ in range(parts - 1):
        out.append(rest * (1 - rng.random(total.shape)))
        rest = rest - out[-1]
    return out + [rest]


class Variant:
    def __init__(self, name, library):
        n_loads, n_pvs, surplus, shaped = VARIANTS[name]
        rng = np.random.default_rng(sum(map(ord, name)))
        base = 10 * rng.random(STEPS)
        extra = 5 * rng.random(STEPS)
        self.load = base + (extra if surplus == "load" else 0)
        self.pv = base + (extra if surplus == "pv" else 0)
        self.n_loads = n_loads or int(rng.integers(3, 10))
        self.n_pvs = n_pvs or int(rng.integers(3, 10))
        modules = [LoadModule(time_series=ts, raise_errors=self.n_loads <= 2) for ts in split_positive(rng, self.load, self.n_loads)]
        modules += [RenewableModule(time_series=ts) for ts in split_positive(rng, self.pv, self.n_pvs)]
        if callable(library):          # resolved when a test runs, not when the module is collected
            library = library()
        from tests.hostsim import select
        select(library)        # the host build of the kernel source (CPU suite) or, with None, the product's CUDA library
        kw = {}
        self.microgrid = pymgrid_b200.Microgrid(modules, **kw)
        if shaped:      # rebuilt from the first microgrid's own modules, slack module included
            self.microgrid = pymgrid_b200.Microgrid(self.microgrid.modules.to_tuples(), add_unbalanced_module=False,
                                                    reward_shaping_func=energy_times_marginal_cost, **kw)


def checks(library):
    names = list(VARIANTS)

    @pytest.mark.parametrize("name", names)
    def test_structure_and_current_values(name):
        v = Variant(name, library)
        m = v.microgrid
        assert hasattr(m.modules, "load") and hasattr(m.modules, "renewable")
        assert len(m.modules) == 1 + v.n_loads + v.n_pvs                      # + the slack module
        np.testing.assert_allclose(sum(x.current_load for x in m.modules.load), v.load[0], **CLOSE)
        np.testing.assert_allclose(sum(x.current_renewable for x in m.modules.renewable), v.pv[0], **CLOSE)
        if v.n_loads == 1:
            assert m.modules.load.item().current_load == v.load[0]
        else:
            with pytest.raises(ValueError):
                m.modules.load.item()

    @pytest.mark.parametrize("name", names)
    def test_actions_and_state_views(name):
        v = Variant(name, library)
        m = v.microgrid
        assert m.sample_action() == {} and m.get_empty_action() == {}         # nothing controllable
        flex = m.sample_action(sample_flex_modules=True)
        assert set(flex) == {"renewable", "balancing"} and len(flex["renewable"]) == v.n_pvs
        state = m.state_dict()
        assert {"load", "renewable", "balancing"} <= set(state)
        assert len(state["load"]) == v.n_loads and len(state["balancing"]) == 1
        series = m.state_series()
        assert set(series.index.get_level_values(0)) == {"load", "renewable"}
        assert series["load"].index.get_level_values(0).nunique() == v.n_loads
        assert series["renewable"].index.get_level_values(0).nunique() == v.n_pvs

    @pytest.mark.parametrize("name", names)
    def test_every_step_balances_and_logs(name):
        v = Variant(name, library)
        m = v.microgrid
        loss_load_cost = m.modules.balancing[0].loss_load_cost
        for t in range(STEPS):
            obs, reward, done, info = m.run(m.get_empty_action())
            met = min(v.load[t], v.pv[t])
            unmet, curtailed = max(v.load[t] - met, 0), max(v.pv[t] - met, 0)
            np.testing.assert_allclose(-reward, loss_load_cost * max(v.load[t] - v.pv[t], 0), **CLOSE)
            log = m.log
            assert len(log) == t + 1 and all(n in log for n in m.modules.names())
            row = log.iloc[t]
            total = lambda module, field: row.loc[pd.IndexSlice[module, :, field]].sum()      # noqa: E731
            assert row["load"].index.get_level_values(0).nunique() == v.n_loads
            want = {("load", "load_current"): -v.load[t], ("load", "load_met"): v.load[t],
                    ("renewable", "renewable_current"): v.pv[t], ("renewable", "renewable_used"): met,
                    ("renewable", "curtailment"): curtailed, ("balancing", "loss_load"): unmet,
                    ("balance", "reward"): -loss_load_cost * unmet,
                    ("balance", "overall_provided_to_microgrid"): v.load[t],
                    ("balance", "overall_absorbed_from_microgrid"): v.load[t],
                    ("balance", "fixed_provided_to_microgrid"): 0.0, ("balance", "fixed_absorbed_from_microgrid"): v.load[t],
                    ("balance", "controllable_provided_to_microgrid"): 0.0,
                    ("balance", "controllable_absorbed_from_microgrid"): 0.0}
            for (module, field), value in want.items():
                np.testing.assert_allclose(total(module, field), value, err_msg=f"{module}.{field} at step {t}", **CLOSE)

    return [test_structure_and_current_values, test_actions_and_state_views, test_every_step_balances_and_logs]
