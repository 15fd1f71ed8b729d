"""`microgrid.modules...` read-only views (what RBC / priority lists / MPC / notebooks read) against values
recorded from the live reference (tests/golden/views.npz).  The views only need the params and the (step, charge, genset)
state, so they are checked here without a GPU by standing in for the engine-backed `_state()`; tests/test_gpu_dropin.py
checks the same attributes once more on top of the real engine."""
import numpy as np
import pytest

from pymgrid_b200.microgrid import Microgrid, ModuleContainerView, ModuleView
from pymgrid_b200.scenario import load_pymgrid25

SCALARS = ("marginal_cost", "production_marginal_cost", "absorption_marginal_cost", "max_production", "max_consumption",
           "min_production", "is_source", "is_sink")
VECTORS = ("state", "min_obs", "max_obs", "min_act", "max_act")


class HostBackedMicrogrid(Microgrid):
    """Microgrid whose live state is a host tuple instead of the device arrays (view logic only)."""

    def __init__(self, params, state):
        self.params = params
        self._st = dict(t=int(state[0]), charge=float(state[1]), genset=tuple(int(x) for x in state[2:6]),
                        soc=params.battery.reported_soc if float(state[1]) == params.battery.current_charge
                        else float(state[1]) / params.battery.max_capacity)
        self._initial_step, self._final_step = params.initial_step, params.final_step
        names = ["load", "pv", "unbalanced_energy"] + (["genset"] if params.has_genset else []) + ["battery"] + \
                (["grid"] if params.has_grid else [])
        self._modules = ModuleContainerView((n, [ModuleView(self, n)]) for n in names)

    def _state(self):
        return self._st

    @property
    def current_step(self):
        return self._st["t"]

    def __del__(self):
        pass


def check_views(m, z, n):
    for name, lst in m.modules.iterdict():
        mod = lst[0]
        for a in SCALARS:
            key = f"s{n}_{name}_{a}"
            if key in z:
                assert float(getattr(mod, a)) == float(z[key]), key
        for a in VECTORS:
            got = np.atleast_1d(np.asarray(getattr(mod, a), dtype=np.float64)).ravel()
            np.testing.assert_array_equal(got, z[f"s{n}_{name}_{a}"], err_msg=f"{name}.{a}")
        assert ",".join(mod.module_type) == str(z[f"s{n}_{name}_type"])
        assert mod.action_space.shape[0] == int(z[f"s{n}_{name}_n_act"]), name
        np.testing.assert_array_equal(np.atleast_1d(mod.to_normalized(mod.state, obs=True)).ravel(), z[f"s{n}_{name}_norm_state"])
        np.testing.assert_array_equal(np.atleast_1d(mod.from_normalized(z[f"s{n}_{name}_norm_state"], obs=True)).ravel().shape,
                                      z[f"s{n}_{name}_state"].shape)
    if m.params.has_genset:
        g = m.modules.genset[0]
        got = [g.next_status(0), g.next_status(1), g.next_max_production(0), g.next_max_production(1),
               g.next_min_production(0), g.next_min_production(1)]
        np.testing.assert_array_equal(np.array(got, dtype=np.float64), z[f"s{n}_genset_next"])
    if m.params.has_grid:
        g = m.modules["grid"][0]
        np.testing.assert_array_equal(np.stack([g.import_price, g.export_price, g.co2_per_kwh]), z[f"s{n}_grid_columns"])
    b = m.modules.battery[0]
    np.testing.assert_array_equal(np.array([b.soc, b.min_soc, b.max_soc]), z[f"s{n}_battery_socs"])
    for kind in ("fixed", "flex", "controllable"):
        for sub in ("sources", "sinks", "source_and_sinks"):
            assert ",".join(getattr(getattr(m, kind), sub).keys()) == str(z[f"s{n}_{kind}_{sub}"]), (kind, sub)
    assert m.get_forecast_horizon() == int(z[f"s{n}_horizon"]) and len(m) == 8760
    # Microgrid.from_normalized on an action dict (microgrid.py:412-432)
    flat, i, ctrl = z[f"s{n}_denorm_in"], 0, {}
    for name in ("genset", "battery", "grid"):
        if name in m.modules:
            w = 2 if name == "genset" else 1
            ctrl[name] = [flat[i:i + w] if w == 2 else float(flat[i])]
            i += w
    out = m.from_normalized(ctrl, act=True)
    np.testing.assert_array_equal(np.concatenate([np.ravel(v[0]) for v in out.values()]), z[f"s{n}_denorm_out"])
    back = m.to_normalized(out, act=True)
    np.testing.assert_allclose(np.concatenate([np.ravel(v[0]) for v in back.values()]), flat, rtol=1e-12)


@pytest.mark.parametrize("n", (0, 1, 2))
def test_module_views_match_reference(golden, n):
    z = golden["views"]
    check_views(HostBackedMicrogrid(load_pymgrid25(n), z[f"s{n}_state"]), z, n)
