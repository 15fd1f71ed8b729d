"""`pymgrid_b200.modules`: the reference's module constructors (src/pymgrid/modules/*.py) as parameter records, and
`Microgrid(modules, ...)`'s folding of them (microgrid/microgrid.py:100-165).  CPU only: the records are compared with
the hand-written parameter sets of the custom golden grids, and the C oracle run on them must reproduce what the LIVE
reference produced for the same constructor arguments (tests/golden/custom.npz, make_golden.build_custom)."""
import dataclasses
import warnings

import numpy as np
import pytest

from oracle.oracle import OracleGrid
from pymgrid_b200 import modules as M
from pymgrid_b200.modules import params_from_modules
from tests.helpers import custom_modules, custom_params, fuzz_modules, fuzz_params, fuzz_spec, jump_to, state_from_oracle


def assert_same_params(a, b):
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        if f.name == "meta":
            continue
        if dataclasses.is_dataclass(x):
            for g in dataclasses.fields(x):
                u, v = getattr(x, g.name), getattr(y, g.name)
                assert np.array_equal(u, v) if isinstance(u, np.ndarray) else u == v, (f.name, g.name, u, v)
        elif isinstance(x, np.ndarray):
            np.testing.assert_array_equal(x, y, err_msg=f.name)
        else:
            assert x == y, (f.name, x, y)


@pytest.mark.parametrize("i", range(6))
def test_modules_fold_into_the_same_record_and_reproduce_the_reference(golden, i):
    z = golden["custom"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # allow_abortion=False warns, as in the reference
        p = params_from_modules(custom_modules(z, i), loss_load_cost=9.0, overgeneration_cost=1.5)
    assert_same_params(p, custom_params(z, i))
    o = OracleGrid(p)
    np.testing.assert_array_equal(o.reset(), z[f"c{i}_reset_obs"])
    for k in range(len(z[f"c{i}_a"])):
        ob, r, d, info, _ = o.run(z[f"c{i}_a"][k])
        assert r == z[f"c{i}_r"][k] and d == bool(z[f"c{i}_d"][k])
        np.testing.assert_array_equal(ob, z[f"c{i}_o"][k])
        np.testing.assert_array_equal(info, z[f"c{i}_i"][k])
        np.testing.assert_array_equal(state_from_oracle(o), z[f"c{i}_s"][k])


@pytest.mark.parametrize("i", range(0, 40, 3))
def test_randomised_modules_fold_into_the_same_record(golden, i):
    """Randomised constructor arguments (tests/golden/fuzz.npz): the module classes fold into the record the parity tests
    build by hand -- including the soc the battery was constructed with, which the reference reports until the battery's
    first update (battery_module.py:89, 125-130) -- and the oracle run on it returns the reference's reset observation."""
    z = golden["fuzz"]
    s = fuzz_spec(z, i)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p = params_from_modules(fuzz_modules(z, i), loss_load_cost=s["llc"], overgeneration_cost=s["ogc"])
    assert p.battery.soc == s["b_init_soc"] and p.battery.reported_soc == z[f"f{i}_soc_before"][0] == z[f"f{i}_soc_before"][1]
    if s["initial_step"]:
        p = jump_to(p, int(s["initial_step"]))
    assert_same_params(p, fuzz_params(z, i))
    np.testing.assert_array_equal(OracleGrid(p).reset(), z[f"f{i}_reset_obs"])


def test_default_names_and_unbalanced_module_defaults(golden):
    """Microgrid(modules) defaults: loss_load_cost 10, overgeneration_cost 2 (microgrid.py:103-104); an un-named renewable is
    called 'renewable' (renewable_module.py:84); an explicit UnbalancedEnergyModule is honoured with add_unbalanced_module=False."""
    z = golden["custom"]
    p = params_from_modules(custom_modules(z, 4, renewable_name=None))
    assert (p.loss_load_cost, p.overgeneration_cost, p.renewable_name) == (10.0, 2.0, "renewable")
    assert p.unbalanced_name == "balancing"     # module_type[0] of the appended module (module_container.py:366-374)
    mods = custom_modules(z, 4) + [M.UnbalancedEnergyModule(raise_errors=False, loss_load_cost=3.0, overgeneration_cost=0.5)]
    p = params_from_modules(mods, add_unbalanced_module=False)
    assert (p.loss_load_cost, p.overgeneration_cost, p.renewable_name, p.unbalanced_name) == (3.0, 0.5, "pv", "balancing")
    named = custom_modules(z, 4) + [("unbalanced_energy", M.UnbalancedEnergyModule(raise_errors=False))]
    assert params_from_modules(named, add_unbalanced_module=False).unbalanced_name == "unbalanced_energy"
    with pytest.raises(NotImplementedError):
        params_from_modules(custom_modules(z, 4), add_unbalanced_module=False)       # no slack module at all
    with pytest.raises(NotImplementedError):
        params_from_modules(mods)                                                     # two slack modules


def test_forecaster_arguments(golden):
    """forecaster=None -> horizon 0 whatever forecast_horizon says (base_timeseries_module.py:42); a number -> Gaussian noise
    with the two flags (forecast/forecaster.py:10-89); callables are rejected."""
    z = golden["custom"]
    assert params_from_modules(custom_modules(z, 2)).forecast_horizon == 0
    load = M.LoadModule(z["load"], forecaster=12.5, forecast_horizon=6, forecaster_increase_uncertainty=True)
    pv = M.RenewableModule(z["pv"], forecaster="oracle", forecast_horizon=6)
    bat = M.BatteryModule(10, 100, 40, 45, 0.9, init_charge=60)
    p = params_from_modules([load, ("pv", pv), bat])
    assert p.forecast_horizon == 6 and set(p.forecasters) == {"load"}
    f = p.forecasters["load"]
    assert (f.noise_std, f.increase_uncertainty, f.relative_noise) == (12.5, True, False)
    assert p.battery.current_charge == 60
    with pytest.raises(NotImplementedError):
        params_from_modules([M.LoadModule(z["load"], forecaster=lambda a, b, n: b, forecast_horizon=6), ("pv", pv), bat])
    with pytest.raises(NotImplementedError):      # different horizons per module
        params_from_modules([M.LoadModule(z["load"], forecaster="oracle", forecast_horizon=5), ("pv", pv), bat])


def test_constructor_checks_match_the_reference():
    ts = np.ones((8, 3))
    with pytest.raises(AssertionError):
        M.BatteryModule(0, 100, 50, 50, 1.5, init_soc=0.5)                     # battery_module.py:78
    with pytest.raises(ValueError, match="Must set one of init_charge and init_soc"):
        M.BatteryModule(0, 100, 50, 50, 0.9)
    with pytest.warns(UserWarning, match="Using init_charge"):
        b = M.BatteryModule(0, 100, 50, 50, 0.9, init_charge=30, init_soc=0.9)
    assert (b.init_charge, b.init_soc) == (30, 0.3)
    with pytest.raises(ValueError, match="min_production must not be greater"):
        M.GensetModule(60, 50, 0.4)
    with pytest.warns(UserWarning, match="do not allow abortions"):
        M.GensetModule(10, 50, 0.4, allow_abortion=False)
    with pytest.raises(ValueError, match="max_import must be non-negative"):
        M.GridModule(-1, 0, ts)
    with pytest.raises(ValueError, match="three or four columns"):
        M.GridModule(1, 1, np.ones((8, 2)))
    with pytest.raises(ValueError, match="binary values"):
        M.GridModule(1, 1, np.full((8, 4), 0.5))
    with pytest.raises(ValueError, match="non-negative"):
        M.GridModule(1, 1, -ts)
    assert M.GridModule(1, 1, ts).time_series.shape == (8, 4) and (M.GridModule(1, 1, ts).time_series[:, 3] == 1).all()
    with pytest.raises(ValueError, match="both positive and negative"):
        M.LoadModule(np.array([1.0, -1.0, 2.0]))
    assert (M.LoadModule(np.array([1.0, 2.0])).time_series <= 0).all()          # sinks are stored negative
    assert (M.RenewableModule(np.array([-1.0, -2.0])).time_series >= 0).all()   # sources positive
    with pytest.raises(NotImplementedError):
        M.GensetModule(10, 50, lambda x: 0.4 * x)
    with pytest.raises(NotImplementedError):
        M.BatteryModule(0, 100, 50, 50, 0.9, init_soc=0.5, battery_transition_model=lambda **kw: 0.0)


def test_module_sets_outside_the_fused_step_fail_loudly(golden):
    z = golden["custom"]
    mods = custom_modules(z, 0)
    bat = M.BatteryModule(10, 100, 40, 45, 0.9, init_soc=0.6)
    for bad in (mods + [bat],                                             # two batteries
                [m for m in mods if not isinstance(m, M.BatteryModule)],   # none
                mods + [M.LoadModule(z["load"], forecaster="oracle", forecast_horizon=4)]):
        with pytest.raises(NotImplementedError, match="exactly one load"):
            params_from_modules(bad)
    with pytest.raises(NotImplementedError, match="renewable module name"):
        params_from_modules(custom_modules(z, 0, renewable_name="d_pv"))
    with pytest.raises(NotImplementedError, match="only the renewable and the slack module can be renamed"):
        params_from_modules([("storage", m) if isinstance(m, M.BatteryModule) else m for m in mods])
    with pytest.raises(TypeError):
        params_from_modules("battery")
    with pytest.raises(TypeError):
        params_from_modules([object()])


@pytest.mark.reference
def test_live_reference_builds_the_same_grid_from_the_same_arguments(golden):
    """Build container only: the reference's own constructors and ours, same keyword arguments -> same parameters, state
    and time series (read back from the live reference objects)."""
    from oracle.from_reference import params_from_reference
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid
    import pymgrid.modules as R
    z = golden["custom"]
    for i in range(6):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = pymgrid.Microgrid(custom_modules(z, i, ns=R), loss_load_cost=9.0, overgeneration_cost=1.5)
            ours = params_from_modules(custom_modules(z, i), loss_load_cost=9.0, overgeneration_cost=1.5)
        want = params_from_reference(ref)
        for name in ("loss_load_cost", "overgeneration_cost", "forecast_horizon", "initial_step", "final_step", "current_step"):
            assert getattr(ours, name) == getattr(want, name), (i, name)
        np.testing.assert_array_equal(ours.load_ts, want.load_ts)
        np.testing.assert_array_equal(ours.pv_ts, want.pv_ts)
        for part in ("battery", "genset", "grid"):
            a, b = getattr(ours, part), getattr(want, part)
            assert (a is None) == (b is None)
            if a is not None:
                for k, v in vars(b).items():
                    u = getattr(a, k)
                    assert np.array_equal(u, v) if isinstance(v, np.ndarray) else u == v, (i, part, k, u, v)
