"""Shared helpers for the parity tests: golden-layout utilities and custom-grid builders."""
import numpy as np

from pymgrid_b200.params import BatteryParams, GensetParams, GridParams, MicrogridParams


def state_from_oracle(o):
    s = o.state
    return np.array([s["t"], s["charge"], *s["genset"]], dtype=np.float64)


def custom_params(golden_custom, i):
    """Rebuild custom grid i of tests/golden/custom.npz (see make_golden.custom_grids / build_custom)."""
    z = golden_custom
    U, D, abort, init, H, final_step, eff, weak, has_gen, has_grid = z[f"c{i}_spec"]
    genset = grid = None
    if has_gen:
        genset = GensetParams.with_init(init_start_up=bool(init), running_min_production=10, running_max_production=50,
                                        genset_cost=0.5, co2_per_unit=1.5, cost_per_unit_co2=0.2, start_up_time=int(U),
                                        wind_down_time=int(D), allow_abortion=bool(abort))
    if has_grid:
        ts = z["grid_ts"].copy()
        if not weak:
            ts[:, 3] = 1.0
        grid = GridParams(max_import=70, max_export=30, time_series=ts, cost_per_unit_co2=0.15)
    battery = BatteryParams(min_capacity=10, max_capacity=100, max_charge=40, max_discharge=45, efficiency=float(eff),
                            battery_cost_cycle=0.05, current_charge=0.6 * 100)
    return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=z["load"], pv_ts=z["pv"],
                           loss_load_cost=9.0, overgeneration_cost=1.5, forecast_horizon=int(H),
                           final_step=int(final_step))


def jump_to(params, step):
    """Same effect as `m.initial_step = step; m.reset()` on the reference object."""
    import copy
    p = copy.copy(params)
    p.initial_step = step
    p.current_step = step
    return p


def generator_params(z, i):
    """Rebuild real MicrogridGenerator grid i of tests/golden/generator.npz in (profile, scale) form."""
    from pymgrid_b200 import generator
    pr = generator.load_profiles()
    rec = z[f"g{i}_rec"]
    (lp, lscale, pp, pscale, bmin, bmax, bch, bdis, beff, bcc, bcharge, has_gen, gmin, gmax, gcost, gco2, gcc,
     has_grid, gimp, gexp, grcc, tariff, cid, llc, ogc, H, final_step) = rec
    battery = BatteryParams(bmin, bmax, bch, bdis, beff, bcc, bcharge)
    genset = grid = None
    if has_gen:
        genset = GensetParams.with_init(running_min_production=gmin, running_max_production=gmax, genset_cost=gcost,
                                        co2_per_unit=gco2, cost_per_unit_co2=gcc)
    if has_grid:
        status = np.unpackbits(z[f"g{i}_status"])[:8760].astype(np.float64)
        ts = generator.grid_table(pr, int(tariff), int(cid))
        assert np.array_equal(ts[:, 0], z[f"g{i}_import_price"])
        grid = GridParams(max_import=gimp, max_export=gexp, time_series=ts, cost_per_unit_co2=grcc, status=status)
    return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=pr["load"][int(lp)], pv_ts=pr["pv"][int(pp)],
                           load_scale=float(lscale), pv_scale=float(pscale), loss_load_cost=llc, overgeneration_cost=ogc,
                           forecast_horizon=int(H), final_step=int(final_step), renewable_name="PV")
