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
                            battery_cost_cycle=0.05, current_charge=0.6 * 100, soc=0.6)
    return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=z["load"], pv_ts=z["pv"],
                           loss_load_cost=9.0, overgeneration_cost=1.5, forecast_horizon=int(H),
                           final_step=int(final_step), unbalanced_name="balancing")


def fuzz_spec(z, i):
    return dict(zip([str(c) for c in z["spec_cols"]], z[f"f{i}_spec"]))


def fuzz_params(z, i):
    """Rebuild randomised grid i of tests/golden/fuzz.npz (see tests/golden/make_fuzz.py: draw_spec / build)."""
    s = fuzz_spec(z, i)
    genset = grid = None
    if s["has_gen"]:
        genset = GensetParams.with_init(init_start_up=bool(s["g_init"]), running_min_production=s["g_min"],
                                        running_max_production=s["g_max"], genset_cost=s["g_cost"], co2_per_unit=s["g_co2"],
                                        cost_per_unit_co2=s["g_cco2"], start_up_time=int(s["g_U"]),
                                        wind_down_time=int(s["g_D"]), allow_abortion=bool(s["g_abort"]))
    if s["has_grid"]:
        grid = GridParams(max_import=s["r_imp"], max_export=s["r_exp"], time_series=z[f"f{i}_grid_ts"],
                          cost_per_unit_co2=s["r_cco2"])
    battery = BatteryParams(min_capacity=s["b_min"], max_capacity=s["b_max"], max_charge=s["b_charge"],
                            max_discharge=s["b_discharge"], efficiency=s["b_eff"], battery_cost_cycle=s["b_cost"],
                            current_charge=s["b_init_soc"] * s["b_max"], soc=s["b_init_soc"])
    return MicrogridParams(battery=battery, genset=genset, grid=grid, load_ts=z[f"f{i}_load"], pv_ts=z[f"f{i}_pv"],
                           loss_load_cost=s["llc"], overgeneration_cost=s["ogc"], forecast_horizon=int(s["H"]),
                           initial_step=int(s["initial_step"]), current_step=int(s["initial_step"]),
                           final_step=int(s["final_step"]), unbalanced_name="balancing")


def fuzz_modules(z, i, ns=None):
    """Grid i of tests/golden/fuzz.npz as a list of reference-style MODULES, with the keyword arguments make_fuzz.build hands
    to the reference's constructors (initial_step is applied afterwards through `Microgrid.initial_step`, as there)."""
    if ns is None:
        from pymgrid_b200 import modules as ns
    s = fuzz_spec(z, i)
    ts_kw = dict(forecaster="oracle" if s["H"] > 0 else None, forecast_horizon=int(s["H"]) if s["H"] > 0 else 23,
                 final_step=int(s["final_step"]))
    mods = [ns.LoadModule(time_series=z[f"f{i}_load"], **ts_kw), ("pv", ns.RenewableModule(time_series=z[f"f{i}_pv"], **ts_kw))]
    if s["has_gen"]:
        mods.append(ns.GensetModule(running_min_production=s["g_min"], running_max_production=s["g_max"],
                                    genset_cost=s["g_cost"], co2_per_unit=s["g_co2"], cost_per_unit_co2=s["g_cco2"],
                                    start_up_time=int(s["g_U"]), wind_down_time=int(s["g_D"]),
                                    allow_abortion=bool(s["g_abort"]), init_start_up=bool(s["g_init"])))
    mods.append(ns.BatteryModule(min_capacity=s["b_min"], max_capacity=s["b_max"], max_charge=s["b_charge"],
                                 max_discharge=s["b_discharge"], efficiency=s["b_eff"], battery_cost_cycle=s["b_cost"],
                                 init_soc=s["b_init_soc"]))
    if s["has_grid"]:
        mods.append(ns.GridModule(max_import=s["r_imp"], max_export=s["r_exp"], time_series=z[f"f{i}_grid_ts"],
                                  cost_per_unit_co2=s["r_cco2"], **ts_kw))
    return mods


def overfull_params(z):
    """tests/golden/fuzz.npz `over_*`: a battery whose charge sits one ulp above max_capacity (make_fuzz.overfull_battery)."""
    mn, mx, ch, dis, eff, cc, imp, exp = z["over_spec"]
    battery = BatteryParams(min_capacity=mn, max_capacity=mx, max_charge=ch, max_discharge=dis, efficiency=eff,
                            battery_cost_cycle=cc, current_charge=float(z["over_charge"]))
    grid = GridParams(max_import=imp, max_export=exp, time_series=z["over_grid_ts"])
    return MicrogridParams(battery=battery, grid=grid, load_ts=z["over_load"], pv_ts=z["over_pv"], loss_load_cost=10.0,
                           overgeneration_cost=1.0, forecast_horizon=3)


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


# ---- numpy restatement of the engine's forecast-noise generator (include/pymgrid_b200.h, mg_forecast_noise) ----------
def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key (k0, k1) -> [..., 4] uint32.  Philox4x32-10, Salmon et al. SC'11."""
    c = np.array(counter, dtype=np.uint64) & 0xffffffff
    k0, k1 = np.uint64(key[0] & 0xffffffff), np.uint64(key[1] & 0xffffffff)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xffffffff)
    for _ in range(10):
        p0, p1 = M0 * c[..., 0], M1 * c[..., 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[..., 1] ^ k0, lo1, hi0 ^ c[..., 3] ^ k1, lo0], axis=-1)
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def engine_noise_normals(env, t, n_pairs, seed, call):
    """the standard normals the kernel draws for env id `env` at step `t`: [2 * n_pairs] (element f uses entry f)"""
    p = np.arange(n_pairs, dtype=np.uint64)
    ctr = np.stack([np.full(n_pairs, env & 0xffffffff, dtype=np.uint64), ((env >> 32) & 0xffff) | (p << np.uint64(16)),
                    np.full(n_pairs, t & 0xffffffff, dtype=np.uint64), np.full(n_pairs, call & 0xffffffff, dtype=np.uint64)], axis=-1)
    x = philox4x32_10(ctr, (seed & 0xffffffff, ((seed >> 32) ^ (call >> 32)) & 0xffffffff)).astype(np.float64)
    x = x.astype(np.uint64)
    u1 = ((x[:, 0] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (x[:, 1] >> np.uint64(6)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = ((x[:, 2] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (x[:, 3] >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    r = np.sqrt(-2.0 * np.log(u1))
    return np.stack([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)], axis=-1).reshape(-1)


def engine_noisy_row(clean, p, order, sigma, increase, env, t, seed, call, series_len):
    """Expected observation row after mg_forecast_noise.  clean: the oracle-forecast row; sigma: dict name -> per-column
    normalised std (list); increase: dict name -> bool."""
    from pymgrid_b200 import views
    H = p.forecast_horizon
    n_fc = H * (2 + 4 * p.has_grid)
    z = engine_noise_normals(env, t, (n_fc + 1) // 2, seed, call)
    n_real = min(max(series_len - (t + 1), 0), H)
    out = np.array(clean, dtype=np.float64)
    sl = views.obs_slices(p, order)
    for f in range(n_fc):
        if f < H:
            name, k, col, off = "load", f, 0, sl["load"].start + 1 + f
        elif f < 2 * H:
            name, k, col, off = "pv", f - H, 0, sl["pv"].start + 1 + (f - H)
        else:
            q = f - 2 * H
            name, k, col, off = "grid", q // 4, q % 4, sl["grid"].start + 4 + q
        s = sigma[name][col]
        if k >= n_real or s == 0:
            continue
        if increase[name]:
            s = s * (1.0 + np.log(1.0 + k))
        out[off] = min(max(out[off] + z[f] * s, 0.0), 1.0)
    return out


def custom_modules(golden_custom, i, ns=None, renewable_name="pv"):
    """Custom grid i of tests/golden/custom.npz as a list of reference-style MODULES, built with the very keyword arguments
    make_golden.build_custom hands to the reference's constructors.  `ns`: the namespace providing the classes
    (pymgrid_b200.modules by default; pymgrid.modules for a live cross-check)."""
    if ns is None:
        from pymgrid_b200 import modules as ns
    z = golden_custom
    U, D, abort, init, H, final_step, eff, weak, has_gen, has_grid = z[f"c{i}_spec"]
    fc = "oracle" if H > 0 else None
    Hc = int(H) if H > 0 else 23
    pv = ns.RenewableModule(time_series=z["pv"], forecaster=fc, forecast_horizon=Hc, final_step=int(final_step))
    mods = [ns.LoadModule(time_series=z["load"], forecaster=fc, forecast_horizon=Hc, final_step=int(final_step)),
            (renewable_name, pv) if renewable_name else pv]
    if has_gen:
        mods.append(ns.GensetModule(running_min_production=10, running_max_production=50, genset_cost=0.5, co2_per_unit=1.5,
                                    cost_per_unit_co2=0.2, start_up_time=int(U), wind_down_time=int(D),
                                    allow_abortion=bool(abort), init_start_up=bool(init)))
    mods.append(ns.BatteryModule(min_capacity=10, max_capacity=100, max_charge=40, max_discharge=45, efficiency=float(eff),
                                 battery_cost_cycle=0.05, init_soc=0.6))
    if has_grid:
        ts = z["grid_ts"].copy()
        if not weak:
            ts[:, 3] = 1.0
        mods.append(ns.GridModule(max_import=70, max_export=30, time_series=ts, forecaster=fc, forecast_horizon=Hc,
                                  final_step=int(final_step), cost_per_unit_co2=0.15))
    return mods
