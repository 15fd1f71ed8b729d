"""Convert the reference's pymgrid25 scenario files into the bundled `pymgrid_b200/data/pymgrid25.npz`.

Run in the build container (needs /root/reference/src/pymgrid/data).  Only our own reader
(`pymgrid_b200.scenario.read_reference_scenario`) touches the files; no reference code is imported.
Grid series are stored factored (price columns, CO2 profile id, packed status bits) when that
reproduces the original columns bit for bit -- verified below -- which keeps the bundle small.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pymgrid_b200.scenario import (N_PYMGRID25, load_pymgrid25, pack_scalars, read_reference_scenario,  # noqa: E402
                                   reference_scenario_path, _bundle)

DATA_ROOT = os.environ.get("PYMGRID_DATA_ROOT", "/root/reference/src/pymgrid/data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pymgrid_b200", "data", "pymgrid25.npz")


def main():
    grids = [read_reference_scenario(reference_scenario_path(n, DATA_ROOT)) for n in range(N_PYMGRID25)]
    T = len(grids[0])
    assert all(len(g) == T for g in grids)
    scalars = np.stack([pack_scalars(g) for g in grids])
    load = np.stack([g.load_ts for g in grids])
    pv = np.stack([g.pv_ts for g in grids])
    co2_profiles, grid_index = [], np.full(N_PYMGRID25, -1, dtype=np.int32)
    imp, exp, co2_id, status = [], [], [], []
    for n, g in enumerate(grids):
        if g.grid is None:
            continue
        ts = g.grid.time_series
        for k, prof in enumerate(co2_profiles):
            if np.array_equal(prof, ts[:, 2]):
                break
        else:
            co2_profiles.append(ts[:, 2].copy())
            k = len(co2_profiles) - 1
        grid_index[n] = len(imp)
        imp.append(ts[:, 0]); exp.append(ts[:, 1]); co2_id.append(k)
        status.append(np.packbits(ts[:, 3].astype(np.uint8)))
    np.savez_compressed(OUT, scalars=scalars, load=load, pv=pv, grid_index=grid_index,
                        grid_import_price=np.stack(imp), grid_export_price=np.stack(exp),
                        grid_co2_profile=np.array(co2_id, dtype=np.int32), co2_profiles=np.stack(co2_profiles),
                        grid_status_bits=np.stack(status))
    _bundle.cache_clear()
    for n, g in enumerate(grids):   # bit-exact round trip
        b = load_pymgrid25(n)
        assert np.array_equal(b.load_ts, g.load_ts) and np.array_equal(b.pv_ts, g.pv_ts)
        assert (b.grid is None) == (g.grid is None)
        if g.grid is not None:
            assert np.array_equal(b.grid.time_series, g.grid.time_series)
            assert (b.grid.max_import, b.grid.max_export, b.grid.cost_per_unit_co2) == \
                   (g.grid.max_import, g.grid.max_export, g.grid.cost_per_unit_co2)
        assert b.battery == g.battery and b.genset == g.genset
        assert (b.loss_load_cost, b.overgeneration_cost, b.forecast_horizon, b.initial_step, b.final_step) == \
               (g.loss_load_cost, g.overgeneration_cost, g.forecast_horizon, g.initial_step, g.final_step)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB, {len(co2_profiles)} CO2 profiles, {len(imp)} grids")


if __name__ == "__main__":
    main()
