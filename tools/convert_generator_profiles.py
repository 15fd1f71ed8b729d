"""Bundle the raw load / PV / CO2 profiles MicrogridGenerator draws from (reference: src/pymgrid/data/{load,pv,co2}/*.csv,
read by MicrogridGenerator._get_random_file, MicrogridGenerator.py:119-135) into pymgrid_b200/data/generator_profiles.npz.
File order = sorted file names (recorded in the bundle).  Run in the build container."""
import glob
import os

import numpy as np

DATA_ROOT = os.environ.get("PYMGRID_DATA_ROOT", "/root/reference/src/pymgrid/data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pymgrid_b200", "data", "generator_profiles.npz")


def read(folder):
    files = sorted(glob.glob(os.path.join(DATA_ROOT, folder, "*.csv")))
    arrs = [np.genfromtxt(f, delimiter=",", skip_header=1, dtype=np.float64).reshape(-1) for f in files]
    assert all(len(a) == 8760 for a in arrs), [len(a) for a in arrs]
    return np.stack(arrs), np.array([os.path.basename(f) for f in files])


if __name__ == "__main__":
    load, load_names = read("load")
    pv, pv_names = read("pv")
    co2, co2_names = read("co2")
    np.savez_compressed(OUT, load=load, pv=pv, co2=co2, load_names=load_names, pv_names=pv_names, co2_names=co2_names)
    print(OUT, os.path.getsize(OUT) / 1e6, "MB", load.shape, pv.shape, co2.shape, load.max(axis=1), pv.max(axis=1))
