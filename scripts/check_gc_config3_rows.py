"""Offline check of the full-size great-circle case (BASELINE configs[2]: 1/4 degree tripolar -> 1 degree lat-lon):
the GPU's exchange cells for the 24 northernmost source rows (the bipolar cap, where the grid is least lat-lon like),
dumped by tests/test_gc_gpu.py with XGB_DUMP_GC_ROWS, against the CPU oracle run on the same rows.  Runs in the build
container (needs oracle/_ref for the tripolar grid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xgtest  # noqa: E402

pkg = xgtest.package()
g = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "gc_config3_toprows.npz"))
tl, ta = xgtest.tripolar_grid(2880, 2160)
ny = tl.shape[0] - 1
j0 = ny - 24
lon2, lat2 = pkg.latlon_grid(360, 180)
ref = xgtest.oracle_setup([tl[j0:]], [ta[j0:]], lon2, lat2, 1 | xgtest.GREAT_CIRCLE)
print("oracle nxgrid", ref["nxgrid"], "gpu", g["area"].size)
assert ref["nxgrid"] == g["area"].size
assert np.array_equal(ref["i_in"], g["i_in"]) and np.array_equal(ref["j_in"] + j0, g["j_in"])
assert np.array_equal(ref["i_out"], g["i_out"]) and np.array_equal(ref["j_out"], g["j_out"])
d = np.abs(ref["area"] - g["area"]) / 6371000.0 ** 2
print("lists identical; max |dA| = %.3g sr, bit-identical areas: %.1f %%" % (d.max(), 100 * np.mean(ref["area"] == g["area"])))
