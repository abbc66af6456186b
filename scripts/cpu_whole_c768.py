#!/usr/bin/env python
"""ONE whole-problem CPU run of the headline configuration with the UNMODIFIED reference (oracle/_ref): C768 -> 2880x1440,
conserve_order2 exchange-grid generation, every destination row, all host cores — the fregrid_parallel-equivalent
(destination row bands, one process per band at a time, fregrid_util.c:592-603 layout {1,npes}).

Writes  profiles/r02_cpu_whole_c768.json   wall time, cores, xcells/s (the measured CPU baseline, no extrapolation)
        tests/golden/c768_rowhash.npz      per destination row: count and order-independent 64-bit sums over the row's
                                           exchange cells of (source cell, destination cell), xgrid_area bits, xgrid_clon
                                           bits, xgrid_clat bits.  tests/test_xgrid_gpu.py compares the GPU's whole list
                                           with these, row by row: the full-size list against the reference, all of it.

    python scripts/cpu_whole_c768.py [rows_per_band]      (build container: needs oracle/_ref; ~15-25 min on 8 cores)
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xgtest  # noqa: E402

N1, NLON, NLAT, ORDER = 768, 2880, 1440, 2
_G = {}


def job(band):
    jsc, jec = band
    lonc, latc, lon2, lat2 = _G["grids"]
    t0 = time.perf_counter()
    x = xgtest.ref_band_xgrid(lonc, latc, lon2, lat2, ORDER, jsc, jec)
    return xgtest.row_hashes(x, NLAT, N1, NLON), x["nxgrid"], time.perf_counter() - t0


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    assert xgtest.ref_lib() is not None, "oracle/_ref not built"
    lonc, latc = xgtest.ref_cubed_sphere(N1)
    lon2, lat2 = xgtest.latlon_grid_np(NLON, NLAT)       # get_output_grid_by_size (fregrid_util.c:588-603), no product library
    _G["grids"] = (lonc, latc, lon2, lat2)
    bands = [(j, min(j + rows, NLAT) - 1) for j in range(0, NLAT, rows)]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    tot = np.zeros((5, NLAT), np.uint64)
    nx = 0
    cpu_s = 0.0
    with mp.get_context("fork").Pool(cores) as pool:
        for h, n, dt in pool.imap_unordered(job, bands, chunksize=1):
            with np.errstate(over="ignore"):
                tot += h
            nx += n; cpu_s += dt
    wall = time.perf_counter() - t0
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c768_rowhash.npz"), count=tot[0], key=tot[1], area=tot[2], clon=tot[3], clat=tot[4])
    rec = {"workload": "C768 gnomonic_ed cubed sphere -> 2880x1440 lat-lon, conserve_order2 exchange-grid generation, WHOLE problem",
           "impl": "unmodified reference create_xgrid_2dx2d_order2 (oracle/_ref), one process per destination band of %d rows with "
                   "setup_conserve_interp's latitude trim, %d bands over %d host cores (fregrid_parallel-equivalent)" % (rows, len(bands), cores),
           "nxgrid": int(nx), "wall_s": wall, "cpu_core_s": cpu_s, "cores": cores, "xcells_per_s": nx / wall,
           "xcells_per_s_per_core": nx / cpu_s, "host": os.uname().nodename, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "r02_cpu_whole_c768.json"), "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
