"""Algorithmic FP64 work per exchange cell, from the op-counting build of the oracle.

TEST INFRASTRUCTURE.  Runs the oracle built with -DORC_COUNT_OPS (oracle/liboracle_xgrid_ops.so), which
counts how often each primitive step executes, and multiplies by the source-level operation count of
that step in the reference (each + - * / compare fabs = 1, each sin/cos = 1):

  step                       ops  reference lines
  inside_edge                  8  create_xgrid.c:2347-2348   4 sub 2 mul 1 add 1 cmp
  edge intersection           23  create_xgrid.c:1307-1319   10 sub 10 mul 2 div 1 fabs... (4+3+3+3+2+4+4)
  poly_area general edge      22  mosaic_util.c:423-448      incl. 2 sin, 1 div
  poly_area flat edge         17  mosaic_util.c:423-444      incl. 1 sin
  poly_area pole edge         12  mosaic_util.c:423-437
  poly_area epilogue           3  mosaic_util.c:455-458
  mask, min, ratio, test       4  create_xgrid.c:805-807
  poly_ctrlon edge            29  create_xgrid.c:2180-2204   incl. 4 trig, 1 div
  poly_ctrlat general edge    22  create_xgrid.c:2101-2117   incl. 4 trig, 1 div
  poly_ctrlat flat edge       19  create_xgrid.c:2101-2115   incl. 3 trig
  per-pair setup               9  create_xgrid.c:777-801     lat test 2, dx 1, 2 cmp, shift, lon test 2 (+2 adds rarely)

    python oracle/count_ops.py [ni nlon nlat]      (default C96 -> 1440x720, BASELINE config 2)
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xgtest  # noqa: E402

NAMES = ["clip_calls", "inside_edge", "intersect", "area_calls", "area_edge_gen", "area_edge_flat", "area_edge_pole",
         "ratio_tests", "accepted", "ctrlon_edge", "ctrlat_edge_gen", "ctrlat_edge_flat", "clip_nonempty"]
OPS = {"inside_edge": 8, "intersect": 23, "area_edge_gen": 22, "area_edge_flat": 17, "area_edge_pole": 12,
       "ctrlon_edge": 29, "ctrlat_edge_gen": 22, "ctrlat_edge_flat": 19}


def main():
    ni, nlon, nlat = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (96, 1440, 720)
    subprocess.run(["make", "-s", "-C", HERE, "ops"], check=True)
    L = C.CDLL(os.path.join(HERE, "liboracle_xgrid_ops.so"))
    dp, ip, vp = xgtest.dp, xgtest.ip, xgtest.vp
    L.orc_setup_conserve_interp.restype = C.c_long
    L.orc_setup_conserve_interp.argtypes = [C.c_int, ip, ip, dp, dp, C.c_int, C.c_int, dp, dp, C.c_uint, C.c_long] + [ip] * 5 + [dp, vp, vp]
    pkg = xgtest.package()
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    nx, ny, lon, lat = xgtest._tiles(lonc, latc)
    out = {}
    for order in (1, 2):
        cap = 12 * max(6 * ni * ni, nlon * nlat)
        bi = [np.zeros(cap, np.int32) for _ in range(5)]
        bd = [np.zeros(cap) for _ in range(3)]
        L.orc_counts_reset()
        # setup also evaluates get_grid_area for both grids: subtract those area calls (not per-xcell work)
        n = L.orc_setup_conserve_interp(6, nx, ny, lon, lat, nlon, nlat, lon2.ravel(), lat2.ravel(), order, cap, *bi, bd[0],
                                        bd[1].ctypes.data, bd[2].ctypes.data)
        cnt = (C.c_long * len(NAMES))()
        L.orc_counts_get(cnt)
        c = dict(zip(NAMES, [int(v) for v in cnt]))
        cell_area_calls = c["area_calls"] - c["clip_nonempty"]          # get_grid_area of both grids
        frac_cells = cell_area_calls / max(c["area_calls"], 1)
        # area edges belonging to grid cells (4 general-or-flat edges each) are removed proportionally
        area_ops = sum(c[k] * OPS[k] for k in ("area_edge_gen", "area_edge_flat", "area_edge_pole"))
        area_ops_x = area_ops * (1 - frac_cells) * (c["clip_nonempty"] * 1.0 / max(c["clip_nonempty"], 1))
        ops = (c["inside_edge"] * OPS["inside_edge"] + c["intersect"] * OPS["intersect"] + area_ops_x
               + c["clip_nonempty"] * 3 + c["ratio_tests"] * 4 + c["clip_calls"] * 9
               + c["ctrlon_edge"] * OPS["ctrlon_edge"] + c["ctrlat_edge_gen"] * OPS["ctrlat_edge_gen"]
               + c["ctrlat_edge_flat"] * OPS["ctrlat_edge_flat"] + (c["accepted"] * 4 if order == 2 else 0))
        out[f"order{order}"] = {"nxgrid": int(n), "clip_calls": c["clip_calls"], "counts": c,
                               "fp64_ops_total": ops, "fp64_ops_per_xcell": ops / n,
                               "fp64_ops_per_candidate_pair": ops / c["clip_calls"]}
        print(f"order {order}: nxgrid {n}, candidate pairs {c['clip_calls']}, algorithmic FP64 ops/xcell = {ops / n:.1f} "
              f"(per candidate pair {ops / c['clip_calls']:.1f}); inside_edge/pair {c['inside_edge'] / c['clip_calls']:.2f}, "
              f"intersections/pair {c['intersect'] / c['clip_calls']:.2f}")
    out["workload"] = f"C{ni} -> {nlon}x{nlat}"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"algorithmic_ops_c{ni}_{nlon}x{nlat}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
