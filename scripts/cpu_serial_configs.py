#!/usr/bin/env python
"""Serial fregrid baseline: the UNMODIFIED reference's setup_conserve_interp (oracle/_ref), ONE process, whole
problem, for BASELINE configs[0] (C48 -> 360x180 order 1) and configs[1] weights (C96 -> 1440x720 order 2).
Writes profiles/r02_cpu_serial_configs.json.   python scripts/cpu_serial_configs.py   (build container, ~2 min)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xgtest  # noqa: E402

out = []
for name, n, nlon, nlat, order in (("configs[0]", 48, 360, 180, 1), ("configs[1] weights", 96, 1440, 720, 2)):
    lonc, latc = xgtest.ref_cubed_sphere(n)
    lon2, lat2 = xgtest.latlon_grid_np(nlon, nlat)
    t0 = time.perf_counter()
    r = xgtest.ref_setup(lonc, latc, lon2, lat2, order)
    dt = time.perf_counter() - t0
    out.append({"config": name, "workload": f"C{n} -> {nlon}x{nlat} conserve_order{order}, serial fregrid (one process, whole problem)",
                "nxgrid": r["nxgrid"], "seconds": dt, "xcells_per_s": r["nxgrid"] / dt, "cores": 1,
                "impl": "unmodified reference setup_conserve_interp (oracle/_ref)"})
    print(out[-1], flush=True)
json.dump({"host": os.uname().nodename, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "runs": out},
          open(os.path.join(ROOT, "profiles", "r02_cpu_serial_configs.json"), "w"), indent=1)
