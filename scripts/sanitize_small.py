#!/usr/bin/env python
"""Developer tool: a small pass over every kernel family, meant to run under compute-sanitizer on a B200:
  compute-sanitizer --tool memcheck python scripts/sanitize_small.py
  compute-sanitizer --tool racecheck python scripts/sanitize_small.py
Coarse source on a fine lat-lon destination (heavy cells, bitmap scatter, long finalize segments), a curvilinear destination
(pyramid walk + level-synchronous heavy path), several source windows, orders 1 and 2, generate_to_host, the apply / gradient
path and the great-circle generator."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402
import xgtest  # noqa: E402

pkg = g.load_package()
lonc, latc, lont, latt = pkg.cubed_sphere_grid(8, centers=True)
plan = pkg.XgridPlan(0)
for dst in (pkg.latlon_grid(720, 360), pkg.latlon_grid(60, 30, -180.0, 180.0, -80.0, 80.0)):
    plan.set_dst(*dst)
    plan.set_src(lonc, latc)
    for order in (pkg.CONSERVE_ORDER1, pkg.CONSERVE_ORDER2):
        n = plan.generate(order)
        print("lat-lon", dst[0].shape, "order", order, "nxgrid", n)
plan.set_dst_latlon(96, 48)
plan.set_src(lonc, latc)
b = plan.partition(4)
plan.set_src_windows([(b[0], b[1]), (b[2], b[3])])
print("windows", plan.generate(pkg.CONSERVE_ORDER2), plan.window_counts())
lond, latd = pkg.cubed_sphere_grid(20)
plan.set_dst(lond[2], latd[2])                       # curvilinear pole tile: pyramid + heavy expand
plan.set_src(lonc, latc)
print("curvilinear", plan.generate(pkg.CONSERVE_ORDER2))
# apply / gradient path
plan.set_dst_latlon(72, 36)
plan.set_src(lonc, latc)
n = plan.generate(pkg.CONSERVE_ORDER2)
plan.apply_setup()
hm = xgtest.cubed_sphere_halo_map(lonc, latc)
plan.grad_setup(xgtest.with_halo(lont.reshape(-1), hm), xgtest.with_halo(latt.reshape(-1), hm))
f = np.stack([xgtest.smooth_field(lont, latt, k, 0) for k in range(3)])
f[1, ::17] = -1e10
out = plan.regrid(pkg.CONSERVE_ORDER2, xgtest.with_halo(f, hm).reshape(-1), 3, has_missing=True, missing=-1e10)
print("regrid", float(np.asarray(out).reshape(3, -1)[0].mean()))
# great circle
n = plan.generate(pkg.CONSERVE_ORDER1 | pkg.GREAT_CIRCLE)
print("great circle", n)
plan.close()
print("done")
