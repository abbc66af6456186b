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
# round 2: regrid without missing values over 70 field-levels (two 64-level chunks of the record layout), box generators,
# the coupler's exchange grids (own land mosaic, partly wet ocean, artificial southern row), one-process multi-GPU entry
plan = pkg.XgridPlan(0)
plan.set_dst_latlon(72, 36); plan.set_src(lonc, latc)
plan.generate(pkg.CONSERVE_ORDER2); plan.apply_setup()
plan.grad_setup(xgtest.with_halo(lont.reshape(-1), hm), xgtest.with_halo(latt.reshape(-1), hm))
f = np.stack([xgtest.smooth_field(lont, latt, k, 0) for k in range(70)])
out = plan.regrid(pkg.CONSERVE_ORDER2, xgtest.with_halo(f, hm).reshape(-1), 70)
print("regrid 70", float(np.asarray(out).reshape(70, -1)[69].mean()))
plan.close()
lon1d = np.linspace(0.0, 2 * np.pi, 37); lat1d = np.linspace(-0.5 * np.pi, 0.5 * np.pi, 19)
print("1dx2d", pkg.create_xgrid_1dx2d_order2(lon1d, lat1d, lonc[0], latc[0])[0],
      "2dx1d", pkg.create_xgrid_2dx1d_order1(lonc[0], latc[0], lon1d, lat1d)[0])
lonl, latl = pkg.cubed_sphere_grid(6)
xo = np.repeat(np.linspace(-280.0, 80.0, 25)[None, :], 13, 0) * np.pi / 180.0
yo = np.repeat(np.linspace(-90.0, 90.0, 13)[:, None], 25, 1) * np.pi / 180.0
om = np.random.default_rng(0).choice([0.0, 1.0, 0.4], size=(12, 24))
x = pkg.make_coupler_xgrid([(lonc[t], latc[t]) for t in range(6)], [(xo, yo)], [om], lnd=[(lonl[t], latl[t]) for t in range(6)], interp_order=2)
print("coupler", {k: x[k]["area"].size for k in ("atmxlnd", "atmxocn", "lndxocn")})
x = pkg.make_coupler_xgrid([(lonc[t], latc[t]) for t in range(6)], [(xo, yo)], [om], lnd=None, interp_order=1)
print("coupler, land on the atmosphere mosaic", {k: x[k]["area"].size for k in ("atmxlnd", "atmxocn", "lndxocn")})
print("done")
