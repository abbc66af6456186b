#!/usr/bin/env python
"""Developer tool: great-circle weight generation of BASELINE configs[2] (1/4 degree tripolar -> 1 degree lat-lon) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import xgtest  # noqa: E402

pkg = xgtest.package()
tl, ta = xgtest.tripolar_grid(2880, 2160)
lon2, lat2 = xgtest.latlon_grid_np(360, 180)
plan = pkg.XgridPlan(0)
plan.set_dst(lon2, lat2); plan.set_src([tl], [ta])
op = pkg.CONSERVE_ORDER1 | pkg.GREAT_CIRCLE
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    n = plan.generate(op)
plan.sync()
import hashlib  # noqa: E402
res = plan.result_host()
h = hashlib.md5()
for k in sorted(res):
    h.update(res[k].tobytes())
print("nxgrid", n, "pairs", plan.npairs, "lib", os.environ.get("XGRID_B200_LIB", ""), "md5", h.hexdigest(),
      {k: round(v, 3) for k, v in plan.phase_ms()[0].items()})
