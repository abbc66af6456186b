#!/usr/bin/env python
"""Developer tool: one batched order-2 regrid of BASELINE configs[1] (C96 -> 1440x720, 396 field-levels) for ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import xgtest  # noqa: E402

pkg = xgtest.package()
ni, nlon, nlat, B = 96, 1440, 720, (int(sys.argv[3]) if len(sys.argv) > 3 else 396)
lonc, latc, lont, latt = pkg.cubed_sphere_grid(ni, centers=True)
lon2, lat2 = pkg.latlon_grid(nlon, nlat)
hm = xgtest.cubed_sphere_halo_map(lonc, latc)
plan = pkg.XgridPlan(0)
plan.set_dst(lon2, lat2); plan.set_src(lonc, latc)
plan.generate(2); plan.apply_setup()
plan.grad_setup(xgtest.with_halo(lont.reshape(-1), hm), xgtest.with_halo(latt.reshape(-1), hm))
rng = np.random.default_rng(1)
f = rng.uniform(0, 1, (B, 6 * ni * ni))
d_in = torch.from_numpy(xgtest.with_halo(f, hm).reshape(-1)).cuda()
d_out = torch.empty(B * nlon * nlat, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    plan.regrid(2, d_in, B, out=d_out)
torch.cuda.synchronize(); plan.sync()
if len(sys.argv) > 2 and sys.argv[2] == "time":      # CUDA-event time of the whole regrid + a hash of the result (variant runs)
    import hashlib
    st = torch.cuda.ExternalStream(plan.stream, device=torch.device("cuda", 0))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(20):
            plan.regrid(2, d_in, B, out=d_out)
        e1.record()
    torch.cuda.synchronize(); plan.sync()
    print({"lib": os.environ.get("XGRID_B200_LIB", ""), "regrid_ms": round(e0.elapsed_time(e1) / 20, 4),
           "md5": hashlib.md5(d_out.cpu().numpy().tobytes()).hexdigest()})
print("ok")
