#!/usr/bin/env python
"""Developer tool: run the windows that rank R of a WORLD-GPU job would own (bench.py's dealing: WORLD*8 windows of equal
candidate-pair count, round-robin) on ONE GPU and print the phase times — for looking at the polar ranks' extra work with
ncu without paying for an 8-GPU box.   python scripts/profile_rank.py R WORLD [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as g  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
pkg = g.load_package()
lonc, latc = pkg.cubed_sphere_grid(768)
lon2, lat2 = pkg.latlon_grid(2880, 1440)
plan = pkg.XgridPlan(0)
plan.set_dst(lon2, lat2)
plan.set_src(lonc, latc)
WPR = int(os.environ.get('XGB_WPR', '8'))
bounds = plan.partition(world * WPR)
wins = [(bounds[w], bounds[w + 1]) for w in range(rank, world * WPR, world)]
plan.set_src_windows(wins)
for _ in range(3):
    n = plan.generate(pkg.CONSERVE_ORDER2)
plan.reset_phase_ms()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(plan.stream, device=torch.device("cuda", 0))
with torch.cuda.stream(st):
    e0.record()
    for _ in range(steps):
        plan.generate(pkg.CONSERVE_ORDER2)
    e1.record()
torch.cuda.synchronize()
_, ph, ngen = plan.phase_ms()
import hashlib  # noqa: E402
res = plan.result_host()
h = hashlib.md5()
for k in sorted(res):
    h.update(res[k].tobytes())
print(json.dumps({"rank": rank, "world": world, "wpr": WPR, "nxgrid": int(n), "ms_per_step": e0.elapsed_time(e1) / steps, "md5": h.hexdigest(),
                  "heavy_stage": os.environ.get("XGB_HEAVY_STAGE", "1"),
                  "phase_ms": {k: round(v / max(ngen, 1), 4) for k, v in ph.items()}}))
