#!/usr/bin/env python
"""Developer tool: build kernel variants of libxgrid_b200 (compile-time -D switches) and compare them on a GPU.

  python scripts/clip_variants.py build 0 1 3 7          # here (nvcc cross-compiles): variants/libxgrid_b200_v<k>.so
  python scripts/clip_variants.py run 0 1 3 7            # on a B200: per variant, a digest of the full C768 result
                                                         # (must equal variant 0's, which the GPU test suite pins to the
                                                         # oracle) and the phase times of the device-resident step

A variant spec is "<clip variant bits>[:<extra define>...]", e.g. "7:XGB_CLIP_BLOCKS=6".
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "fre-nctools_b200", "variants")


def vpath(spec):
    return os.path.join(VDIR, "libxgrid_b200_v%s.so" % spec.replace(":", "_").replace("=", "-"))


def build(specs):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        parts = spec.split(":")
        defs = ["XGB_CLIP_VARIANT=" + parts[0]] + parts[1:]
        pkg._build.build(out=vpath(spec), defines=defs)
        print("built", vpath(spec))


def child(workload_n, nlon, nlat, order, steps):
    """runs in a subprocess with XGRID_B200_LIB set"""
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    lonc, latc = pkg.cubed_sphere_grid(workload_n)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    n = plan.generate(order)
    res = plan.result_host()
    h = hashlib.md5()
    for k in sorted(res):
        if hasattr(res[k], "tobytes"):
            h.update(k.encode()); h.update(np.ascontiguousarray(res[k]).tobytes())
    for _ in range(3):
        plan.generate(order)
    plan.reset_phase_ms()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(plan.stream, device=torch.device("cuda", 0))
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(steps):
            plan.generate(order)
        e1.record()
    torch.cuda.synchronize()
    _, ph, ngen = plan.phase_ms()
    print(json.dumps({"nxgrid": int(n), "md5": h.hexdigest(), "ms_per_step": e0.elapsed_time(e1) / steps,
                      "phase_ms": {k: v / max(ngen, 1) for k, v in ph.items()}}))


def run(specs, order=2):
    base = None
    for spec in specs:
        env = dict(os.environ, XGRID_B200_LIB=vpath(spec))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(order)], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(spec, "FAILED", r.stderr[-800:])
            continue
        d = json.loads(r.stdout.strip().split("\n")[-1])
        base = base or d["md5"]
        print("variant %-28s order %d  step %.3f ms  clip %.3f ms  nxgrid %d  %s" % (
            spec, order, d["ms_per_step"], d["phase_ms"].get("clip", 0.0), d["nxgrid"],
            "same result as first" if d["md5"] == base else "RESULT DIFFERS"), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        run(sys.argv[2:], 2)
        run(sys.argv[2:], 1)
    elif sys.argv[1] == "child":
        child(768, 2880, 1440, int(sys.argv[2]), 10)
