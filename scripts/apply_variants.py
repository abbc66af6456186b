#!/usr/bin/env python
"""Developer tool: time the apply leg of bench.py (C96 -> 1440x720 order 2, 396 field-levels) for library variants built by
scripts/clip_variants.py (XGRID_B200_LIB).   python scripts/apply_variants.py <spec> [<spec> ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

if sys.argv[1] == "child":
    import torch
    import bench
    import __graft_entry__ as g
    r = bench.apply_leg(g.load_package(), torch, None, 0, 1, 0, 10, 3)
    print(json.dumps({"ms": r["ms_per_step"], "gbs": r["value"]}))
else:
    import clip_variants
    for spec in sys.argv[1:]:
        env = dict(os.environ, XGRID_B200_LIB=clip_variants.vpath(spec))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print("variant %-44s %s" % (spec, r.stdout.strip().split("\n")[-1] if r.returncode == 0 else "FAILED " + r.stderr[-300:]), flush=True)
