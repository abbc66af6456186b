"""Run the apply leg of bench.py alone (configs[1]: C96 -> 1440x720 order 2, 396 field-levels), for ncu captures:
    ncu --set full -k regex:apply_kernel\\|grad_c2l -o gpurun_out/apply python scripts/profile_apply.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

pkg = ge.load_package()
torch.cuda.set_device(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
r = bench.apply_leg(pkg, torch, None, 0, 1, 0, steps, 1)
print({k: r[k] for k in ("value", "ms_per_step", "field_levels_per_sec")}, r["e2e"])
