#!/usr/bin/env python
"""Summarise an `ncu --set full` report into profiles/<tag>_ncu_full_kernels.{txt,json}.

usage: python scripts/ncu_summary.py gpurun_out/<tag>_kernels.ncu-rep <tag> ["header line"]
One row per captured launch; durations are cold-cache / serialised (profiler): use them for shares only.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

COLS = [
    ("duration_ms", "gpu__time_duration.sum"),
    ("fp64_pipe_active_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_slots_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("warp_insts", "smsp__inst_executed.sum"),
    ("dram_read", "dram__bytes_read.sum"),
    ("dram_write", "dram__bytes_write.sum"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    header = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("xgb::", "")
        d = {"kernel": short}
        for key, metric in COLS:
            if metric in hdr:
                i = hdr.index(metric)
                v = r[i].replace(",", "")
                try:
                    d[key] = float(v)
                except ValueError:
                    d[key] = v
                if key in ("dram_read", "dram_write"):
                    d[key + "_unit"] = units[i]
                if key == "duration_ms" and units[i] == "us":
                    d[key] = d[key] / 1000.0
                if key == "duration_ms" and units[i] == "ns":
                    d[key] = d[key] / 1e6
        out.append(d)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = os.path.join(root, "profiles", f"{tag}_ncu_full_kernels")
    json.dump(out, open(base + ".json", "w"), indent=1)
    keys = ["kernel"] + [k for k, _ in COLS]
    with open(base + ".txt", "w") as f:
        if header:
            f.write(header + "\n")
        f.write("one row per captured launch; durations are cold-cache/serialised (profiler), use them for shares only\n\n")
        f.write(" | ".join(keys) + "\n")
        for d in out:
            f.write(" | ".join(("%.4g" % d[k]) if isinstance(d.get(k), float) else str(d.get(k, "")) for k in keys)
                    + " (dram_read in %s, dram_write in %s)\n" % (d.get("dram_read_unit", "?"), d.get("dram_write_unit", "?")))
    print(open(base + ".txt").read())


if __name__ == "__main__":
    main()
