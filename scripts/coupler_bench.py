"""make_coupler_mosaic exchange grids: the unmodified reference tool (oracle/_ref/make_coupler_mosaic_ref, one host core) and
xgb_make_coupler_xgrid on one GPU, on the same synthetic mosaics; prints one JSON line.  Usage:
  python scripts/coupler_bench.py [--atm 48] [--ocn 360x200] [--order 2] [--own-land 0|N]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atm", type=int, default=48)
    ap.add_argument("--ocn", default="360x200")
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--own-land", type=int, default=0, help="C<N> land mosaic of its own (0: land on the atmosphere mosaic)")
    ap.add_argument("--no-ref", action="store_true")
    a = ap.parse_args()
    pkg = ge.load_package()
    import test_coupler_gpu as T
    nxo, nyo = (int(v) for v in a.ocn.split("x"))
    d = tempfile.mkdtemp(prefix="cplbench")
    atm = T._write_mosaic(pkg, d, a.atm)
    lnd = T._write_mosaic(pkg, d, a.own_land) if a.own_land else None
    ocn = T._write_ocean(d, nxo, nyo, -80.0, seed=1, frac_land=0.3, area_frac=True)
    land_file = f"C{a.own_land}_mosaic.nc" if a.own_land else f"C{a.atm}_mosaic.nc"
    out = dict(atm=f"C{a.atm}", ocn=a.ocn, land=(f"C{a.own_land}" if a.own_land else "atmosphere mosaic"), order=a.order)
    if not a.no_ref:
        t0 = time.monotonic()
        r = subprocess.run([T._ref_tool(), "--atmos_mosaic", f"C{a.atm}_mosaic.nc", "--land_mosaic", land_file, "--ocean_mosaic",
                            "ocean_mosaic.nc", "--ocean_topog", "topog.nc", "--interp_order", str(a.order), "--mosaic_name", "grid_spec"],
                           cwd=d, capture_output=True, text=True)
        out["reference_tool_s"] = round(time.monotonic() - t0, 3)
        assert r.returncode == 0, r.stderr[-500:]
    args = (T._tiles(atm), [(ocn["lon"], ocn["lat"])], [ocn["omask"]])
    kw = dict(lnd=(T._tiles(lnd) if lnd else None), interp_order=a.order)
    times = []
    for _ in range(3):
        t0 = time.monotonic()
        x = pkg.make_coupler_xgrid(*args, **kw)
        times.append(round(time.monotonic() - t0, 4))
    out["gpu_call_s"] = times          # host arrays in, host lists out; the first call includes CUDA context creation
    out["cells"] = {k: int(x[k]["area"].size) for k in ("atmxlnd", "atmxocn", "lndxocn")}
    if not a.no_ref:
        name = f"C{a.atm}_mosaic"
        n = T._compare_lists(d, x["atmxocn"], name, "ocean_mosaic", a.order, ext2=1)
        n += T._compare_lists(d, x["atmxlnd"], name, (f"C{a.own_land}_mosaic" if a.own_land else name), a.order)
        if a.own_land:
            n += T._compare_lists(d, x["lndxocn"], f"C{a.own_land}_mosaic", "ocean_mosaic", a.order, ext2=1)
        out["cells_bit_identical_to_reference"] = n
        out["speedup_warm"] = round(out["reference_tool_s"] / min(times), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
