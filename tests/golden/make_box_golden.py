"""Golden vectors for the 1-D x 2-D exchange-grid generators (create_xgrid_1dx2d_order1/2, create_xgrid_2dx1d_order1/2,
create_xgrid.c:208-598) from the UNMODIFIED reference (oracle/_ref).  Build container only.

    python tests/golden/make_box_golden.py      ->  tests/golden/xgrid_box.npz
Cases: a global box grid against one cubed-sphere tile with a pole inside (fix_lon twin poles), against an equatorial tile
straddling the date line, against a regional 2-D lat-lon grid with a random mask, a one-column box grid (the
get_grid_area_no_adjust branch of 1dx2d_order1), and a tripolar cap."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import xgtest  # noqa: E402

R = xgtest.ref_lib()
assert R is not None
D2R = np.pi / 180


def run(fn, nxi, nyi, nxo, nyo, lon_in, lat_in, lon_out, lat_out, mask, order):
    f = getattr(R, fn)
    f.restype = C.c_int
    cap = 400000
    bi = [np.zeros(cap, np.int32) for _ in range(4)]
    xa = np.zeros(cap); xc = np.zeros(cap); yc = np.zeros(cap)
    ci = lambda v: C.byref(C.c_int(v))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    a = [np.ascontiguousarray(v, np.float64) for v in (lon_in, lat_in, lon_out, lat_out, mask)]
    args = [ci(nxi), ci(nyi), ci(nxo), ci(nyo)] + [pv(v) for v in a] + [pv(v) for v in bi] + [pv(xa)]
    if order == 2:
        args += [pv(xc), pv(yc)]
    n = f(*args)
    out = {"n": np.int32(n), "idx": np.stack([v[:n] for v in bi]), "area": xa[:n].copy()}
    if order == 2:
        out["clon"] = xc[:n].copy(); out["clat"] = yc[:n].copy()
    return out


def main():
    rng = np.random.default_rng(7)
    lonc, latc = xgtest.ref_cubed_sphere(12)
    tl, ta = xgtest.tripolar_grid(48, 36)
    cases = {}
    boxes = {"g36x18": (np.linspace(0, 360, 37) * D2R, np.linspace(-90, 90, 19) * D2R),
             "g1x12": (np.array([0.0, 360.0]) * D2R, np.linspace(-90, 90, 13) * D2R),
             "r20x10": (np.linspace(100, 160, 21) * D2R, np.linspace(10, 60, 11) * D2R)}
    reg_lon, reg_lat = np.meshgrid(np.linspace(95, 170, 31) * D2R, np.linspace(5, 65, 25) * D2R)
    grids2d = {"c12t2": (lonc[2], latc[2]), "c12t0": (lonc[0], latc[0]), "c12t3": (lonc[3], latc[3]), "reg30x24": (reg_lon, reg_lat), "tripolar": (tl, ta)}
    combos = [("g36x18", "c12t2"), ("g36x18", "c12t0"), ("g36x18", "c12t3"), ("r20x10", "reg30x24"), ("g1x12", "c12t0"), ("g36x18", "tripolar")]
    out = {}
    for bname, (lb, ab) in boxes.items():
        out[f"box_{bname}_lon"] = lb; out[f"box_{bname}_lat"] = ab
    for gname, (lg, ag) in grids2d.items():
        out[f"grid_{gname}_lon"] = np.ascontiguousarray(lg); out[f"grid_{gname}_lat"] = np.ascontiguousarray(ag)
    names = []
    for bname, gname in combos:
        lb, ab = boxes[bname]; lg, ag = grids2d[gname]
        nxb, nyb = lb.size - 1, ab.size - 1
        nyc, nxc = lg.shape[0] - 1, lg.shape[1] - 1
        mask_box = (rng.uniform(size=nxb * nyb) > 0.15).astype(np.float64)
        mask_cell = (rng.uniform(size=nxc * nyc) > 0.15).astype(np.float64)
        key = f"{bname}__{gname}"
        names.append(key)
        out[key + "_mask_box"] = mask_box; out[key + "_mask_cell"] = mask_cell
        for order in (1, 2):
            r = run(f"create_xgrid_1dx2d_order{order}", nxb, nyb, nxc, nyc, lb, ab, lg, ag, mask_box, order)
            for k, v in r.items():
                out[f"{key}_1dx2d_o{order}_{k}"] = v
            r2 = run(f"create_xgrid_2dx1d_order{order}", nxc, nyc, nxb, nyb, lg, ag, lb, ab, mask_cell, order)
            for k, v in r2.items():
                out[f"{key}_2dx1d_o{order}_{k}"] = v
            print(key, "order", order, "1dx2d", int(r["n"]), "2dx1d", int(r2["n"]))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "xgrid_box.npz"), **out)


if __name__ == "__main__":
    main()
