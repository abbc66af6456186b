"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference): it compiles the reference sources into
oracle/_ref/libfrenc_ref.so (oracle/Makefile) and records its outputs, gcc 13.3 -O2, glibc 2.39,
x86-64 — the compiler and libm matter for the last bits of the floating-point fields.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import xgtest  # noqa: E402

R = xgtest.ref_lib()
assert R is not None, "reference not built"
D2R = np.pi / 180
MV = 50


def ref_fix_lon(x, y):
    xx = np.zeros(MV); yy = np.zeros(MV)
    xx[:4] = x; yy[:4] = y
    n = R.fix_lon(xx, yy, 4, np.pi)
    return n, xx[:10].copy(), yy[:10].copy()


def polygon_cases(rng):
    cases = []
    def quad(lon0, lat0, dlon, dlat, jitter):
        x = np.array([lon0, lon0 + dlon, lon0 + dlon, lon0]) + rng.normal(0, jitter, 4)
        y = np.array([lat0, lat0, lat0 + dlat, lat0 + dlat]) + rng.normal(0, jitter, 4)
        return x * D2R, np.clip(y, -90, 90) * D2R
    for _ in range(260):                                    # generic overlapping quads anywhere on the sphere
        lon0 = rng.uniform(0, 358); lat0 = rng.uniform(-88, 86); d = rng.uniform(0.2, 3.0)
        a = quad(lon0, lat0, d, d, 0.05 * d)
        b = quad(lon0 + rng.uniform(-d, d), lat0 + rng.uniform(-d, d), d * rng.uniform(0.5, 2), d * rng.uniform(0.5, 2), 0.0)
        cases.append((a, b))
    for _ in range(40):                                     # date-line / negative longitudes
        lon0 = rng.uniform(-3, 1); lat0 = rng.uniform(-60, 60)
        cases.append((quad(lon0, lat0, 2, 2, 0.05), quad(lon0 + 359, lat0 + 0.5, 2.5, 2, 0.0)))
    for _ in range(40):                                     # pole row of a lat-lon grid against a polar cap quad
        lon0 = rng.uniform(0, 350)
        b = (np.array([lon0, lon0 + 5, lon0 + 5, lon0]) * D2R, np.array([85, 85, 90, 90.]) * D2R)
        c = rng.uniform(0, 360)
        a = (np.array([c, c + 90, c + 180, c + 270]) * D2R % (2 * np.pi), np.array([88, 87.5, 88.2, 87.9]) * D2R)
        cases.append((a, b))
    for _ in range(20):                                     # a single pole vertex (fix_lon pairs it: 5 vertices)
        lon0 = rng.uniform(0, 340); s = rng.choice([-1.0, 1.0])
        a = (np.array([lon0, lon0 + 10, lon0 + 5, lon0 - 3]) * D2R, s * np.array([88, 88.2, 90, 89]) * D2R)
        b = quad(lon0 + 2, s * 88.5 - 0.5, 4, 1.2, 0.0)
        cases.append((a, b))
    for _ in range(20):                                     # a side running through the pole (twin pole vertices: 6)
        lon0 = rng.uniform(0, 170); s = rng.choice([-1.0, 1.0])
        a = (np.array([lon0, lon0 + 180, lon0 + 170, lon0 + 10]) * D2R, s * np.array([89, 89, 88, 88]) * D2R)
        b = quad(lon0 + 2, s * 88.6 - 0.4, 6, 0.8, 0.0)
        cases.append((a, b))
    for _ in range(20):                                     # identical and nested boxes
        lon0 = rng.uniform(10, 300); lat0 = rng.uniform(-70, 70)
        a = quad(lon0, lat0, 2, 2, 0.0)
        cases.append((a, a))
        cases.append((a, quad(lon0 + 0.5, lat0 + 0.5, 1, 1, 0.0)))
    return cases


def make_polys():
    rng = np.random.default_rng(20260101)
    cases = polygon_cases(rng)
    n = len(cases)
    out = {k: np.zeros((n, 10)) for k in ("fx1", "fy1", "fx2", "fy2")}
    out.update({"x1": np.zeros((n, 4)), "y1": np.zeros((n, 4)), "x2": np.zeros((n, 4)), "y2": np.zeros((n, 4)),
                "n1": np.zeros(n, np.int32), "n2": np.zeros(n, np.int32), "n_out": np.zeros(n, np.int32),
                "ox": np.zeros((n, 20)), "oy": np.zeros((n, 20)), "area1": np.zeros(n), "area2": np.zeros(n),
                "xarea": np.zeros(n), "ctrlon": np.zeros(n), "ctrlat": np.zeros(n)})
    for i, ((x1, y1), (x2, y2)) in enumerate(cases):
        out["x1"][i], out["y1"][i], out["x2"][i], out["y2"][i] = x1, y1, x2, y2
        n1, fx1, fy1 = ref_fix_lon(x1, y1)
        n2, fx2, fy2 = ref_fix_lon(x2, y2)
        out["n1"][i], out["n2"][i] = n1, n2
        out["fx1"][i], out["fy1"][i], out["fx2"][i], out["fy2"][i] = fx1, fy1, fx2, fy2
        a1 = np.zeros(MV); b1 = np.zeros(MV); a2 = np.zeros(MV); b2 = np.zeros(MV)
        a1[:10], b1[:10], a2[:10], b2[:10] = fx1, fy1, fx2, fy2
        out["area1"][i] = R.poly_area(a1, b1, n1)
        out["area2"][i] = R.poly_area(a2, b2, n2)
        # the generators shift polygon 2 by +-2pi towards polygon 1 before clipping (create_xgrid.c:786-796)
        dx = a2[:n2].mean() - a1[:n1].mean()
        if dx < -np.pi: a2[:n2] += 2 * np.pi
        elif dx > np.pi: a2[:n2] -= 2 * np.pi
        out["fx2"][i] = a2[:10]
        ox = np.zeros(MV); oy = np.zeros(MV)
        no = R.clip_2dx2d(a1, b1, n1, a2, b2, n2, ox, oy)
        out["n_out"][i] = no
        out["ox"][i], out["oy"][i] = ox[:20], oy[:20]
        if no > 0:
            out["xarea"][i] = R.poly_area(ox, oy, no)
            out["ctrlon"][i] = R.poly_ctrlon(ox, oy, no, a1[:n1].mean())
            out["ctrlat"][i] = R.poly_ctrlat(ox, oy, no)
    np.savez_compressed(os.path.join(HERE, "polys.npz"), **out)
    print("polys.npz:", n, "cases;", int((out["n_out"] > 0).sum()), "non-empty clips; n1 max", out["n1"].max())


def make_xgrid(tag, lonc, latc, lon2, lat2, opcode):
    r = xgtest.ref_setup(lonc, latc, lon2, lat2, opcode)
    nx, ny, lon, lat = xgtest._tiles(lonc, latc)
    d = {"nx": nx, "ny": ny, "lon_in": lon, "lat_in": lat, "lon_out": np.asarray(lon2), "lat_out": np.asarray(lat2),
         "opcode": np.int32(opcode)}
    d.update({k: v for k, v in r.items() if k != "nxgrid"})
    np.savez_compressed(os.path.join(HERE, f"xgrid_{tag}.npz"), **d)
    print(f"xgrid_{tag}.npz: nxgrid", r["nxgrid"])


def make_apply(tag, ni, nlon, nlat):
    """remapped fields of the reference's do_scalar_conserve_interp + grad_c2l + calc_c2l_grid_info on a small case:
    C<ni> -> nlon x nlat, two field-levels (smooth, random), order 1 / order 2 / order 2 with 5 % missing / monotonic"""
    lonc, latc, lont, latt = xgtest.ref_cubed_sphere(ni, centers=True)
    lo, la = latlon(nlon, nlat)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    xt = xgtest.with_halo(lont.reshape(-1), hm); yt = xgtest.with_halo(latt.reshape(-1), hm)
    nh = (ni + 2) ** 2; nc = ni * ni
    metrics = [xgtest.c2l_metrics("ref", ni, ni, xt[t * nh:(t + 1) * nh], yt[t * nh:(t + 1) * nh], lonc[t], latc[t]) for t in range(6)]
    rng = np.random.default_rng(1234)
    fields = np.stack([xgtest.smooth_field(lont, latt, 3, 2), rng.uniform(0, 1, 6 * nc)])
    miss = -999.0
    holes = np.random.default_rng(4321).uniform(size=fields.shape) < 0.05
    fields_m = np.where(holes, miss, fields)
    d = {"ni": np.int32(ni), "nlon": np.int32(nlon), "nlat": np.int32(nlat), "fields": fields, "fields_missing": fields_m,
         "missing": np.float64(miss), "xt": xt, "yt": yt}
    for k in xgtest.METRICS:
        d["m_" + k] = np.concatenate([m[k] for m in metrics])
    r1 = xgtest.ref_setup(lonc, latc, lo, la, 1, keep=True)
    d["out_o1"] = np.stack([xgtest.ref_apply(r1["handle"], 1, f, nlon * nlat) for f in fields])
    d["out_o1_missing"] = np.stack([xgtest.ref_apply(r1["handle"], 1, f, nlon * nlat, has_missing=True, missing=miss) for f in fields_m])
    r2 = xgtest.ref_setup(lonc, latc, lo, la, 2, keep=True)
    for name, src, hm_flag in (("", fields, False), ("_missing", fields_m, True)):
        outs, outs_mono, gxs, gys, gms = [], [], [], [], []
        for f in src:
            fh = xgtest.with_halo(f, hm, corner=0.0)
            gx = np.zeros(6 * nc); gy = np.zeros(6 * nc); gm = np.zeros(6 * nc, np.int32)
            for t in range(6):
                a, b = xgtest.grad_c2l("ref", ni, ni, fh[t * nh:(t + 1) * nh], metrics[t])
                gx[t * nc:(t + 1) * nc] = a; gy[t * nc:(t + 1) * nc] = b
                if hm_flag:
                    gm[t * nc:(t + 1) * nc] = xgtest.grad_mask(ni, ni, fh[t * nh:(t + 1) * nh], miss)
            outs.append(xgtest.ref_apply(r2["handle"], 2, fh, nlon * nlat, gx, gy, gm, has_missing=hm_flag, missing=miss))
            outs_mono.append(xgtest.ref_apply(r2["handle"], 2, fh, nlon * nlat, gx, gy, gm, has_missing=hm_flag, missing=miss, monotonic=True))
            gxs.append(gx); gys.append(gy); gms.append(gm)
        d["out_o2" + name] = np.stack(outs); d["out_o2_mono" + name] = np.stack(outs_mono)
        d["grad_x" + name] = np.stack(gxs); d["grad_y" + name] = np.stack(gys); d["grad_mask" + name] = np.stack(gms)
    np.savez_compressed(os.path.join(HERE, f"apply_{tag}.npz"), **d)
    print(f"apply_{tag}.npz: nxgrid", r1["nxgrid"], r2["nxgrid"])


def make_gc():
    """great-circle path: polygon pairs through the reference's clip_2dx2d_great_circle + great_circle_area, and whole
    exchange grids (cubed sphere and a small tripolar ocean grid -> lat-lon) through create_xgrid_great_circle"""
    cases = xgtest.gc_quad_cases(1000, seed=20260102)
    n = len(cases)
    d = {"p1": np.zeros((n, 3, 4)), "p2": np.zeros((n, 3, 4)), "n_out": np.zeros(n, np.int32), "out": np.zeros((n, 3, 10)),
         "area": np.zeros(n)}
    for k, (a, b) in enumerate(cases):
        d["p1"][k] = np.stack(a); d["p2"][k] = np.stack(b)
        o = [np.zeros(60) for _ in range(3)]
        m = R.clip_2dx2d_great_circle(*a, 4, *b, 4, *o)
        d["n_out"][k] = m
        for c in range(3):
            d["out"][k, c] = o[c][:10]
        if m > 0:
            d["area"][k] = R.great_circle_area(m, *o)
    np.savez_compressed(os.path.join(HERE, "gc_polys.npz"), **d)
    print("gc_polys.npz:", n, "cases;", int((d["n_out"] > 0).sum()), "non-empty; max n_out", d["n_out"].max())
    GC = xgtest.GREAT_CIRCLE
    c8 = xgtest.ref_cubed_sphere(8)
    lo, la = latlon(36, 18)
    make_xgrid("gc_c8_36x18", c8[0], c8[1], lo, la, 1 | GC)
    tl, ta = xgtest.tripolar_grid(48, 36)
    make_xgrid("gc_tripolar24x18_36x18", [tl], [ta], lo, la, 1 | GC)
    a = np.zeros(24 * 18)
    R.get_grid_great_circle_area(C.byref(C.c_int(24)), C.byref(C.c_int(18)), np.ascontiguousarray(tl).reshape(-1), np.ascontiguousarray(ta).reshape(-1), a)
    b = np.zeros(36 * 18)
    R.get_grid_great_circle_area(C.byref(C.c_int(36)), C.byref(C.c_int(18)), lo.reshape(-1), la.reshape(-1), b)
    np.savez_compressed(os.path.join(HERE, "gc_areas.npz"), tripolar=a, latlon=b)


def latlon(nlon, nlat, lon0=0.0, lon1=360.0, lat0=-90.0, lat1=90.0):
    lon = np.array([(lon0 + i * ((lon1 - lon0) / nlon)) * D2R for i in range(nlon + 1)])
    lat = np.array([(lat0 + j * ((lat1 - lat0) / nlat)) * D2R for j in range(nlat + 1)])
    return np.ascontiguousarray(np.tile(lon, (nlat + 1, 1))), np.ascontiguousarray(np.tile(lat[:, None], (1, nlon + 1)))


def main():
    make_polys()
    c8 = xgtest.ref_cubed_sphere(8, centers=True)
    np.savez_compressed(os.path.join(HERE, "grid_c8.npz"), lonc=c8[0], latc=c8[1], lont=c8[2], latt=c8[3])
    lo, la = latlon(36, 18)
    make_xgrid("c8_36x18_o1", c8[0], c8[1], lo, la, 1)
    make_xgrid("c8_36x18_o2", c8[0], c8[1], lo, la, 2)
    c12 = xgtest.ref_cubed_sphere(12)
    lo, la = latlon(72, 36)
    make_xgrid("c12_72x36_o2", c12[0], c12[1], lo, la, 2)
    l1 = latlon(40, 20)
    l2 = latlon(25, 20, -30.0, 95.0, -63.0, 77.0)
    make_xgrid("ll40x20_regional_o2", [l1[0]], [l1[1]], l2[0], l2[1], 2)
    c10 = xgtest.ref_cubed_sphere(10)
    make_xgrid("c12_to_c10tile3_o2", c12[0], c12[1], c10[0][2], c10[1][2], 2)
    make_apply("c8_36x18", 8, 36, 18)
    make_gc()


if __name__ == "__main__":
    main()
