"""CPU-side checks (no GPU): the oracle restatement against golden vectors produced by the unmodified
reference (tests/golden/, made by tests/golden/make_golden.py), and against the compiled reference
itself when oracle/_ref exists.  Integer results must be bit-exact; floating-point fields are compared
bit-for-bit when the host libm is the one the vectors were made with, else to 1e-12 of the parent scale."""
import os

import numpy as np
import pytest

import xgtest

G = xgtest.GOLDEN_DIR
MV = 50


def _load(name):
    return np.load(os.path.join(G, name))


def _close(a, b, scale=None):
    if np.array_equal(a, b):
        return True
    den = np.maximum(np.abs(b), 1e-300) if scale is None else scale
    return float(np.max(np.abs(a - b) / den)) <= 1e-12


def test_polygon_primitives_match_golden():
    O = xgtest.oracle_lib()
    g = _load("polys.npz")
    n = g["n1"].size
    for i in range(n):
        for tag in ("1", "2"):
            x = np.zeros(MV); y = np.zeros(MV)
            x[:4] = g["x" + tag][i]; y[:4] = g["y" + tag][i]
            m = O.orc_fix_lon(x, y, 4, np.pi)
            assert m == g["n" + tag][i]
            assert np.array_equal(y[:m], g["fy" + tag][i][:m])
            if tag == "1":
                assert np.array_equal(x[:m], g["fx1"][i][:m])
                assert _close(np.array([O.orc_poly_area(x, y, m)]), g["area1"][i:i + 1])
        n1, n2 = int(g["n1"][i]), int(g["n2"][i])
        a1 = np.zeros(MV); b1 = np.zeros(MV); a2 = np.zeros(MV); b2 = np.zeros(MV)
        a1[:10], b1[:10], a2[:10], b2[:10] = g["fx1"][i], g["fy1"][i], g["fx2"][i], g["fy2"][i]
        ox = np.zeros(MV); oy = np.zeros(MV)
        no = O.orc_clip_2dx2d(a1, b1, n1, a2, b2, n2, ox, oy)
        assert no == g["n_out"][i], i
        assert np.array_equal(ox[:no], g["ox"][i][:no]) and np.array_equal(oy[:no], g["oy"][i][:no]), i
        if no > 0:
            sc = np.array([min(g["area1"][i], g["area2"][i])])
            assert _close(np.array([O.orc_poly_area(ox, oy, no)]), g["xarea"][i:i + 1], sc), i
            assert _close(np.array([O.orc_poly_ctrlon(ox, oy, no, a1[:n1].mean())]), g["ctrlon"][i:i + 1], sc), i
            assert _close(np.array([O.orc_poly_ctrlat(ox, oy, no)]), g["ctrlat"][i:i + 1], sc), i


@pytest.mark.parametrize("tag", ["c8_36x18_o1", "c8_36x18_o2", "c12_72x36_o2", "ll40x20_regional_o2", "c12_to_c10tile3_o2"])
def test_setup_conserve_interp_matches_golden(tag):
    g = _load(f"xgrid_{tag}.npz")
    nx, ny = g["nx"], g["ny"]
    lons, lats, off = [], [], 0
    for t in range(nx.size):
        nv = (nx[t] + 1) * (ny[t] + 1)
        lons.append(g["lon_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1))
        lats.append(g["lat_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1))
        off += nv
    opcode = int(g["opcode"])
    got = xgtest.oracle_setup(lons, lats, g["lon_out"], g["lat_out"], opcode)
    assert got["nxgrid"] == g["area"].size
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(got[k], g[k]), k
    sc = xgtest.parent_scale(got, lons, lats, g["lon_out"], g["lat_out"])
    assert _close(got["area"], g["area"], sc)
    if opcode & 2:
        assert np.max(np.abs(got["di"] - g["di"])) <= 1e-11 and np.max(np.abs(got["dj"] - g["dj"])) <= 1e-11
    if xgtest.libm_matches_ref_trig():      # same libm as the one the vectors were made with: everything is bit-identical
        assert np.array_equal(got["area"], g["area"])
        if opcode & 2:
            assert np.array_equal(got["di"], g["di"]) and np.array_equal(got["dj"], g["dj"])


def test_grid_synthesis_matches_golden(pkg):
    g = _load("grid_c8.npz")
    lonc, latc, lont, latt = pkg.cubed_sphere_grid(8, centers=True)
    for a, b in ((lonc, g["lonc"]), (latc, g["latc"]), (lont, g["lont"]), (latt, g["latt"])):
        assert np.max(np.abs(a - b)) <= 4e-16          # bit-identical with the generating libm
    lo, la = pkg.latlon_grid(36, 18)
    x = _load("xgrid_c8_36x18_o1.npz")
    assert np.array_equal(lo, x["lon_out"]) and np.array_equal(la, x["lat_out"])


def test_oracle_equals_compiled_reference(reflib, pkg):
    """live check against the unmodified reference (only where oracle/_ref exists)"""
    for ni, nlon, nlat, opcode in ((16, 90, 45, 1), (16, 90, 45, 2), (20, 64, 40, 2)):
        lonc, latc = xgtest.ref_cubed_sphere(ni)
        assert np.array_equal(lonc, pkg.cubed_sphere_grid(ni)[0])
        lon2, lat2 = pkg.latlon_grid(nlon, nlat)
        ref = xgtest.ref_setup(lonc, latc, lon2, lat2, opcode)
        got = xgtest.oracle_setup(lonc, latc, lon2, lat2, opcode)
        assert got["nxgrid"] == ref["nxgrid"]
        for k in ref:
            if k != "nxgrid":
                assert np.array_equal(got[k], ref[k]), k


def test_row_band_decomposition_concatenates(reflib, pkg):
    """fregrid_parallel's layout {1,npes}: destination row bands, concatenated in rank order, hold the same
    cells as the serial run after a canonical sort (conserve_interp.c:404-437 gathers in rank order)."""
    lonc, latc = pkg.cubed_sphere_grid(12)
    lon2, lat2 = pkg.latlon_grid(48, 24)
    full = xgtest.ref_setup(lonc, latc, lon2, lat2, 1)
    ib = np.zeros(3, np.int32); ie = np.zeros(3, np.int32)
    reflib.ref_compute_extent(24, 3, ib, ie)
    parts = [xgtest.ref_setup(lonc, latc, lon2, lat2, 1, jsc=int(a), jec=int(b)) for a, b in zip(ib, ie)]
    for p, a in zip(parts, ib):
        p["j_out"] = p["j_out"] + a                      # band-relative on each rank (conserve_interp.c:421)
    cat = {k: np.concatenate([p[k] for p in parts]) for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area")}
    cat["nxgrid"] = cat["area"].size
    xgtest.assert_xgrid_equal(cat, full, 1, same_order=False, exact=True)


def test_reference_embedded_polygon_cases_match_the_oracle():
    """The 26 hand-built cases of the reference's own print-only harness (create_xgrid.c:2383-3010: poles, tripolar fold,
    identical boxes, containment, sides through the south pole, twin pole vertices), recorded in full precision from the
    unmodified reference by tests/golden/make_polycases_golden.py.  The oracle restatement reproduces every one bit for
    bit: clip_2dx2d_great_circle (1-10), create_xgrid_great_circle (11-14), clip_2dx2d + fix_lon + poly_area (15-26), and the
    2dx2d / great-circle generators on the quadrilateral pairs taken as 1x1 grids."""
    O = xgtest.oracle_lib()
    g = _load("ref_polycases.npz")
    seen = 0
    for n in range(1, 27):
        k = f"c{n:02d}_"
        x1, y1, x2, y2 = g[k + "lon1"], g[k + "lat1"], g[k + "lon2"], g[k + "lat2"]
        n1, n2, nlon1, nlat1, nlon2, nlat2 = (int(v) for v in g[k + "dims"])
        if n <= 10:
            a, b = g[k + "xyz1"], g[k + "xyz2"]
            o = [np.zeros(MV) for _ in range(3)]
            no = O.orc_clip_2dx2d_great_circle(*[np.ascontiguousarray(v) for v in a], 4, *[np.ascontiguousarray(v) for v in b], n2, *o)
            assert no == int(g[k + "gc_n"]), n
            assert np.array_equal(np.stack([v[:no] for v in o]), g[k + "gc_xyz"]), n
        elif n <= 14:
            cap = 4096
            bi = [np.zeros(cap, np.int32) for _ in range(4)]
            xa = np.zeros(cap)
            nx = O.orc_create_xgrid_great_circle(nlon1, nlat1, nlon2, nlat2, np.ascontiguousarray(x1), np.ascontiguousarray(y1),
                                                 np.ascontiguousarray(x2), np.ascontiguousarray(y2), np.ones(max(nlon1 * nlat1, 1)), cap,
                                                 *bi, xa, None, None)
            assert nx == int(g[k + "gcx_n"]), n
            assert np.array_equal(np.stack([v[:nx] for v in bi]), g[k + "gcx_idx"]), n
            assert np.array_equal(xa[:nx], g[k + "gcx_area"]), n
        else:
            a1 = np.zeros(MV); b1 = np.zeros(MV); a2 = np.zeros(MV); b2 = np.zeros(MV)
            a1[:n1] = x1; b1[:n1] = y1; a2[:n2] = x2; b2[:n2] = y2
            lo = np.zeros(MV); la = np.zeros(MV)
            no = O.orc_clip_2dx2d(a1, b1, n1, a2, b2, n2, lo, la)
            assert no == int(g[k + "clip_n"]), n
            assert np.array_equal(lo[:no], g[k + "clip_lon"]) and np.array_equal(la[:no], g[k + "clip_lat"]), n
            f1 = O.orc_fix_lon(a1, b1, n1, np.pi); f2 = O.orc_fix_lon(a2, b2, n2, np.pi); fo = O.orc_fix_lon(lo, la, no, np.pi)
            assert [f1, f2, fo] == g[k + "fix_n"].tolist(), n
            assert np.array_equal(np.stack([a1[:f1], b1[:f1]]), g[k + "fix1"]) and np.array_equal(np.stack([lo[:fo], la[:fo]]), g[k + "fixo"]), n
            areas = np.array([O.orc_poly_area(a1, b1, f1), O.orc_poly_area(a2, b2, f2), O.orc_poly_area(lo, la, fo)])
            assert _close(areas, g[k + "areas"], np.maximum(np.abs(g[k + "areas"]), 1.0)), n
        if k + "cell_o2" in g:
            cell = lambda v: np.ascontiguousarray(np.array([v[0], v[1], v[3], v[2]]))
            bi = [np.zeros(64, np.int32) for _ in range(4)]
            xa = np.zeros(64); xc = np.zeros(64); yc = np.zeros(64)
            nx = O.orc_create_xgrid_2dx2d(2, 1, 1, 1, 1, cell(x1), cell(y1), cell(x2), cell(y2), np.ones(1), 64, *bi, xa,
                                          xc.ctypes.data, yc.ctypes.data)
            want = g[k + "cell_o2"]
            assert nx == int(want[0]), n
            assert _close(np.concatenate([xa[:nx], xc[:nx], yc[:nx]]), want[1:], np.maximum(np.abs(want[1:]), 1.0)), n
            nx = O.orc_create_xgrid_great_circle(1, 1, 1, 1, cell(x1), cell(y1), cell(x2), cell(y2), np.ones(1), 64, *bi, xa, None, None)
            want = g[k + "cell_gc"]
            assert nx == int(want[0]) and np.array_equal(xa[:nx], want[1:]), n
            seen += 1
    assert seen >= 14


def test_reference_embedded_great_circle_cases_match_the_product_clip(pkg):
    """cases 1-10 through the product's own great-circle clip (host build of csrc/gc_clip.cuh, the code the kernel runs):
    same vertex counts; coordinates to one ulp (the reference solves a 3x3 system in x87 long double, the product in
    double-double)"""
    g = _load("ref_polycases.npz")
    L = pkg.lib()
    for n in range(1, 11):
        k = f"c{n:02d}_"
        a, b = g[k + "xyz1"], g[k + "xyz2"]
        o = [np.zeros(MV) for _ in range(3)]
        no = L.xgb_gc_clip_host(*[np.ascontiguousarray(v).ctypes.data for v in a], 4, *[np.ascontiguousarray(v).ctypes.data for v in b],
                                int(g[k + "dims"][1]), *[v.ctypes.data for v in o], None)
        assert no == int(g[k + "gc_n"]), n
        if no:
            assert np.max(np.abs(np.stack([v[:no] for v in o]) - g[k + "gc_xyz"])) <= 4.5e-16, n


def test_order2_distance_restatement_equals_the_setup_restatement():
    """oracle_order2_distance (the centroid correction alone, used to check full-size GPU lists) against the oracle's whole
    setup_conserve_interp, which the golden vectors and the compiled reference pin"""
    pkg = xgtest.package()
    lonc, latc = pkg.cubed_sphere_grid(12)
    lon2, lat2 = pkg.latlon_grid(72, 36)
    full = xgtest.oracle_setup(lonc, latc, lon2, lat2, 2, raw=True)
    di, dj = xgtest.oracle_order2_distance(lonc, latc, full)
    assert np.array_equal(di, full["di"]) and np.array_equal(dj, full["dj"])
    if xgtest.ref_lib() is not None:
        ref = xgtest.ref_setup(lonc, latc, lon2, lat2, 2)
        assert np.array_equal(full["di"], ref["di"]) and np.array_equal(full["dj"], ref["dj"])
        for jsc, jec in ((0, 0), (17, 19), (35, 35)):
            b = xgtest.ref_band_xgrid(lonc, latc, lon2, lat2, 2, jsc, jec)
            m = (full["j_out"] >= jsc) & (full["j_out"] <= jec)
            for key in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "xgrid_clon", "xgrid_clat"):
                assert np.array_equal(b[key], full[key][m]), (jsc, key)
