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
