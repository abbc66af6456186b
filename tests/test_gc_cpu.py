"""CPU-side checks of the great-circle path (no GPU):
  * the oracle restatement (oracle/xgrid_oracle_gc.c) against golden vectors produced by the unmodified reference and,
    where oracle/_ref exists, live against the compiled reference: bit-exact (both run the x87 long double solve);
  * the HOST BUILD of the product's device code (csrc/gc_clip.cuh, through xgb_gc_clip_host) against the same vectors:
    vertex counts exact, vertices to 1 ulp, areas to 1e-14 steradian (double-double instead of x87, see gc_clip.cuh)."""
import ctypes as C
import os

import numpy as np
import pytest

import xgtest

R2 = 6371000.0 ** 2


def _golden_polys():
    return np.load(os.path.join(xgtest.GOLDEN_DIR, "gc_polys.npz"))


def test_oracle_gc_clip_matches_golden():
    O = xgtest.oracle_lib()
    g = _golden_polys()
    for k in range(g["n_out"].size):
        a = [np.ascontiguousarray(v) for v in g["p1"][k]]; b = [np.ascontiguousarray(v) for v in g["p2"][k]]
        o = [np.zeros(60) for _ in range(3)]
        n = O.orc_clip_2dx2d_great_circle(*a, 4, *b, 4, *o)
        assert n == g["n_out"][k], k
        for c in range(3):
            assert np.array_equal(o[c][:n], g["out"][k, c, :n]), k
        if n > 0:
            assert O.orc_great_circle_area(n, *o) == g["area"][k], k


def test_product_host_build_gc_clip_matches_golden(pkg):
    L = pkg.lib()
    g = _golden_polys()
    ndiff = 0
    for k in range(g["n_out"].size):
        a = [np.ascontiguousarray(v) for v in g["p1"][k]]; b = [np.ascontiguousarray(v) for v in g["p2"][k]]
        o = [np.zeros(60) for _ in range(3)]
        area = C.c_double(0)
        n = L.xgb_gc_clip_host(*[v.ctypes.data for v in a], 4, *[v.ctypes.data for v in b], 4, *[v.ctypes.data for v in o], C.byref(area))
        assert n == g["n_out"][k], k
        for c in range(3):
            assert np.max(np.abs(o[c][:n] - g["out"][k, c, :n]), initial=0.0) <= 2.3e-16, k
            ndiff += int(not np.array_equal(o[c][:n], g["out"][k, c, :n]))
        if n > 0:
            assert abs(area.value - g["area"][k]) / R2 <= 1e-14, k
    assert ndiff <= g["n_out"].size // 100          # double-double vs x87: all but a handful of vertices are bit-identical


def test_oracle_gc_exchange_grids_match_golden():
    for tag in ("gc_c8_36x18", "gc_tripolar24x18_36x18"):
        g = np.load(os.path.join(xgtest.GOLDEN_DIR, f"xgrid_{tag}.npz"))
        nx, ny = g["nx"], g["ny"]
        lons, lats, off = [], [], 0
        for t in range(nx.size):
            nv = (nx[t] + 1) * (ny[t] + 1)
            lons.append(g["lon_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1)); lats.append(g["lat_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1))
            off += nv
        got = xgtest.oracle_setup(lons, lats, g["lon_out"], g["lat_out"], int(g["opcode"]))
        assert got["nxgrid"] == g["area"].size
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area"):
            assert np.array_equal(got[k], g[k]), (tag, k)
    a = np.load(os.path.join(xgtest.GOLDEN_DIR, "gc_areas.npz"))
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "xgrid_gc_tripolar24x18_36x18.npz"))
    out = np.zeros(24 * 18)
    xgtest.oracle_lib().orc_get_grid_great_circle_area(24, 18, np.ascontiguousarray(g["lon_in"]), np.ascontiguousarray(g["lat_in"]), out)
    assert np.array_equal(out, a["tripolar"])
    out = np.zeros(36 * 18)
    xgtest.oracle_lib().orc_get_grid_great_circle_area(36, 18, np.ascontiguousarray(g["lon_out"]).reshape(-1), np.ascontiguousarray(g["lat_out"]).reshape(-1), out)
    assert np.array_equal(out, a["latlon"])
    # the great-circle exchange grid tiles the sphere
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "xgrid_gc_c8_36x18.npz"))
    assert abs(g["area"].sum() / (4 * np.pi * R2) - 1) < 1e-12


def test_oracle_gc_equals_compiled_reference(reflib, pkg):
    """live: other grids than the goldens, incl. a tripolar grid whose bipolar cap crosses the lat-lon pole rows"""
    GC = xgtest.GREAT_CIRCLE
    lonc, latc = xgtest.ref_cubed_sphere(10)
    lon2, lat2 = pkg.latlon_grid(30, 20)
    cases = [((lonc, latc), (lon2, lat2))]
    tl, ta = xgtest.tripolar_grid(64, 48)
    cases.append((([tl], [ta]), pkg.latlon_grid(24, 12)))
    for (a, b), (lo, la) in cases:
        ref = xgtest.ref_setup(a, b, lo, la, 1 | GC)
        got = xgtest.oracle_setup(a, b, lo, la, 1 | GC)
        assert got["nxgrid"] == ref["nxgrid"]
        for k in ref:
            if k != "nxgrid":
                assert np.array_equal(got[k], ref[k]), k


def test_gc_acos_host_against_acosl(pkg):
    """gc_acos (csrc/gc_clip.cuh) = acosl(x) rounded to double as spherical_angle leaves it (mosaic_util.c:834).  numpy's
    longdouble arccos IS libm's acosl; the restatement must give its bits except where the x87 fpatan is not the correctly
    rounded 64-bit result AND that result sits on a double rounding boundary (measured here: 3.5 per 100 000 arguments; plain
    double acos differs on 5 %)."""
    import ctypes as C
    L = pkg.lib()
    L.xgb_gc_acos_host.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]; L.xgb_gc_acos_host.restype = None
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 1_000_000), 1 - 10.0 ** rng.uniform(-16, 0, 300_000), -1 + 10.0 ** rng.uniform(-16, 0, 300_000),
                        rng.uniform(-0.05, 0.05, 400_000), [1.0, -1.0, 0.0, 0.5, -0.5]])
    out = np.empty_like(x)
    L.xgb_gc_acos_host(x.size, x.ctypes.data, out.ctypes.data)
    ref = np.arccos(x.astype(np.longdouble)).astype(np.float64)
    if np.finfo(np.longdouble).nmant != 63:
        pytest.skip("no 80-bit long double on this machine")
    bad = out != ref
    assert bad.mean() < 2e-4, bad.mean()
    assert np.max(np.abs(out - ref)) <= 4.5e-16
    assert (np.arccos(x) != ref).mean() > 0.01          # what the toolchain's acos would give
