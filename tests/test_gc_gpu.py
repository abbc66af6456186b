"""GPU parity tests of the great-circle exchange-grid path (create_xgrid_great_circle through the C ABI) against the CPU
oracle and the reference's golden vectors.  Bar: integer cell lists bit-exact in the reference's emission order;
xgrid_area to 8e-15 steradian absolute — the spherical excess is a sum of O(1) angles minus (n-2)*pi, each angle an
acosl() rounded to double in the reference (x87), so one ulp of one angle is 2e-16 sr whatever the size of the cell
(see csrc/gc_clip.cuh); relative to the cell that is 1e-11 for a quarter-degree cell."""
import os

import numpy as np
import pytest

import xgtest

pytestmark = pytest.mark.gpu

R2 = 6371000.0 ** 2
AREA_ATOL_SR = 8e-15
GC = xgtest.GREAT_CIRCLE


def test_gc_acos_device_equals_host_equals_acosl(pkg):
    """spherical_angle's acosl() rounded to double (mosaic_util.c:834): the double-double restatement gives the same bits on
    the device as on the host, and the bits of this machine's x87 acosl but for the few arguments per 100 000 where fpatan is
    not the correctly rounded 64-bit result"""
    import ctypes as C
    L = pkg.lib()
    L.xgb_gc_acos_host.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]; L.xgb_gc_acos_host.restype = None
    L.xgb_gc_acos_device.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1, 1, 1_000_000), 1 - 10.0 ** rng.uniform(-16, 0, 200_000), -1 + 10.0 ** rng.uniform(-16, 0, 200_000),
                        rng.uniform(-0.05, 0.05, 600_000), [1.0, -1.0, 0.0, -0.0, 0.5, -0.5, 1 - 2.0 ** -53, -1 + 2.0 ** -53]])
    h = np.empty_like(x); d = np.empty_like(x)
    L.xgb_gc_acos_host(x.size, x.ctypes.data, h.ctypes.data)
    assert L.xgb_gc_acos_device(x.size, x.ctypes.data, d.ctypes.data) == 0
    assert np.array_equal(h.view(np.int64), d.view(np.int64))
    ref = np.arccos(x.astype(np.longdouble)).astype(np.float64)
    bad = h != ref
    assert bad.mean() < 2e-4, bad.mean()
    assert np.max(np.abs(h - ref)) <= 4.5e-16


def _gen(pkg, lonc, latc, lon2, lat2, mask=None):
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc, mask)
    n = plan.generate(1 | GC)
    got = plan.result_host()
    got["nxgrid"] = n
    got["npairs"] = plan.npairs
    got["area_src"] = plan.great_circle_area("src"); got["area_dst"] = plan.great_circle_area("dst")
    plan.close()
    return got


def _check(got, ref):
    assert got["nxgrid"] == ref["nxgrid"], (got["nxgrid"], ref["nxgrid"])
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(got[k], ref[k]), k
    d = np.abs(got["area"] - ref["area"]) / R2
    assert d.max(initial=0.0) <= AREA_ATOL_SR, d.max()
    return float(np.mean(got["area"] == ref["area"]))


@pytest.mark.parametrize("tag", ["gc_c8_36x18", "gc_tripolar24x18_36x18"])
def test_gc_matches_reference_golden(pkg, tag):
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, f"xgrid_{tag}.npz"))
    nx, ny = g["nx"], g["ny"]
    lons, lats, off = [], [], 0
    for t in range(nx.size):
        nv = (nx[t] + 1) * (ny[t] + 1)
        lons.append(g["lon_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1)); lats.append(g["lat_in"][off:off + nv].reshape(ny[t] + 1, nx[t] + 1))
        off += nv
    got = _gen(pkg, lons, lats, g["lon_out"], g["lat_out"])
    ref = {k: g[k] for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area")}
    ref["nxgrid"] = g["area"].size
    _check(got, ref)
    if tag.startswith("gc_tripolar"):
        a = np.load(os.path.join(xgtest.GOLDEN_DIR, "gc_areas.npz"))
        assert np.max(np.abs(got["area_src"] - a["tripolar"])) / R2 <= AREA_ATOL_SR
        assert np.max(np.abs(got["area_dst"] - a["latlon"])) / R2 <= AREA_ATOL_SR


@pytest.mark.parametrize("ni,nlon,nlat", [(16, 90, 45), (48, 360, 180)])
def test_gc_cubed_sphere_matches_oracle(pkg, ni, nlon, nlat):
    """C48 -> 1 degree: SURVEY 8c sanity value nxgrid 146016 for the great-circle algorithm"""
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    got = _gen(pkg, lonc, latc, lon2, lat2)
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, 1 | GC)
    same = _check(got, ref)
    if ni == 48:
        assert got["nxgrid"] == 146016
    assert abs(got["area"].sum() - ref["area"].sum()) / ref["area"].sum() < 1e-13
    assert abs(got["area"].sum() / (4 * np.pi * R2) - 1) < 1e-8      # slivers below the 1e-6 area ratio are dropped
    print(f"great-circle C{ni}: {same * 100:.3f} % of the areas bit-identical")
    assert same > 0.995                                 # the rest differ by an ulp of one angle (fpatan not correctly rounded) or of a vertex


def test_gc_latlon_to_latlon_shared_edges_and_mask(pkg):
    """vertices of one grid lying exactly on the sides of the other (the fragile inside/outside decisions of insidePolygon),
    pole triangles on both sides, and a source mask"""
    lon1, lat1 = pkg.latlon_grid(40, 20)
    lon2, lat2 = pkg.latlon_grid(60, 30)
    rng = np.random.default_rng(3)
    mask = (rng.uniform(size=40 * 20) > 0.2).astype(np.float64)
    got = _gen(pkg, [lon1], [lat1], lon2, lat2, mask)
    L = xgtest.oracle_lib()
    cap = 40000
    out = {k: np.zeros(cap, np.int32) for k in ("i_in", "j_in", "i_out", "j_out")}
    area = np.zeros(cap)
    n = L.orc_create_xgrid_great_circle(40, 20, 60, 30, lon1.reshape(-1), lat1.reshape(-1), lon2.reshape(-1), lat2.reshape(-1), mask, cap,
                                        out["i_in"], out["j_in"], out["i_out"], out["j_out"], area, None, None)
    assert n == got["nxgrid"], (n, got["nxgrid"])
    for k in out:
        assert np.array_equal(got[k], out[k][:n]), k
    assert np.max(np.abs(got["area"] - area[:n])) / R2 <= AREA_ATOL_SR


def test_gc_reference_signature_entry_points(pkg):
    lonc, latc = pkg.cubed_sphere_grid(8)
    lon2, lat2 = pkg.latlon_grid(36, 18)
    r = pkg.create_xgrid_great_circle(lonc[2], latc[2], lon2, lat2)
    ref = xgtest.oracle_setup([lonc[2]], [latc[2]], lon2, lat2, 1 | GC)
    assert r[0] == ref["nxgrid"]
    for a, k in zip(r[1:5], ("i_in", "j_in", "i_out", "j_out")):
        assert np.array_equal(a, ref[k]), k
    assert np.max(np.abs(r[5] - ref["area"])) / R2 <= AREA_ATOL_SR
    assert not r[6].any() and not r[7].any()
    a = pkg.get_grid_great_circle_area(lon2, lat2)
    want = np.zeros(36 * 18)
    xgtest.oracle_lib().orc_get_grid_great_circle_area(36, 18, lon2.reshape(-1), lat2.reshape(-1), want)
    assert np.max(np.abs(a.reshape(-1) - want)) / R2 <= AREA_ATOL_SR
    with pytest.raises(pkg.XgridError):                 # fregrid.c:763: great circle is first order only
        p = pkg.XgridPlan(0); p.set_dst(lon2, lat2); p.set_src(lonc, latc); p.generate(2 | GC)


def test_gc_tripolar_reduced_config3_matches_oracle(pkg):
    """BASELINE configs[2] at reduced size: 1 degree tripolar ocean grid (360x270, bipolar cap north of 65N) -> 2 degree
    lat-lon with the great-circle algorithm, against the oracle (O(N1*N2) range checks: a few seconds)"""
    if xgtest.ref_lib() is None:
        pytest.skip("tripolar grid generator lives in oracle/_ref")
    tl, ta = xgtest.tripolar_grid(720, 540)
    lon2, lat2 = pkg.latlon_grid(180, 90)
    got = _gen(pkg, [tl], [ta], lon2, lat2)
    ref = xgtest.oracle_setup([tl], [ta], lon2, lat2, 1 | GC)
    _check(got, ref)


def test_gc_tripolar_config3_full_size_properties(pkg):
    """BASELINE configs[2] at full size: 1/4 degree tripolar (1440x1080) -> 1 degree lat-lon, great circle.  Too big for the
    oracle; size-independent properties instead: the exchange cells of a source cell tile it (sum of their areas == its
    spherical-excess area), likewise for every destination cell the ocean grid covers completely, list is in emission
    order, no duplicates."""
    import time
    if xgtest.ref_lib() is None:
        pytest.skip("tripolar grid generator lives in oracle/_ref")
    tl, ta = xgtest.tripolar_grid(2880, 2160)
    lon2, lat2 = pkg.latlon_grid(360, 180)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src([tl], [ta])
    t0 = time.perf_counter(); n = plan.generate(1 | GC); plan.sync(); t1 = time.perf_counter()
    n = plan.generate(1 | GC); plan.sync(); t2 = time.perf_counter()
    x = plan.result_host()
    a_src = plan.great_circle_area("src"); a_dst = plan.great_circle_area("dst")
    last, _, _ = plan.phase_ms()
    print(f"config3 great circle: nxgrid {n}, candidate pairs {plan.npairs}, generate {1e3 * (t2 - t1):.1f} ms (first call {1e3 * (t1 - t0):.1f}), phases {last}")
    plan.close()
    nx1 = tl.shape[1] - 1
    s = x["j_in"].astype(np.int64) * nx1 + x["i_in"]
    d = x["j_out"].astype(np.int64) * 360 + x["i_out"]
    key = s * (360 * 180) + d
    assert np.all(np.diff(key) > 0)                                   # emission order, no duplicates
    per_src = np.bincount(s, weights=x["area"], minlength=a_src.size)
    over = np.nonzero(per_src > a_src * (1 + 1e-9))[0]
    under = np.nonzero(per_src < a_src * (1 - 2e-5))[0]
    print(f"source cells over-covered: {over.size} (max ratio {np.max(per_src / a_src):.12f}) rows {np.unique(over // nx1)[:10]}..; "
          f"under-covered: {under.size} (min ratio {np.min(per_src / a_src):.12f}) rows {np.unique(under // nx1)[:10]}")
    # slivers below 1e-6 of the smaller parent are dropped by the reference's accept test: the sums fall short by at most that
    # optional dump of the bipolar-cap rows for an offline comparison with the oracle (scripts/check_gc_config3_rows.py)
    dump = os.environ.get("XGB_DUMP_GC_ROWS")
    if dump:
        m = x["j_in"] >= (tl.shape[0] - 1) - 24
        np.savez_compressed(dump, **{k: x[k][m] for k in ("i_in", "j_in", "i_out", "j_out", "area")})
    # against the CPU oracle on slabs of source rows (the oracle scans every destination cell per source cell, so a slab of the
    # full-size problem takes seconds): five mid-latitude rows and the three northernmost rows of the bipolar cap, where the
    # grid is least lat-lon like.  Same lists in the same order; areas to the great-circle tolerance.
    ny1 = tl.shape[0] - 1
    for ja, jb in ((500, 505), (ny1 - 3, ny1)):
        ref = xgtest.oracle_setup([tl[ja:jb + 1]], [ta[ja:jb + 1]], lon2, lat2, 1 | GC)
        m = (x["j_in"] >= ja) & (x["j_in"] < jb)
        assert int(m.sum()) == ref["nxgrid"] > 0, (ja, int(m.sum()), ref["nxgrid"])
        assert np.array_equal(x["i_in"][m], ref["i_in"]) and np.array_equal(x["j_in"][m], ref["j_in"] + ja)
        assert np.array_equal(x["i_out"][m], ref["i_out"]) and np.array_equal(x["j_out"][m], ref["j_out"])
        assert np.max(np.abs(x["area"][m] - ref["area"])) / R2 <= AREA_ATOL_SR
    # tolerances: one ulp of one angle is 2e-16 sr whatever the cell size (absolute term), and the reference's accept test
    # keeps slivers down to 1e-6 of the smaller parent and drops the rest (relative term)
    slack = 2e-5 * a_src + 1e-13 * R2
    assert np.all(per_src <= a_src + slack)
    assert np.all(per_src >= a_src - slack)
    assert abs(per_src.sum() / a_src.sum() - 1) < 1e-6
    per_dst = np.bincount(d, weights=x["area"], minlength=a_dst.size)
    assert np.all(per_dst <= a_dst * (1 + 2e-5) + 1e-13 * R2)
