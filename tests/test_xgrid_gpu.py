"""GPU parity tests of exchange-grid weight generation: CUDA path (through the C ABI) vs the CPU oracle.

Bar (BASELINE.json north_star): integer cell lists bit-exact, xgrid_area within 1e-12 relative;
tile1_distance (a difference of near-equal centroids) within 1e-11 rad absolute.
"""
import os
import numpy as np
import pytest

import xgtest

pytestmark = pytest.mark.gpu

AREA_RTOL = 1e-12
DIST_ATOL = 1e-9


def test_ref_trig_device_equals_host_and_libm(pkg):
    """csrc/ref_trig.cuh: device build == host build == the libm the reference links (bit for bit)."""
    x = xgtest.trig_samples(300000)
    host = [np.empty_like(x) for _ in range(4)]
    dev = [np.empty_like(x) for _ in range(4)]
    pkg.lib().xgb_ref_trig_host(x.size, *[a.ctypes.data for a in [x] + host])
    assert pkg.lib().xgb_ref_trig_device(x.size, *[a.ctypes.data for a in [x] + dev]) == 0
    for a, b in zip(host, dev):
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    if xgtest.libm_matches_ref_trig():
        for a, b in zip(dev, xgtest.libm_trig(x)):
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def _gen(pkg, lonc, latc, lon2, lat2, opcode, mask=None):
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc, mask)
    n = plan.generate(opcode)
    got = plan.result_host()
    got["nxgrid"] = n
    got["npairs"] = plan.npairs
    plan.close()
    return got


@pytest.mark.parametrize("ni,nlon,nlat", [(8, 36, 18), (16, 180, 90), (48, 360, 180)])
@pytest.mark.parametrize("order", [1, 2])
def test_cubed_sphere_to_latlon_matches_oracle(pkg, ni, nlon, nlat, order):
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    got = _gen(pkg, lonc, latc, lon2, lat2, order)
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, order)
    sc = xgtest.parent_scale(ref, lonc, latc, lon2, lat2)
    xgtest.assert_xgrid_equal(got, ref, order, AREA_RTOL, DIST_ATOL, same_order=True, scale=sc)
    # per-exchange-cell relative difference for cells that are not slivers (>= 1% of the parent cell)
    big = ref["area"] >= 1e-2 * sc
    assert np.max(np.abs(got["area"][big] - ref["area"][big]) / ref["area"][big]) <= 1e-10


def test_c48_known_answer(pkg):
    """SURVEY 8c sanity values: C48 -> 360x180 order-1 nxgrid 146032, sum(area)/4piR^2 = 0.9999999995881"""
    lonc, latc = pkg.cubed_sphere_grid(48)
    lon2, lat2 = pkg.latlon_grid(360, 180)
    got = _gen(pkg, lonc, latc, lon2, lat2, 1)
    assert got["nxgrid"] == 146032
    frac = np.sum(got["area"]) / (4 * np.pi * xgtest.RADIUS ** 2)
    assert abs(frac - 0.999999999588102) < 1e-12


def test_c96_quarter_degree_order2_matches_oracle(pkg):
    """BASELINE config 2 weights: C96 -> 1440x720 order 2 (nxgrid 1645856)."""
    lonc, latc = pkg.cubed_sphere_grid(96)
    lon2, lat2 = pkg.latlon_grid(1440, 720)
    got = _gen(pkg, lonc, latc, lon2, lat2, 2)
    assert got["nxgrid"] == 1645856
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, 2)
    sc = xgtest.parent_scale(ref, lonc, latc, lon2, lat2)
    xgtest.assert_xgrid_equal(got, ref, 2, AREA_RTOL, DIST_ATOL, same_order=True, scale=sc)


def test_latlon_to_latlon_and_regional(pkg):
    """lat-lon source (pole rows, cyclic seam) onto a shifted regional lat-lon window."""
    lon1, lat1 = pkg.latlon_grid(72, 36)
    lon2, lat2 = pkg.latlon_grid(50, 40, lonbegin=-30.0, lonend=95.0, latbegin=-63.0, latend=77.0)
    for order in (1, 2):
        got = _gen(pkg, [lon1], [lat1], lon2, lat2, order)
        ref = xgtest.oracle_setup([lon1], [lat1], lon2, lat2, order)
        sc = xgtest.parent_scale(ref, [lon1], [lat1], lon2, lat2)
        xgtest.assert_xgrid_equal(got, ref, order, AREA_RTOL, DIST_ATOL, scale=sc)


def test_cubed_sphere_to_cubed_sphere_tile(pkg):
    """curvilinear destination (a rotated cube face incl. the pole tile): exercises the generic pyramid."""
    lonc, latc = pkg.cubed_sphere_grid(24)
    lond, latd = pkg.cubed_sphere_grid(20)
    for tile in (0, 2, 5):
        got = _gen(pkg, lonc, latc, lond[tile], latd[tile], 2)
        ref = xgtest.oracle_setup(lonc, latc, lond[tile], latd[tile], 2)
        sc = xgtest.parent_scale(ref, lonc, latc, lond[tile], latd[tile])
        xgtest.assert_xgrid_equal(got, ref, 2, AREA_RTOL, DIST_ATOL, scale=sc)


def test_mask_and_empty(pkg):
    lonc, latc = pkg.cubed_sphere_grid(12)
    lon2, lat2 = pkg.latlon_grid(60, 30)
    rng = np.random.default_rng(7)
    mask = (rng.random(6 * 12 * 12) > 0.4).astype(np.float64)
    got = _gen(pkg, lonc, latc, lon2, lat2, 1, mask=mask)
    # masked source cells emit nothing (create_xgrid.c:752)
    src = (got["t_in"].astype(np.int64) * 144 + got["j_in"] * 12 + got["i_in"])
    assert np.all(mask[src] > 0.5)
    full = _gen(pkg, lonc, latc, lon2, lat2, 1)
    srcf = (full["t_in"].astype(np.int64) * 144 + full["j_in"] * 12 + full["i_in"])
    keep = mask[srcf] > 0.5
    assert got["nxgrid"] == int(keep.sum())
    assert np.array_equal(got["i_out"], full["i_out"][keep]) and np.array_equal(got["area"], full["area"][keep])
    none = _gen(pkg, lonc, latc, lon2, lat2, 1, mask=np.zeros(6 * 144))
    assert none["nxgrid"] == 0


def test_reference_signature_entry_points(pkg):
    """create_xgrid_2dx2d_order1/2 through the reference's own C signatures, per tile with the
    latitude-trimmed slab exactly as conserve_interp.c:169-200 calls them."""
    lonc, latc = pkg.cubed_sphere_grid(16)
    lon2, lat2 = pkg.latlon_grid(90, 45)
    O = xgtest.oracle_lib()
    for tile in (0, 2):
        n, ii, ji, io, jo, xa, xc, yc = pkg.create_xgrid_2dx2d_order2(lonc[tile], latc[tile], lon2, lat2)
        cap = 200000
        bi = [np.zeros(cap, np.int32) for _ in range(4)]
        bd = [np.zeros(cap) for _ in range(3)]
        m = O.orc_create_xgrid_2dx2d(2, 16, 16, 90, 45, np.ascontiguousarray(lonc[tile]).ravel(), np.ascontiguousarray(latc[tile]).ravel(),
                                     lon2.ravel(), lat2.ravel(), np.ones(256), cap, *bi, bd[0], bd[1].ctypes.data, bd[2].ctypes.data)
        assert n == m
        for a, b in zip((ii, ji, io, jo), bi):
            assert np.array_equal(a, b[:m])
        if xgtest.libm_matches_ref_trig():
            assert np.array_equal(xa, bd[0][:m]) and np.array_equal(xc, bd[1][:m]) and np.array_equal(yc, bd[2][:m])
        else:
            assert np.max(np.abs(xa - bd[0][:m]) / bd[0][:m]) <= 1e-10
    area = pkg.get_grid_area(lon2, lat2)
    oa = np.zeros(90 * 45)
    O.orc_get_grid_area(90, 45, lon2.ravel(), lat2.ravel(), oa)
    if xgtest.libm_matches_ref_trig():
        assert np.array_equal(area.ravel(), oa)
    assert np.max(np.abs(area.ravel() - oa) / oa) <= AREA_RTOL
    assert pkg.get_maxxgrid() == 5000000


def test_windows_concatenate_to_full_result(pkg):
    """multi-GPU sharding unit: source-cell windows, concatenated in order, equal the single-window result."""
    lonc, latc = pkg.cubed_sphere_grid(24)
    lon2, lat2 = pkg.latlon_grid(120, 60)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    plan.generate(2)
    full = plan.result_host()
    bounds = plan.partition(5)
    assert bounds[0] == 0 and bounds[-1] == 6 * 24 * 24 and all(a <= b for a, b in zip(bounds, bounds[1:]))
    parts = []
    for a, b in zip(bounds, bounds[1:]):
        plan.set_src_window(a, b)
        plan.generate(2)
        parts.append(plan.result_host())
    for k in full:
        cat = np.concatenate([p[k] for p in parts])
        assert np.array_equal(cat, full[k]), k


def test_conservation_c192_half_degree(pkg):
    """size-independent property at a size the oracle is not run at: exchange areas tile both grids."""
    lonc, latc = pkg.cubed_sphere_grid(192)
    lon2, lat2 = pkg.latlon_grid(720, 360)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    n = plan.generate(2)
    r = plan.result_host()
    a_src = plan.src_area(); a_dst = plan.dst_area()
    s = np.bincount(r["t_in"].astype(np.int64) * 192 * 192 + r["j_in"] * 192 + r["i_in"], weights=r["area"], minlength=a_src.size)
    d = np.bincount(r["j_out"].astype(np.int64) * 720 + r["i_out"], weights=r["area"], minlength=a_dst.size)
    assert np.max(np.abs(s - a_src) / a_src) < 2e-4      # poly_area's straight-line sides vs great-circle cube edges
    assert np.max(np.abs(d - a_dst) / a_dst) < 2e-4
    assert abs(r["area"].sum() / (4 * np.pi * xgtest.RADIUS ** 2) - 1) < 1e-8
    # sortedness: reference emission order
    key = (r["t_in"].astype(np.int64) * 192 * 192 + r["j_in"] * 192 + r["i_in"]) * (720 * 360) + r["j_out"].astype(np.int64) * 720 + r["i_out"]
    assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("order", [1, 2])
def test_generate_to_host_chunks_equal_one_shot(pkg, order):
    """xgb_plan_generate_to_host: the window generated in pieces with the download overlapped; host arrays and the
    device-resident copy must equal the one-shot result bit for bit, for any number of pieces"""
    lonc, latc = pkg.cubed_sphere_grid(24)
    lon2, lat2 = pkg.latlon_grid(144, 72)
    want = _gen(pkg, lonc, latc, lon2, lat2, order)
    n = want["nxgrid"]
    keys_i = ("t_in", "i_in", "j_in", "i_out", "j_out")
    keys_f = ("area", "di", "dj", "xgrid_clon", "xgrid_clat") if order == 2 else ("area",)
    for nchunks in (1, 3, 8, 50):
        plan = pkg.XgridPlan(0)
        plan.set_dst(lon2, lat2)
        plan.set_src(lonc, latc)
        bufs = {k: np.full(n + 7, -1, np.int32) for k in keys_i}
        bufs.update({k: np.full(n + 7, np.nan) for k in keys_f})
        got_n = plan.generate_to_host(order, bufs, nchunks=nchunks)
        assert got_n == n
        for k in keys_i + keys_f:
            assert np.array_equal(bufs[k][:n], want[k]), (nchunks, k)
            assert np.all(bufs[k][n:] == -1) if k in keys_i else np.all(np.isnan(bufs[k][n:]))
        dev = plan.result_host()
        for k in keys_i + keys_f:
            assert np.array_equal(dev[k], want[k]), (nchunks, k)
        plan.close()
    # too small a capacity is an error, not an overrun
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2); plan.set_src(lonc, latc)
    small = {k: np.zeros(n // 2, np.int32) for k in keys_i}
    small.update({k: np.zeros(n // 2) for k in keys_f})
    with pytest.raises(pkg.XgridError):
        plan.generate_to_host(order, small, nchunks=4)
    plan.close()


def test_ragged_tiles_coarse_on_fine_and_fine_on_coarse(pkg):
    """source tiles of different shapes in one mosaic; a very coarse source on a fine destination (every source cell is a
    'heavy' cell of the candidate search and has hundreds of exchange cells) and the opposite"""
    lon_a, lat_a = pkg.latlon_grid(7, 5, lonbegin=0.0, lonend=140.0, latbegin=-50.0, latend=20.0)
    lon_b, lat_b = pkg.latlon_grid(13, 3, lonbegin=140.0, lonend=360.0, latbegin=-10.0, latend=80.0)
    lon_c, lat_c = pkg.latlon_grid(4, 9, lonbegin=20.0, lonend=60.0, latbegin=20.0, latend=90.0)
    srcs = ([lon_a, lon_b, lon_c], [lat_a, lat_b, lat_c])
    for (nlon, nlat) in ((180, 90), (12, 6)):
        lon2, lat2 = pkg.latlon_grid(nlon, nlat)
        for order in (1, 2):
            got = _gen(pkg, srcs[0], srcs[1], lon2, lat2, order)
            ref = xgtest.oracle_setup(srcs[0], srcs[1], lon2, lat2, order)
            sc = xgtest.parent_scale(ref, srcs[0], srcs[1], lon2, lat2)
            xgtest.assert_xgrid_equal(got, ref, order, AREA_RTOL, DIST_ATOL, scale=sc)
    # cubed sphere C4 on one degree: ~700 exchange cells per source cell
    lonc, latc = pkg.cubed_sphere_grid(4)
    lon2, lat2 = pkg.latlon_grid(360, 180)
    got = _gen(pkg, lonc, latc, lon2, lat2, 2)
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, 2)
    sc = xgtest.parent_scale(ref, lonc, latc, lon2, lat2)
    xgtest.assert_xgrid_equal(got, ref, 2, AREA_RTOL, DIST_ATOL, scale=sc)
    # and C96 on 10 degrees
    lonc, latc = pkg.cubed_sphere_grid(96)
    lon2, lat2 = pkg.latlon_grid(36, 18)
    got = _gen(pkg, lonc, latc, lon2, lat2, 1)
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, 1)
    xgtest.assert_xgrid_equal(got, ref, 1, AREA_RTOL, DIST_ATOL, scale=xgtest.parent_scale(ref, lonc, latc, lon2, lat2))


def test_disjoint_grids_give_empty_list(pkg):
    lon1, lat1 = pkg.latlon_grid(10, 10, lonbegin=10.0, lonend=50.0, latbegin=-40.0, latend=-5.0)
    lon2, lat2 = pkg.latlon_grid(12, 8, lonbegin=100.0, lonend=160.0, latbegin=10.0, latend=60.0)
    for opcode in (1, 2, 1 | xgtest.GREAT_CIRCLE):
        plan = pkg.XgridPlan(0)
        plan.set_dst(lon2, lat2)
        plan.set_src([lon1], [lat1])
        assert plan.generate(opcode) == 0
        r = plan.result_host()
        assert all(v.size == 0 for v in r.values())
        plan.close()
    assert xgtest.oracle_setup([lon1], [lat1], lon2, lat2, 1)["nxgrid"] == 0


_BAND = {}


def _band_job(args):
    jsc, jec, order = args
    lonc, latc, lon2, lat2 = _BAND["grids"]
    return xgtest.ref_band_xgrid(lonc, latc, lon2, lat2, order, jsc, jec)


def _ref_bands(lonc, latc, lon2, lat2, order, bands):
    """the unmodified reference generator on destination row bands, one forked process per band"""
    import multiprocessing as mp
    _BAND["grids"] = (lonc, latc, lon2, lat2)
    with mp.get_context("fork").Pool(min(len(bands), os.cpu_count() or 1)) as pool:
        return pool.map(_band_job, [(a, b, order) for a, b in bands], chunksize=1)


def test_c768_eighth_degree_full_size_properties(pkg):
    """BASELINE configs[3] at full size (the bench workload): C768 -> 2880x1440 order 2.  The oracle needs minutes for it;
    size-independent properties instead: known cell count, reference emission order without duplicates, exchange areas
    tile the sphere and every destination cell, windows concatenate to the whole, tile1_distance is centred."""
    n1, nlon, nlat = 768, 2880, 1440
    lonc, latc = pkg.cubed_sphere_grid(n1)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    n = plan.generate(2)
    assert n == 16673872                                  # = the unmodified reference's whole-problem count (profiles/r02_cpu_whole_c768.json)
    r = plan.result_host()
    a_src = plan.src_area(); a_dst = plan.dst_area()
    s = r["t_in"].astype(np.int64) * n1 * n1 + r["j_in"].astype(np.int64) * n1 + r["i_in"]
    d = r["j_out"].astype(np.int64) * nlon + r["i_out"]
    assert np.all(np.diff(s * (nlon * nlat) + d) > 0)
    assert abs(r["area"].sum() / (4 * np.pi * xgtest.RADIUS ** 2) - 1) < 1e-9
    per_dst = np.bincount(d, weights=r["area"], minlength=a_dst.size)
    assert np.max(np.abs(per_dst - a_dst) / a_dst) < 2e-4
    per_src = np.bincount(s, weights=r["area"], minlength=a_src.size)
    assert np.max(np.abs(per_src - a_src) / a_src) < 2e-4
    # tile1_distance: area-weighted mean over a source cell's exchange cells vanishes where the cell is fully covered
    mx = np.bincount(s, weights=r["area"] * r["di"], minlength=a_src.size) / per_src
    my = np.bincount(s, weights=r["area"] * r["dj"], minlength=a_src.size) / per_src
    assert np.percentile(np.abs(mx), 99) < 1e-9 and np.percentile(np.abs(my), 99) < 1e-9
    # ---- against the compiled reference on destination row bands (the headline config had no oracle comparison):
    # the unmodified create_xgrid_2dx2d_order2 on 16 evenly spaced rows plus the two polar rows, one process per band like
    # a fregrid_parallel rank; the GPU list restricted to those rows must be the same list in the same order, areas and
    # raw centroids bit for bit.  tile1_distance (whole-grid sums per source cell) is then checked for the WHOLE list with
    # the centroid-correction restatement (pinned to the reference on small grids, tests/test_oracle_cpu.py).
    if xgtest.ref_lib() is not None:
        rows = sorted(set([0, nlat - 1] + [int((k + 0.5) * nlat / 16) for k in range(16)]))
        bands = _ref_bands(lonc, latc, lon2, lat2, 2, [(j, j) for j in rows])
        checked = 0
        for j, b in zip(rows, bands):
            m = r["j_out"] == j
            assert int(m.sum()) == b["nxgrid"] > 0, (j, int(m.sum()), b["nxgrid"])
            for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
                assert np.array_equal(r[k][m], b[k]), (j, k)
            if xgtest.libm_matches_ref_trig():
                for k in ("area", "xgrid_clon", "xgrid_clat"):
                    assert np.array_equal(r[k][m], b[k]), (j, k)
            else:
                assert np.max(np.abs(r["area"][m] - b["area"]) / b["area"]) < 1e-9
            checked += b["nxgrid"]
        assert checked > 150000
        di, dj = xgtest.oracle_order2_distance(lonc, latc, r)
        if xgtest.libm_matches_ref_trig():
            assert np.array_equal(di, r["di"]) and np.array_equal(dj, r["dj"])
        else:
            assert np.max(np.abs(di - r["di"])) < 1e-9 and np.max(np.abs(dj - r["dj"])) < 1e-9
    # ---- the WHOLE list against the unmodified reference's whole-problem run (scripts/cpu_whole_c768.py: every destination
    # row on all host cores, 552 s on 8 cores; profiles/r02_cpu_whole_c768.json), row by row: count, cell pairs, and the bit
    # patterns of xgrid_area / xgrid_clon / xgrid_clat as order-independent 64-bit sums per destination row
    gold = np.load(os.path.join(xgtest.GOLDEN_DIR, "c768_rowhash.npz"))
    assert int(gold["count"].sum()) == n
    h = xgtest.row_hashes(r, nlat, n1, nlon)
    assert np.array_equal(h[0], gold["count"]) and np.array_equal(h[1], gold["key"])
    if xgtest.libm_matches_ref_trig():
        assert np.array_equal(h[2], gold["area"]) and np.array_equal(h[3], gold["clon"]) and np.array_equal(h[4], gold["clat"])
    # a checksum of the integer lists that the 2-window run must reproduce
    chk = int(np.bitwise_xor.reduce((s * 1315423911 + d * 2654435761) & 0xffffffffffff))
    bounds = plan.partition(2)
    parts = []
    for a, b in zip(bounds, bounds[1:]):
        plan.set_src_window(a, b)
        plan.generate(2)
        parts.append(plan.result_host())
    plan.close()
    assert sum(p["area"].size for p in parts) == n
    s2 = np.concatenate([p["t_in"].astype(np.int64) * n1 * n1 + p["j_in"].astype(np.int64) * n1 + p["i_in"] for p in parts])
    d2 = np.concatenate([p["j_out"].astype(np.int64) * nlon + p["i_out"] for p in parts])
    assert int(np.bitwise_xor.reduce((s2 * 1315423911 + d2 * 2654435761) & 0xffffffffffff)) == chk
    assert np.array_equal(np.concatenate([p["area"] for p in parts]), r["area"])
    assert np.array_equal(np.concatenate([p["di"] for p in parts]), r["di"])


def test_heavy_work_lists_grow_on_their_own(pkg):
    """C24 on a 1/4 degree grid: every source cell has hundreds of candidates, far beyond the initial capacity of the
    level-synchronous work lists; generation (and the partition's count pass) must grow them and still match the oracle"""
    lonc, latc = pkg.cubed_sphere_grid(24)
    lon2, lat2 = pkg.latlon_grid(1440, 720)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    b = plan.partition(3)
    assert b[0] == 0 and b[-1] == 6 * 24 * 24
    n = plan.generate(1)
    got = plan.result_host(); got["nxgrid"] = n
    plan.close()
    ref = xgtest.oracle_setup(lonc, latc, lon2, lat2, 1)
    xgtest.assert_xgrid_equal(got, ref, 1, AREA_RTOL, DIST_ATOL, scale=xgtest.parent_scale(ref, lonc, latc, lon2, lat2))


def test_interleaved_windows_reassemble_to_serial_list(pkg):
    """xgb_plan_set_src_windows: two 'GPUs' take alternating windows of a 12-way partition; putting the windows' pieces back
    in window order gives the serial list bit for bit (this is how bench.py shards for N > 1)"""
    lonc, latc = pkg.cubed_sphere_grid(24)
    lon2, lat2 = pkg.latlon_grid(120, 60)
    plan = pkg.XgridPlan(0)
    plan.set_dst(lon2, lat2)
    plan.set_src(lonc, latc)
    plan.generate(2)
    full = plan.result_host()
    nparts, world = 12, 2
    b = plan.partition(nparts)
    pieces = {}
    for rank in range(world):
        wins = [(b[w], b[w + 1]) for w in range(rank, nparts, world)]
        plan.set_src_windows(wins)
        n = plan.generate(2)
        r = plan.result_host()
        counts = plan.window_counts()
        assert sum(counts) == n and len(counts) == len(wins)
        off = 0
        for k, w in enumerate(range(rank, nparts, world)):
            pieces[w] = {key: v[off:off + counts[k]] for key, v in r.items()}
            off += counts[k]
    for key in full:
        cat = np.concatenate([pieces[w][key] for w in range(nparts)])
        assert np.array_equal(cat, full[key]), key
    # cost-balanced windows (xgb_plan_partition_shares): equal shares are the plain partition, unequal shares move the cuts
    # to the cumulative shares of the pair count, and the pieces still reassemble to the serial list
    assert plan.partition(nparts, [1.0] * nparts) == b or max(abs(x - y) for x, y in zip(plan.partition(nparts, [1.0] * nparts), b)) <= 1
    shares = [0.5 if w in (0, nparts - 1) else 1.0 for w in range(nparts)]
    plan.set_src_window(0, plan.ncell_src)
    plan.generate(2)
    bs = plan.partition(nparts, shares)
    assert bs[0] == 0 and bs[-1] == plan.ncell_src and all(x <= y for x, y in zip(bs, bs[1:])) and bs != b
    pieces = {}
    for rank in range(world):
        plan.set_src_windows([(bs[w], bs[w + 1]) for w in range(rank, nparts, world)])
        n = plan.generate(2)
        r = plan.result_host()
        counts = plan.window_counts()
        off = 0
        for k, w in enumerate(range(rank, nparts, world)):
            pieces[w] = {key: v[off:off + counts[k]] for key, v in r.items()}
            off += counts[k]
    for key in full:
        assert np.array_equal(np.concatenate([pieces[w][key] for w in range(nparts)]), full[key]), key
    small = sum(len(pieces[w]["area"]) for w in (0, nparts - 1)) / 2.0
    mid = sum(len(pieces[w]["area"]) for w in range(1, nparts - 1)) / (nparts - 2.0)
    assert 0.35 < small / mid < 0.65
    # the host-download path works on several windows too
    wins = [(b[w], b[w + 1]) for w in range(0, nparts, 2)]
    plan.set_src_windows(wins)
    n = plan.generate(2)
    want = plan.result_host()
    bufs = {k: np.zeros(n + 3, v.dtype) for k, v in want.items()}
    assert plan.generate_to_host(2, bufs, nchunks=5) == n
    for k in want:
        assert np.array_equal(bufs[k][:n], want[k]), k
    plan.close()


def _generate(pkg, lonc, latc, lon2, lat2, order, no_rect):
    os.environ["XGB_NO_RECT"] = "1" if no_rect else "0"       # read by xgb_plan_set_dst
    try:
        plan = pkg.XgridPlan(0)
        plan.set_dst(lon2, lat2)
        plan.set_src(lonc, latc)
        n = plan.generate(pkg.CONSERVE_ORDER2 if order == 2 else pkg.CONSERVE_ORDER1)
        out = plan.result_host()
        out["nxgrid"] = n
        plan.close()
        return out
    finally:
        os.environ.pop("XGB_NO_RECT", None)


@pytest.mark.parametrize("order", [1, 2])
def test_separable_destination_search_equals_pyramid_walk_and_oracle(pkg, order):
    """the 1-D row / column candidate search used for separable (lat-lon) destinations against the pyramid walk and the oracle:
    global grids, a window of negative longitudes, a window that wraps past 360 degrees (the 2*pi shifts of create_xgrid.c:786-
    801), unevenly spaced rows and columns, a coarse source on a fine destination (heavy cells) and a destination that is not
    separable at all (one displaced vertex: the check must fall back to the pyramid)"""
    lonc, latc = pkg.cubed_sphere_grid(24)
    rng = np.random.default_rng(5)
    cases = [pkg.latlon_grid(144, 72), pkg.latlon_grid(90, 40, -180.0, 180.0, -80.0, 80.0), pkg.latlon_grid(120, 30, 100.0, 460.0, -30.0, 60.0),
             pkg.latlon_grid(720, 360)]
    xs = np.sort(rng.uniform(0.0, 2 * np.pi, 100)); ys = np.sort(rng.uniform(-0.5 * np.pi, 0.5 * np.pi, 50))
    xs[0], xs[-1], ys[0], ys[-1] = 0.0, 2 * np.pi, -0.5 * np.pi, 0.5 * np.pi
    cases.append(tuple(np.ascontiguousarray(a) for a in np.meshgrid(xs, ys)))
    bent = [a.copy() for a in pkg.latlon_grid(72, 36)]
    bent[0][10, 20] += 1e-3; bent[1][10, 20] += 2e-3
    cases.append(tuple(bent))
    for k, (lon2, lat2) in enumerate(cases):
        a = _generate(pkg, lonc, latc, lon2, lat2, order, no_rect=False)
        b = _generate(pkg, lonc, latc, lon2, lat2, order, no_rect=True)
        assert a["nxgrid"] == b["nxgrid"] and a["nxgrid"] > 0, k
        for key in a:
            if key != "nxgrid":
                assert np.array_equal(a[key], b[key]), (k, key)
        if k in (1, 2, 4):
            want = xgtest.oracle_setup(lonc, latc, lon2, lat2, xgtest.ORDER2 if order == 2 else xgtest.ORDER1)
            sc = xgtest.parent_scale(want, lonc, latc, lon2, lat2)
            xgtest.assert_xgrid_equal(a, want, order, area_tol=1e-12, dist_atol=1e-9, scale=sc)


def test_destination_grid_built_on_the_device_equals_the_uploaded_one(pkg):
    """xgb_plan_set_dst_latlon (fregrid's --nlon/--nlat grid, get_output_grid_by_size) against the same grid uploaded from the
    host: identical exchange grids, bit for bit"""
    lonc, latc = pkg.cubed_sphere_grid(24)
    for args in ((96, 48, 0.0, 360.0, -90.0, 90.0), (50, 40, -30.0, 95.0, -63.0, 77.0)):
        lon2, lat2 = pkg.latlon_grid(*args)
        want = _generate(pkg, lonc, latc, lon2, lat2, 2, no_rect=False)
        plan = pkg.XgridPlan(0)
        plan.set_dst_latlon(*args)
        plan.set_src(lonc, latc)
        n = plan.generate(pkg.CONSERVE_ORDER2)
        got = plan.result_host()
        plan.close()
        assert n == want["nxgrid"] and n > 0
        for key in got:
            assert np.array_equal(got[key], want[key]), key


def test_generate_async_queues_windows_and_finish_returns_the_same_result(pkg):
    """xgb_plan_generate_async / _finish: several windows queued back to back without a host wait leave the same result as the
    blocking call; per-window counts taken on the device equal the host's; a call before any blocking generate is refused"""
    import torch
    lonc, latc = pkg.cubed_sphere_grid(32)
    plan = pkg.XgridPlan(0)
    plan.set_dst_latlon(180, 90)
    plan.set_src(lonc, latc)
    with pytest.raises(pkg.XgridError):
        plan.generate_async(pkg.CONSERVE_ORDER2)
    b = plan.partition(6)
    wins = [(b[0], b[1]), (b[2], b[3]), (b[5], b[6])]
    plan.set_src_windows(wins)
    n = plan.generate(pkg.CONSERVE_ORDER2)
    want = plan.result_host(); wc = plan.window_counts()
    dev_counts = torch.zeros(len(wins), dtype=torch.int64, device="cuda:0")
    for _ in range(4):
        plan.generate_async(pkg.CONSERVE_ORDER2)
        plan.window_counts_device(dev_counts)
    assert plan.generate_finish() == n
    got = plan.result_host()
    torch.cuda.synchronize()
    assert dev_counts.tolist() == wc == plan.window_counts() and sum(wc) == n
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    # a larger window after a small one: the buffers sized for the small one overflow, finish repeats the window by itself
    plan.set_src_window(b[0], b[1])
    small = plan.generate(pkg.CONSERVE_ORDER1)
    plan.set_src_window(b[0], b[6])
    plan.generate_async(pkg.CONSERVE_ORDER1)
    full = plan.generate_finish()
    assert full > small and full == plan.generate(pkg.CONSERVE_ORDER1)
    plan.close()


def test_c3072_thirtysecond_degree_source_slabs_against_the_compiled_reference(pkg):
    """BASELINE configs[4] (the reference's one published figure: C3072 -> 11520x5760, conserve_order1; 2.7e8 exchange cells
    for the whole job).  The whole mosaic and the destination are set up on the device at full size; four slabs of source
    rows (mid-latitude face, rotated face, polar face, rotated face next to its edge) are generated as source windows —
    the code path of a multi-GPU rank — and, on three destination rows each (first, middle and last row the slab reaches),
    compared with the unmodified create_xgrid_2dx2d_order1 run on the same source rows and that destination row: same
    lists, same order, bit-identical areas."""
    if xgtest.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    import ctypes as C
    n1, nlon, nlat = 3072, 11520, 5760
    lonc, latc = pkg.cubed_sphere_grid(n1)
    plan = pkg.XgridPlan(0)
    plan.set_dst_latlon(nlon, nlat)                      # built on the device, bit-identical to latlon_grid (tested elsewhere)
    plan.set_src(lonc, latc)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    R = xgtest.ref_lib()
    R.create_xgrid_2dx2d_order1.restype = C.c_int
    ci = lambda v: C.byref(C.c_int(v))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    total = 0
    for tile, j0, nrows in ((0, 1500, 3), (4, 777, 3), (2, 1530, 2), (5, 40, 3)):
        begin = tile * n1 * n1 + j0 * n1
        plan.set_src_window(begin, begin + nrows * n1)
        n = plan.generate(1)
        g = plan.result_host()
        assert n > 0 and np.all(g["t_in"] == tile) and g["j_in"].min() >= j0 and g["j_in"].max() < j0 + nrows
        lo1 = np.ascontiguousarray(lonc[tile, j0:j0 + nrows + 1]); la1 = np.ascontiguousarray(latc[tile, j0:j0 + nrows + 1])
        ja, jb = int(g["j_out"].min()), int(g["j_out"].max())
        mask = np.ones(n1 * nrows)
        for j in sorted({ja, (ja + jb) // 2, jb}):
            m = g["j_out"] == j
            k = int(m.sum())
            lo2 = np.ascontiguousarray(lon2[j:j + 2]); la2 = np.ascontiguousarray(lat2[j:j + 2])
            cap = 4 * k + 65536
            bi = [np.empty(cap, np.int32) for _ in range(4)]
            xa = np.empty(cap)
            nr = R.create_xgrid_2dx2d_order1(ci(n1), ci(nrows), ci(nlon), ci(1), pv(lo1), pv(la1), pv(lo2), pv(la2), pv(mask),
                                             pv(bi[0]), pv(bi[1]), pv(bi[2]), pv(bi[3]), pv(xa))
            assert nr == k > 0, (tile, j0, j, nr, k)
            assert np.array_equal(g["i_in"][m], bi[0][:k]) and np.array_equal(g["j_in"][m], bi[1][:k] + j0)
            assert np.array_equal(g["i_out"][m], bi[2][:k]) and np.all(bi[3][:k] == 0)
            if xgtest.libm_matches_ref_trig():
                assert np.array_equal(g["area"][m], xa[:k]), (tile, j0, j)
            else:
                assert np.max(np.abs(g["area"][m] - xa[:k]) / xa[:k]) < 1e-9
            total += k
        # rows beyond the slab's reach stay empty in the reference too
        for j in (ja - 1, jb + 1):
            if 0 <= j < nlat:
                lo2 = np.ascontiguousarray(lon2[j:j + 2]); la2 = np.ascontiguousarray(lat2[j:j + 2])
                bi = [np.empty(65536, np.int32) for _ in range(4)]
                xa = np.empty(65536)
                assert R.create_xgrid_2dx2d_order1(ci(n1), ci(nrows), ci(nlon), ci(1), pv(lo1), pv(la1), pv(lo2), pv(la2), pv(mask),
                                                   pv(bi[0]), pv(bi[1]), pv(bi[2]), pv(bi[3]), pv(xa)) == 0
    plan.close()
    assert total > 2000


def test_reference_embedded_polygon_cases_through_the_c_abi(pkg):
    """The reference's own 26 hand-built polygon cases (create_xgrid.c:2383-3010; tests/golden/ref_polycases.npz, recorded from
    the unmodified reference): every quadrilateral pair taken as two 1x1 grids through create_xgrid_2dx2d_order2 and
    create_xgrid_great_circle of the C ABI, and the small-grid great-circle cases 11-13 — poles, sides through the south
    pole, identical boxes, containment.  Counts and lists exact; 2dx2d areas / centroids bit-identical; great-circle areas
    to the great-circle tolerance."""
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "ref_polycases.npz"))
    cell = lambda v: np.array([[v[0], v[1]], [v[3], v[2]]])
    seen = 0
    for n in range(1, 27):
        k = f"c{n:02d}_"
        x1, y1, x2, y2 = g[k + "lon1"], g[k + "lat1"], g[k + "lon2"], g[k + "lat2"]
        if k + "cell_o2" in g:
            want = g[k + "cell_o2"]
            r = pkg.create_xgrid_2dx2d_order2(cell(x1), cell(y1), cell(x2), cell(y2))
            assert r[0] == int(want[0]), n
            got = np.concatenate([r[5], r[6], r[7]])
            if xgtest.libm_matches_ref_trig():
                assert np.array_equal(got, want[1:]), n
            else:
                assert np.allclose(got, want[1:], rtol=1e-9, atol=1.0), n
            want = g[k + "cell_gc"]
            r = pkg.create_xgrid_great_circle(cell(x1), cell(y1), cell(x2), cell(y2))
            assert r[0] == int(want[0]), n
            assert np.max(np.abs(r[5] - want[1:]), initial=0.0) / xgtest.RADIUS ** 2 <= 8e-15, n
            seen += 1
        if 11 <= n <= 13:
            n1, n2, nlon1, nlat1, nlon2, nlat2 = (int(v) for v in g[k + "dims"])
            r = pkg.create_xgrid_great_circle(x1.reshape(nlat1 + 1, nlon1 + 1), y1.reshape(nlat1 + 1, nlon1 + 1),
                                              x2.reshape(nlat2 + 1, nlon2 + 1), y2.reshape(nlat2 + 1, nlon2 + 1))
            assert r[0] == int(g[k + "gcx_n"]), n
            assert np.array_equal(np.stack(r[1:5]), g[k + "gcx_idx"]), n
            assert np.max(np.abs(r[5] - g[k + "gcx_area"]), initial=0.0) / xgtest.RADIUS ** 2 <= 8e-15, n
    assert seen >= 14


def test_1dx2d_and_2dx1d_generators_match_the_reference_golden(pkg):
    """create_xgrid_1dx2d_order1/2 and create_xgrid_2dx1d_order1/2 (create_xgrid.c:208-598: runoff_regrid / interp.c's
    generators between a regular grid given by 1-D bounds and a 2-D grid) through the reference-signature C ABI against
    vectors recorded from the unmodified reference (tests/golden/make_box_golden.py): a pole inside a cube face, the date
    line, a regional pair with masks, a one-column grid (get_grid_area_no_adjust branch) and a tripolar cap.  Lists in the
    reference's emission order, areas and centroid integrals bit-identical."""
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "xgrid_box.npz"))
    exact = xgtest.libm_matches_ref_trig()
    for key in g["names"]:
        bname, gname = str(key).split("__")
        lb, ab = g[f"box_{bname}_lon"], g[f"box_{bname}_lat"]
        lg, ag = g[f"grid_{gname}_lon"], g[f"grid_{gname}_lat"]
        for order in (1, 2):
            for kind in ("1dx2d", "2dx1d"):
                fn = getattr(pkg, f"create_xgrid_{kind}_order{order}")
                if kind == "1dx2d":
                    r = fn(lb, ab, lg, ag, g[f"{key}_mask_box"])
                else:
                    r = fn(lg, ag, lb, ab, g[f"{key}_mask_cell"])
                pre = f"{key}_{kind}_o{order}_"
                assert r[0] == int(g[pre + "n"]), (key, kind, order, r[0], int(g[pre + "n"]))
                assert np.array_equal(np.stack(r[1:5]), g[pre + "idx"]), (key, kind, order)
                fields = [("area", r[5])] + ([("clon", r[6]), ("clat", r[7])] if order == 2 else [])
                for name, got in fields:
                    want = g[pre + name]
                    if exact:
                        assert np.array_equal(got, want), (key, kind, order, name)
                    else:
                        assert np.allclose(got, want, rtol=1e-9, atol=1e-3), (key, kind, order, name)


def test_sharded_source_upload_equals_the_whole_mosaic(pkg):
    """xgb_plan_set_dst_latlon + xgb_plan_set_src_sharded (only the windows' vertex rows uploaded, only their cells precomputed,
    no host synchronisation) generate exactly what xgb_plan_set_src + xgb_plan_set_src_windows generate"""
    ni, nlon, nlat = 48, 360, 180
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    full = pkg.XgridPlan(0)
    full.set_dst(lon2, lat2); full.set_src(lonc, latc)
    b = full.partition(6)
    wins = [(b[1], b[2]), (b[4], b[5])]
    full.set_src_windows(wins)
    n = full.generate(2)
    want = full.result_host(); wc = full.window_counts()
    p = pkg.XgridPlan(0)
    for _ in range(2):                                   # twice: the second call reuses the plan's buffers
        p.set_dst_latlon(nlon, nlat)
        nbytes = p.set_src_sharded([ni] * 6, [ni] * 6, lonc.reshape(-1), latc.reshape(-1), wins)
        assert 0 < nbytes < lonc.size * 16
        assert p.generate(2) == n
        got = p.result_host()
        assert p.window_counts() == wc
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    full.close(); p.close()
