"""Test-side helpers: ctypes access to the CPU oracle (oracle/liboracle_xgrid.so), to the compiled
reference (oracle/_ref/libfrenc_ref.so, when present) and loading of the product package.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

dp = np.ctypeslib.ndpointer(np.float64, flags="C")
ip = np.ctypeslib.ndpointer(np.int32, flags="C")
vp = C.c_void_p

ORDER1, ORDER2, GREAT_CIRCLE, MONOTONIC, LEGACY_CLIP = 1, 2, 4096, 16384, 2048
RADIUS = 6371000.0

_oracle = None
_ref = None


def package():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ge.load_package()


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle_xgrid.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.startswith("xgrid_oracle")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
        L = C.CDLL(path)
        L.orc_fix_lon.argtypes = [dp, dp, C.c_int, C.c_double]
        L.orc_poly_area.restype = C.c_double
        L.orc_poly_area.argtypes = [dp, dp, C.c_int]
        L.orc_poly_ctrlon.restype = C.c_double
        L.orc_poly_ctrlon.argtypes = [dp, dp, C.c_int, C.c_double]
        L.orc_poly_ctrlat.restype = C.c_double
        L.orc_poly_ctrlat.argtypes = [dp, dp, C.c_int]
        L.orc_clip_2dx2d.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, dp, dp]
        L.orc_get_grid_area.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_create_xgrid_2dx2d.restype = C.c_long
        L.orc_create_xgrid_2dx2d.argtypes = [C.c_int] * 5 + [dp] * 5 + [C.c_long] + [ip] * 4 + [dp, vp, vp]
        L.orc_setup_conserve_interp.restype = C.c_long
        L.orc_setup_conserve_interp.argtypes = [C.c_int, ip, ip, dp, dp, C.c_int, C.c_int, dp, dp, C.c_uint, C.c_long] + [ip] * 5 + [dp, vp, vp]
        L.orc_setup_conserve_interp_ex.restype = C.c_long
        L.orc_setup_conserve_interp_ex.argtypes = [C.c_int, ip, ip, dp, dp, C.c_int, C.c_int, dp, dp, C.c_uint, C.c_long] + [ip] * 5 + [dp, vp, vp, vp, vp]
        L.orc_order2_distance.argtypes = [C.c_int, ip, ip, dp, dp, C.c_long, ip, ip, ip, dp, dp, dp, dp, dp]
        L.orc_conserve_apply.argtypes = [C.c_int, C.c_long] + [ip] * 5 + [dp, vp, vp, C.c_int, ip, ip, dp, vp, vp, vp,
                                         C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_libm_trig.argtypes = [C.c_long] + [dp] * 5
        L.orc_conserve_apply_ex.argtypes = [C.c_int, C.c_long] + [ip] * 5 + [dp, vp, vp, C.c_int, ip, ip, dp, vp, vp, vp, C.c_int, C.c_double, C.c_int,
                                            C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, dp]
        L.orc_clip_2dx2d_great_circle.argtypes = [dp, dp, dp, C.c_int, dp, dp, dp, C.c_int, dp, dp, dp]
        L.orc_great_circle_area.restype = C.c_double
        L.orc_great_circle_area.argtypes = [C.c_int, dp, dp, dp]
        L.orc_create_xgrid_great_circle.restype = C.c_long
        L.orc_create_xgrid_great_circle.argtypes = [C.c_int] * 4 + [dp] * 5 + [C.c_long] + [ip] * 4 + [dp, vp, vp]
        L.orc_get_grid_great_circle_area.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_grad_c2l.argtypes = [C.c_int, C.c_int] + [dp] * 14
        L.orc_grad_mask.argtypes = [C.c_int, C.c_int, dp, C.c_double, ip]
        L.orc_calc_c2l_grid_info.argtypes = [C.c_int, C.c_int] + [dp] * 15
        _oracle = L
    return _oracle


def trig_samples(n=200000, seed=99):
    """arguments covering every branch of csrc/ref_trig.cuh: bulk (-2.42, 2.42), Taylor range, tiny, near pi/2, and the
    large-argument range reduction"""
    rng = np.random.default_rng(seed)
    x = np.concatenate([rng.uniform(-2.42, 2.42, n), rng.uniform(-0.13, 0.13, n // 4), rng.uniform(-1e-3, 1e-3, n // 8),
                        rng.uniform(-1e-8, 1e-8, n // 8), np.pi / 2 - rng.uniform(0, 1e-4, n // 8),
                        -np.pi / 2 + rng.uniform(0, 1e-4, n // 8),
                        rng.uniform(-7.0, 7.0, n // 2), rng.uniform(-400.0, 400.0, n // 4),      # longitudes: reduce_sincos branch
                        np.array([0.0, -0.0, np.pi / 2, -np.pi / 2, 0.126, 0.855469, 0.85546875, 2.426265, 2.0 ** -26, 2.0 ** -27,
                                  np.pi, -np.pi, 2 * np.pi, 1.5 * np.pi, 3.0, 4.0, 5.0, 6.0])])
    return np.ascontiguousarray(x)


def libm_trig(x):
    L = oracle_lib()
    out = [np.empty_like(x) for _ in range(4)]
    L.orc_libm_trig(x.size, x, *out)
    return out


_libm_ok = None


def libm_matches_ref_trig():
    """True when the host libm rounds sin/cos/sincos exactly like csrc/ref_trig.cuh (glibc 2.39 x86-64 FMA build):
    then the GPU results are expected to equal the oracle's bit for bit, areas and centroids included."""
    global _libm_ok
    if _libm_ok is None:
        pk = package()
        x = trig_samples(50000)
        got = [np.empty_like(x) for _ in range(4)]
        pk.lib().xgb_ref_trig_host(x.size, *[a.ctypes.data for a in [x] + got])
        want = libm_trig(x)
        _libm_ok = all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(got, want))
    return _libm_ok


def ref_lib():
    """The unmodified reference compiled by oracle/Makefile, or None when it was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libfrenc_ref.so")
        if os.path.isdir("/root/reference/tools/libfrencutils"):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_cubed_sphere_grid.argtypes = [C.c_int, dp, dp, vp, vp]
        L.ref_tripolar_grid.argtypes = [C.c_int, C.c_int] + [C.c_double] * 5 + [dp, dp]
        L.ref_regrid_setup.restype = vp
        L.ref_regrid_setup.argtypes = [C.c_int, ip, ip, dp, dp, vp, vp, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, C.c_uint]
        L.ref_regrid_nxgrid.restype = C.c_long
        L.ref_regrid_nxgrid.argtypes = [vp]
        L.ref_regrid_get.argtypes = [vp] + [ip] * 5 + [dp, vp, vp]
        L.ref_regrid_cell_area.argtypes = [vp, vp, vp]
        L.ref_regrid_apply.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint, dp, vp, vp, vp, dp]
        L.ref_regrid_free.argtypes = [vp]
        L.ref_regrid_apply_ex.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, vp, vp, C.c_double, C.c_uint, dp, vp, vp, vp, dp]
        L.ref_compute_extent.argtypes = [C.c_int, C.c_int, ip, ip]
        L.ref_abi_layout.argtypes = [C.POINTER(C.c_size_t), C.c_int]
        L.ref_next_setup_remap.argtypes = [C.c_char_p, C.c_int]
        L.ref_regrid_setup_through_remap.restype = vp
        L.ref_regrid_setup_through_remap.argtypes = [vp, vp, C.c_char_p, C.c_int]
        for name in ("stub_dim_name", "stub_var_name", "stub_var_att_name", "stub_var_att_value"):
            getattr(L, name).restype = C.c_char_p
        L.stub_find.argtypes = [C.c_char_p]
        L.stub_new_file.argtypes = [C.c_char_p]
        L.stub_add_dim.argtypes = [C.c_int, C.c_char_p, C.c_long]
        L.stub_add_var.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, ip]
        L.stub_set_var_data.argtypes = [C.c_int, C.c_int, vp]
        L.stub_var_data.restype = vp
        L.stub_var_nelem.restype = C.c_long
        L.stub_dim_size.restype = C.c_long
        L.ref_regrid_setup_through.restype = vp
        L.ref_regrid_setup_through.argtypes = [vp, vp]
        L.ref_multi_setup.restype = vp
        L.ref_multi_setup.argtypes = [C.c_int, ip, ip, dp, dp, C.c_int, ip, ip, dp, dp, C.c_uint, vp]
        L.ref_multi_nxgrid.restype = C.c_long
        L.ref_multi_nxgrid.argtypes = [vp, C.c_int]
        L.ref_multi_get.argtypes = [vp, C.c_int] + [ip] * 5 + [dp, vp, vp]
        L.ref_multi_apply.argtypes = [vp, vp, C.c_int, dp, vp, vp, vp, dp]
        L.ref_regrid_apply_through.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint, dp, vp, vp, vp, dp]
        L.fix_lon.argtypes = [dp, dp, C.c_int, C.c_double]
        L.poly_area.restype = C.c_double
        L.poly_area.argtypes = [dp, dp, C.c_int]
        L.poly_ctrlon.restype = C.c_double
        L.poly_ctrlon.argtypes = [dp, dp, C.c_int, C.c_double]
        L.poly_ctrlat.restype = C.c_double
        L.poly_ctrlat.argtypes = [dp, dp, C.c_int]
        L.clip_2dx2d.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, dp, dp]
        L.get_grid_area.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp, dp]
        pi_ = C.POINTER(C.c_int)
        L.clip_2dx2d_great_circle.argtypes = [dp, dp, dp, C.c_int, dp, dp, dp, C.c_int, dp, dp, dp]
        L.great_circle_area.restype = C.c_double
        L.great_circle_area.argtypes = [C.c_int, dp, dp, dp]
        L.get_grid_great_circle_area.argtypes = [pi_, pi_, dp, dp, dp]
        L.grad_c2l.argtypes = [pi_, pi_] + [dp] * 14 + [pi_] * 4
        L.calc_c2l_grid_info.argtypes = [pi_, pi_] + [dp] * 15 + [pi_] * 4
        _ref = L
    return _ref


def _tiles(lonc, latc):
    """[ntiles, ny+1, nx+1] array or list of 2-D arrays -> (nx[], ny[], lon_cat, lat_cat)."""
    if not isinstance(lonc, (list, tuple)):
        lonc = [lonc[t] for t in range(lonc.shape[0])] if lonc.ndim == 3 else [lonc]
        latc = [latc[t] for t in range(latc.shape[0])] if latc.ndim == 3 else [latc]
    nx = np.array([a.shape[1] - 1 for a in lonc], np.int32)
    ny = np.array([a.shape[0] - 1 for a in lonc], np.int32)
    lon = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float64).ravel() for a in lonc]))
    lat = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float64).ravel() for a in latc]))
    return nx, ny, lon, lat


def _alloc(cap, order):
    out = {k: np.zeros(cap, np.int32) for k in ("t_in", "i_in", "j_in", "i_out", "j_out")}
    out["area"] = np.zeros(cap)
    if order == 2:
        out["di"] = np.zeros(cap)
        out["dj"] = np.zeros(cap)
    return out


def _trim(out, n):
    res = {k: v[:n].copy() for k, v in out.items()}
    res["nxgrid"] = int(n)
    return res


def oracle_setup(lonc, latc, lon2, lat2, opcode, cap=None, raw=False):
    """oracle restatement of setup_conserve_interp for a source mosaic and one destination tile.
    raw=True (order 2): also the generators' raw xgrid_clon / xgrid_clat."""
    L = oracle_lib()
    nx, ny, lon, lat = _tiles(lonc, latc)
    lon2 = np.ascontiguousarray(lon2, np.float64); lat2 = np.ascontiguousarray(lat2, np.float64)
    ny2, nx2 = lon2.shape[0] - 1, lon2.shape[1] - 1
    order = 2 if opcode & ORDER2 else 1
    if cap is None:
        cap = int(12 * max(int((nx * ny).sum()), nx2 * ny2)) + 4096
    out = _alloc(cap, order)
    if raw and order == 2:
        out["xgrid_clon"] = np.empty(cap); out["xgrid_clat"] = np.empty(cap)
    n = L.orc_setup_conserve_interp_ex(len(nx), nx, ny, lon, lat, nx2, ny2, lon2.ravel(), lat2.ravel(), opcode, cap,
                                       out["t_in"], out["i_in"], out["j_in"], out["i_out"], out["j_out"], out["area"],
                                       out["di"].ctypes.data if order == 2 else None,
                                       out["dj"].ctypes.data if order == 2 else None,
                                       out["xgrid_clon"].ctypes.data if "xgrid_clon" in out else None,
                                       out["xgrid_clat"].ctypes.data if "xgrid_clat" in out else None)
    if n < 0:
        raise RuntimeError("oracle capacity exceeded")
    return _trim(out, n)


def oracle_order2_distance(lonc, latc, x):
    """tile1_distance (di, dj) of a finished one-output-tile list from its areas and raw centroids: the order-2 centroid
    correction alone (conserve_interp.c:204-221, :319-358), oracle/xgrid_oracle.c orc_order2_distance"""
    L = oracle_lib()
    nx, ny, lon, lat = _tiles(lonc, latc)
    n = int(x["area"].shape[0])
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    di = np.empty(n); dj = np.empty(n)
    L.orc_order2_distance(len(nx), nx, ny, lon, lat, n, i32(x["t_in"]), i32(x["i_in"]), i32(x["j_in"]), f64(x["area"]),
                          f64(x["xgrid_clon"]), f64(x["xgrid_clat"]), di, dj)
    return di, dj


def ref_band_xgrid(lonc, latc, lon2, lat2, order, jsc, jec):
    """The UNMODIFIED reference generators create_xgrid_2dx2d_order1/2 (create_xgrid.c:621, :893) on destination rows
    [jsc, jec] of a lat-lon grid, one call per source tile with setup_conserve_interp's latitude trim (conserve_interp.c:
    169-200) — what a fregrid_parallel rank owning that band computes, but with the raw xgrid_clon / xgrid_clat kept
    (setup_conserve_interp's own di/dj of a band differ from the whole-grid run: its per-cell sums only see the band).
    -> dict t_in, i_in, j_in, i_out, j_out (GLOBAL row), area[, xgrid_clon, xgrid_clat]"""
    R = ref_lib()
    nxs, nys, _, _ = _tiles(lonc, latc)
    lo2 = np.ascontiguousarray(lon2[jsc:jec + 2]); la2 = np.ascontiguousarray(lat2[jsc:jec + 2])
    ny2, nx2 = lo2.shape[0] - 1, lo2.shape[1] - 1
    y_min, y_max = la2.min(), la2.max()
    cap = 64 * nx2 * ny2 + 65536
    bufs = [np.empty(cap, np.int32) for _ in range(4)]
    xa = np.empty(cap); xc = np.empty(cap); yc = np.empty(cap)
    ci = lambda v: C.byref(C.c_int(v))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    parts = {k: [] for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "xgrid_clon", "xgrid_clat")}
    for m in range(len(nxs)):
        lo1 = np.ascontiguousarray(lonc[m], np.float64); la1 = np.ascontiguousarray(latc[m], np.float64)
        ny1, nx1 = lo1.shape[0] - 1, lo1.shape[1] - 1
        rows_gt = np.nonzero((la1 > y_min).any(axis=1))[0]
        rows_lt = np.nonzero((la1 < y_max).any(axis=1))[0]
        jstart = max(0, (rows_gt.min() if rows_gt.size else ny1) - 1)
        jend = min(ny1 - 1, (rows_lt.max() if rows_lt.size else -1) + 1)
        ny_now = jend - jstart + 1
        if ny_now <= 0:
            continue
        mask = np.ones(nx1 * ny_now)
        a_lo = np.ascontiguousarray(lo1[jstart:jstart + ny_now + 1]); a_la = np.ascontiguousarray(la1[jstart:jstart + ny_now + 1])
        if order == 2:
            R.create_xgrid_2dx2d_order2.restype = C.c_int
            n = R.create_xgrid_2dx2d_order2(ci(nx1), ci(ny_now), ci(nx2), ci(ny2), pv(a_lo), pv(a_la), pv(lo2), pv(la2), pv(mask),
                                            pv(bufs[0]), pv(bufs[1]), pv(bufs[2]), pv(bufs[3]), pv(xa), pv(xc), pv(yc))
        else:
            R.create_xgrid_2dx2d_order1.restype = C.c_int
            n = R.create_xgrid_2dx2d_order1(ci(nx1), ci(ny_now), ci(nx2), ci(ny2), pv(a_lo), pv(a_la), pv(lo2), pv(la2), pv(mask),
                                            pv(bufs[0]), pv(bufs[1]), pv(bufs[2]), pv(bufs[3]), pv(xa))
        assert 0 <= n <= cap
        parts["t_in"].append(np.full(n, m, np.int32)); parts["i_in"].append(bufs[0][:n].copy())
        parts["j_in"].append(bufs[1][:n] + jstart); parts["i_out"].append(bufs[2][:n].copy())
        parts["j_out"].append(bufs[3][:n] + jsc); parts["area"].append(xa[:n].copy())
        if order == 2:
            parts["xgrid_clon"].append(xc[:n].copy()); parts["xgrid_clat"].append(yc[:n].copy())
    out = {k: (np.concatenate(v) if v else np.empty(0)) for k, v in parts.items() if v or k in ("area",)}
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        out[k] = out.get(k, np.empty(0, np.int32)).astype(np.int32)
    out["nxgrid"] = int(out["area"].shape[0])
    return out


def ref_setup(lonc, latc, lon2, lat2, opcode, jsc=None, jec=None, keep=False, remap=None):
    """the reference's own setup_conserve_interp (conserve_interp.c:42) through oracle/ref_driver.c"""
    L = ref_lib()
    nx, ny, lon, lat = _tiles(lonc, latc)
    lon2 = np.ascontiguousarray(lon2, np.float64); lat2 = np.ascontiguousarray(lat2, np.float64)
    ny2, nx2 = lon2.shape[0] - 1, lon2.shape[1] - 1
    order = 2 if opcode & ORDER2 else 1
    jsc = 0 if jsc is None else jsc
    jec = ny2 - 1 if jec is None else jec
    if not (opcode & GREAT_CIRCLE):
        opcode |= LEGACY_CLIP
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)   # reference prints a NOTE
    if remap is not None:      # (name, 1 = WRITE | 2 = READ) against the in-memory file store of oracle/shim/io_stubs.c
        L.ref_next_setup_remap(remap[0].encode(), remap[1])
    try:
        r = L.ref_regrid_setup(len(nx), nx, ny, lon, lat, None, None, nx2, ny2, lon2.ravel(), lat2.ravel(), jsc, jec, opcode)
    finally:
        os.dup2(saved, 1); os.close(saved); os.close(devnull)
    n = L.ref_regrid_nxgrid(r)
    out = _alloc(max(n, 1), order)
    L.ref_regrid_get(r, out["t_in"], out["i_in"], out["j_in"], out["i_out"], out["j_out"], out["area"],
                     out["di"].ctypes.data if order == 2 else None, out["dj"].ctypes.data if order == 2 else None)
    res = _trim(out, n)
    if keep:
        res["handle"] = r
    else:
        L.ref_regrid_free(r)
    return res


def ref_multi_setup(lonc, latc, lonc_out, latc_out, opcode, setup_fn=None):
    """setup_conserve_interp over SEVERAL output tiles (oracle/ref_driver.c ref_multi_setup): the reference's own when
    setup_fn is None, else the implementation behind the function pointer.  -> (handle, [per-output-tile dict])"""
    L = ref_lib()
    nx, ny, lon, lat = _tiles(lonc, latc)
    nxo, nyo, lono, lato = _tiles(lonc_out, latc_out)
    order = 2 if opcode & ORDER2 else 1
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)   # the NOTE line
    try:
        h = L.ref_multi_setup(len(nx), nx, ny, lon, lat, len(nxo), nxo, nyo, lono, lato, opcode | LEGACY_CLIP, setup_fn)
    finally:
        os.dup2(saved, 1); os.close(saved); os.close(devnull)
    res = []
    for n in range(len(nxo)):
        k = L.ref_multi_nxgrid(h, n)
        out = _alloc(max(k, 1), order)
        if k:
            L.ref_multi_get(h, n, out["t_in"], out["i_in"], out["j_in"], out["i_out"], out["j_out"], out["area"],
                            out["di"].ctypes.data if order == 2 else None, out["dj"].ctypes.data if order == 2 else None)
        res.append(_trim(out, k))
    return h, res


def latlon_grid_np(nlon, nlat, lonbegin=0.0, lonend=360.0, latbegin=-90.0, latend=90.0):
    """fregrid's --nlon/--nlat output grid (get_output_grid_by_size, fregrid_util.c:588-603) in plain Python/numpy — the
    CPU legs (bench.py --impl reference, scripts/cpu_whole_c768.py) build their input without loading the product library;
    bit-identical to pkg.latlon_grid (tests/test_capi_cpu.py)"""
    d2r = np.pi / 180
    dlon = (lonend - lonbegin) / nlon; dlat = (latend - latbegin) / nlat
    lon1d = np.array([(lonbegin + i * dlon) * d2r for i in range(nlon + 1)])
    lat1d = np.array([(latbegin + j * dlat) * d2r for j in range(nlat + 1)])
    return (np.ascontiguousarray(np.broadcast_to(lon1d, (nlat + 1, nlon + 1))),
            np.ascontiguousarray(np.broadcast_to(lat1d[:, None], (nlat + 1, nlon + 1))))


def row_hashes(x, nlat, n1, nlon):
    """per destination row of an exchange-grid list: count and order-independent 64-bit sums (wrapping) of the cell-pair
    keys and of the bit patterns of xgrid_area / xgrid_clon / xgrid_clat -> uint64 [5, nlat].  Two lists with equal row hashes
    hold the same cells with the same floating-point fields (tests/golden/c768_rowhash.npz: the unmodified reference's
    whole C768 -> 1/8 degree list, scripts/cpu_whole_c768.py)."""
    j = x["j_out"].astype(np.int64)
    s = (x["t_in"].astype(np.uint64) * np.uint64(n1 * n1) + x["j_in"].astype(np.uint64) * np.uint64(n1) + x["i_in"].astype(np.uint64))
    d = (x["j_out"].astype(np.uint64) * np.uint64(nlon) + x["i_out"].astype(np.uint64))
    out = np.zeros((5, nlat), np.uint64)
    with np.errstate(over="ignore"):
        key = s * np.uint64(1315423911) + d * np.uint64(2654435761)
        np.add.at(out[0], j, np.uint64(1))
        np.add.at(out[1], j, key)
        for r, k in ((2, "area"), (3, "xgrid_clon"), (4, "xgrid_clat")):
            if k in x:
                np.add.at(out[r], j, np.ascontiguousarray(x[k]).view(np.uint64))
    return out


def ref_cubed_sphere(ni, centers=False):
    L = ref_lib()
    lonc = np.zeros((6, ni + 1, ni + 1)); latc = np.zeros_like(lonc)
    lont = np.zeros((6, ni, ni)); latt = np.zeros_like(lont)
    rc = L.ref_cubed_sphere_grid(ni, lonc.reshape(-1), latc.reshape(-1),
                                 lont.ctypes.data if centers else None, latt.ctypes.data if centers else None)
    assert rc == 0
    return (lonc, latc, lont, latt) if centers else (lonc, latc)


def canonical_order(x):
    """sort key of an exchange-grid list: (tile, j_in, i_in, j_out, i_out)"""
    return np.lexsort((x["i_out"], x["j_out"], x["i_in"], x["j_in"], x["t_in"]))


def parent_scale(x, lonc, latc, lon2, lat2):
    """min(area of parent source cell, area of parent destination cell) per exchange cell: the
    reference's own normalisation of xgrid_area in its accept test (create_xgrid.c:806-807)."""
    L = oracle_lib()
    nx, ny, lon, lat = _tiles(lonc, latc)
    lon2 = np.ascontiguousarray(lon2, np.float64); lat2 = np.ascontiguousarray(lat2, np.float64)
    ny2, nx2 = lon2.shape[0] - 1, lon2.shape[1] - 1
    a_src, coff, voff, offs = [], 0, 0, []
    for t in range(len(nx)):
        nv = (nx[t] + 1) * (ny[t] + 1)
        a = np.zeros(nx[t] * ny[t])
        L.orc_get_grid_area(int(nx[t]), int(ny[t]), lon[voff:voff + nv].copy(), lat[voff:voff + nv].copy(), a)
        a_src.append(a); offs.append(coff)
        coff += nx[t] * ny[t]; voff += nv
    a_src = np.concatenate(a_src); offs = np.array(offs, np.int64)
    a_dst = np.zeros(nx2 * ny2)
    L.orc_get_grid_area(nx2, ny2, lon2.ravel(), lat2.ravel(), a_dst)
    s = offs[x["t_in"]] + x["j_in"].astype(np.int64) * nx[x["t_in"]] + x["i_in"]
    d = x["j_out"].astype(np.int64) * nx2 + x["i_out"]
    return np.minimum(a_src[s], a_dst[d])


def assert_xgrid_equal(got, ref, order, area_tol=1e-12, dist_atol=1e-11, same_order=True, scale=None, exact=None):
    """Integer lists bit-exact (optionally after canonical sort).
    exact=True (default: whenever the host libm rounds like csrc/ref_trig.cuh): xgrid_area and
    tile1_distance must equal the reference BIT FOR BIT as well.  Otherwise:
    xgrid_area: |got - ref| <= area_tol * scale, where scale is the parent-cell area (parent_scale) when
    given, else the reference value itself.  The line integral in poly_area cancels catastrophically for
    sliver cells (terms ~ dlon*sin(lat) against a result ~ ratio*dlon*dlat*cos(lat)), so a 1-ulp libm
    difference is amplified by up to 1/ratio >= 1e6 relative to the sliver's own area; relative to the
    parent cell it stays at rounding level.  See DESIGN.md "Parity tolerances".
    tile1_distance: dist_atol absolute (radians; a difference of near-equal centroids)."""
    assert got["nxgrid"] == ref["nxgrid"], (got["nxgrid"], ref["nxgrid"])
    if same_order:
        pg = pr = slice(None)
    else:
        pg, pr = canonical_order(got), canonical_order(ref)
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(got[k][pg], ref[k][pr]), f"{k} differs"
    a, b = got["area"][pg], ref["area"][pr]
    if exact is None:
        exact = libm_matches_ref_trig()
    if exact:
        assert np.array_equal(a, b), f"xgrid_area not bit-identical: max rel {np.max(np.abs(a - b) / b)}"
        if order == 2:
            for k in ("di", "dj"):
                assert np.array_equal(got[k][pg], ref[k][pr]), f"{k} not bit-identical: max abs {np.max(np.abs(got[k][pg] - ref[k][pr]))}"
        return 0.0
    den = b if scale is None else scale[pr]
    rel = float(np.max(np.abs(a - b) / den)) if len(b) else 0.0
    assert rel <= area_tol, f"xgrid_area max scaled diff {rel}"
    tot = abs(a.sum() - b.sum()) / b.sum() if len(b) else 0.0
    assert tot <= 1e-13, f"sum(xgrid_area) rel diff {tot}"
    if order == 2:
        for k in ("di", "dj"):
            d = np.max(np.abs(got[k][pg] - ref[k][pr])) if len(b) else 0.0
            assert d <= dist_atol, f"{k} max abs diff {d}"
    return rel


# ---------------------------------------------------------------------------------------------
# apply path helpers: fields, halos, gradient metrics, oracle / reference apply
# ---------------------------------------------------------------------------------------------
METRICS = ("dx", "dy", "area", "edge_w", "edge_e", "edge_s", "edge_n", "en_n", "en_e", "vlon", "vlat")


def metric_sizes(nx, ny):
    return {"dx": nx * (ny + 1), "dy": (nx + 1) * ny, "area": nx * ny, "edge_w": ny + 1, "edge_e": ny + 1, "edge_s": nx + 1,
            "edge_n": nx + 1, "en_n": 3 * nx * (ny + 1), "en_e": 3 * (nx + 1) * ny, "vlon": 3 * nx * ny, "vlat": 3 * nx * ny}


def cubed_sphere_halo_map(lonc, latc):
    """Index map that fills the one-cell halo of every tile of a cubed sphere from the neighbouring tiles
    (what fregrid's update_halo, fregrid_util.c:2614, does from the mosaic contact list): returns an int64
    array [6, n+2, n+2] of indices into the flat [6*n*n] cell array, -1 at the four halo corners.
    Neighbours are found geometrically: the cell across an edge is the one sharing its two corner vertices."""
    nt, npt, _ = lonc.shape
    n = npt - 1
    xyz = np.stack([np.cos(latc) * np.cos(lonc), np.cos(latc) * np.sin(lonc), np.sin(latc)], -1)
    key = lambda v: tuple(np.round(v, 9) + 0.0)
    edges = {}

    def add(t, i, j, a, b):
        k = frozenset((key(xyz[t][a]), key(xyz[t][b])))
        edges.setdefault(k, []).append((t, i, j))
    for t in range(nt):
        for k in range(n):
            add(t, 0, k, (k, 0), (k + 1, 0)); add(t, n - 1, k, (k, n), (k + 1, n))
            add(t, k, 0, (0, k), (0, k + 1)); add(t, k, n - 1, (n, k), (n, k + 1))
    m = -np.ones((nt, n + 2, n + 2), np.int64)
    for t in range(nt):
        m[t, 1:-1, 1:-1] = t * n * n + np.arange(n * n).reshape(n, n)

    def other(t, a, b):
        k = frozenset((key(xyz[t][a]), key(xyz[t][b])))
        c = [e for e in edges[k] if e[0] != t]
        assert len(c) == 1, (t, a, b, edges[k])
        tt, ii, jj = c[0]
        return tt * n * n + jj * n + ii
    for t in range(nt):
        for k in range(n):
            m[t, k + 1, 0] = other(t, (k, 0), (k + 1, 0))
            m[t, k + 1, n + 1] = other(t, (k, n), (k + 1, n))
            m[t, 0, k + 1] = other(t, (0, k), (0, k + 1))
            m[t, n + 1, k + 1] = other(t, (n, k), (n, k + 1))
    return m


def with_halo(flat, hmap, corner=0.0):
    """flat [..., ncell] -> [..., 6*(n+2)^2] using cubed_sphere_halo_map; halo corners get `corner`"""
    src = np.concatenate([np.asarray(flat, np.float64), np.full(flat.shape[:-1] + (1,), corner)], -1)
    idx = np.where(hmap < 0, src.shape[-1] - 1, hmap).reshape(-1)
    return np.ascontiguousarray(src[..., idx])


def smooth_field(lont, latt, k=0, t=0):
    """f = 2 + cos^2(lat) cos(2 lon) + 0.01 k + 0.001 t at cell centres (flat)"""
    return (2.0 + np.cos(latt) ** 2 * np.cos(2 * lont) + 0.01 * k + 0.001 * t).reshape(-1)


def c2l_metrics(fn_kind, nx, ny, xt, yt, xc, yc):
    """calc_c2l_grid_info through the oracle ('oracle') or the compiled reference ('ref') for one tile"""
    sz = metric_sizes(nx, ny)
    m = {k: np.zeros(sz[k]) for k in METRICS}
    args = [np.ascontiguousarray(a, np.float64).reshape(-1) for a in (xt, yt, xc, yc)] + [m[k] for k in METRICS]
    if fn_kind == "oracle":
        oracle_lib().orc_calc_c2l_grid_info(nx, ny, *args)
    else:
        one = C.c_int(1)
        ref_lib().calc_c2l_grid_info(C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), *args, *[C.byref(one)] * 4)
    return m


def grad_c2l(fn_kind, nx, ny, pin, m):
    gx = np.zeros(nx * ny); gy = np.zeros(nx * ny)
    args = [np.ascontiguousarray(pin, np.float64).reshape(-1)] + [m[k] for k in METRICS] + [gx, gy]
    if fn_kind == "oracle":
        oracle_lib().orc_grad_c2l(nx, ny, *args)
    else:
        one = C.c_int(1)
        ref_lib().grad_c2l(C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), *args, *[C.byref(one)] * 4)
    return gx, gy


def grad_mask(nx, ny, pin, missing):
    m = np.zeros(nx * ny, np.int32)
    oracle_lib().orc_grad_mask(nx, ny, np.ascontiguousarray(pin, np.float64).reshape(-1), float(missing), m)
    return m


def oracle_apply(x, order, tiles, data, nx_out, ny_out, grad_x=None, grad_y=None, gmask=None, has_missing=False,
                 missing=0.0, monotonic=False):
    """orc_conserve_apply for ONE field-level (nz = 1); tiles = [(nx, ny), ...]"""
    L = oracle_lib()
    nx = np.array([t[0] for t in tiles], np.int32); ny = np.array([t[1] for t in tiles], np.int32)
    out = np.zeros(nx_out * ny_out)
    ptr = lambda a: None if a is None else a.ctypes.data
    gx = None if grad_x is None else np.ascontiguousarray(grad_x, np.float64)
    gy = None if grad_y is None else np.ascontiguousarray(grad_y, np.float64)
    gm = None if gmask is None else np.ascontiguousarray(gmask, np.int32)
    di = x.get("di"); dj = x.get("dj")
    L.orc_conserve_apply(order, x["area"].size, x["t_in"], x["i_in"], x["j_in"], x["i_out"], x["j_out"], x["area"],
                         ptr(di) if order == 2 else None, ptr(dj) if order == 2 else None, len(tiles), nx, ny,
                         np.ascontiguousarray(data, np.float64), ptr(gx), ptr(gy), ptr(gm), int(has_missing), float(missing),
                         int(monotonic), nx_out, ny_out, 1, out)
    return out


def ref_apply(handle, order, data, nout, grad_x=None, grad_y=None, gmask=None, has_missing=False, missing=0.0, monotonic=False):
    """the reference's do_scalar_conserve_interp for ONE field-level through oracle/ref_driver.c"""
    L = ref_lib()
    out = np.zeros(nout)
    ptr = lambda a: None if a is None else a.ctypes.data
    gx = None if grad_x is None else np.ascontiguousarray(grad_x, np.float64)
    gy = None if grad_y is None else np.ascontiguousarray(grad_y, np.float64)
    gm = None if gmask is None else np.ascontiguousarray(gmask, np.int32)
    L.ref_regrid_apply(handle, order, int(has_missing), float(missing), 0, 1, MONOTONIC if monotonic else 0,
                       np.ascontiguousarray(data, np.float64), ptr(gx), ptr(gy), ptr(gm), out)
    return out


# ---------------------------------------------------------------------------------------------
# great-circle path helpers
# ---------------------------------------------------------------------------------------------
def ll2xyz(lon, lat):
    return (np.ascontiguousarray(np.cos(lat) * np.cos(lon)), np.ascontiguousarray(np.cos(lat) * np.sin(lon)),
            np.ascontiguousarray(np.sin(lat)))


def gc_quad_cases(n, seed=1):
    """pairs of spherical quads in the reference's clockwise corner order: generic overlaps, shared edges with vertices
    lying on the other cell's edges, nested cells sharing a corner, identical cells, pole triangles"""
    rng = np.random.default_rng(seed)
    d2r = np.pi / 180

    def quad(lon0, lat0, dlon, dlat, jit):
        lon = np.array([lon0, lon0, lon0 + dlon, lon0 + dlon]) + rng.normal(0, jit, 4)
        lat = np.clip(np.array([lat0, lat0 + dlat, lat0 + dlat, lat0]) + rng.normal(0, jit, 4), -90, 90)
        return lon * d2r, lat * d2r
    out = []
    for it in range(n):
        lon0 = rng.uniform(0, 355); lat0 = rng.uniform(-88, 84); d = rng.uniform(0.1, 3)
        kind = it % 5
        if kind == 0:
            a = quad(lon0, lat0, d, d, 0.05 * d)
            b = quad(lon0 + rng.uniform(-d, d), lat0 + rng.uniform(-d, d), d * rng.uniform(.5, 2), d * rng.uniform(.5, 2), 0)
        elif kind == 1:
            a = quad(lon0, lat0, d, d, 0); b = quad(lon0 + d / 2, lat0, d, d, 0)
        elif kind == 2:
            a = quad(lon0, lat0, d, d, 0); b = quad(lon0, lat0, d / 2, d / 2, 0)
        elif kind == 3:
            a = quad(lon0, lat0, d, d, 0); b = quad(lon0, lat0, d, d, 0)
        else:                                               # pole row of a lat-lon grid (two corners coincide) vs a polar quad
            a = quad(lon0, 90 - d, 2 * d, d, 0); b = quad(lon0 + d / 3, 90 - 1.5 * d, d, d, 0.02 * d)
        out.append((ll2xyz(*a), ll2xyz(*b)))
    return out


def tripolar_grid(nx_s, ny_s, lat_join=65.0):
    """make_hgrid tripolar grid through the compiled reference: supergrid nx_s x ny_s -> (lonc, latc) [ny_s/2+1, nx_s/2+1]"""
    lonc = np.zeros((ny_s // 2 + 1, nx_s // 2 + 1)); latc = np.zeros_like(lonc)
    ref_lib().ref_tripolar_grid(nx_s, ny_s, -280.0, 80.0, -82.0, 90.0, lat_join, lonc.reshape(-1), latc.reshape(-1))
    return lonc, latc


TARGET = 16          # globals.h:50


def oracle_apply_ex(x, order, tiles, data, nx_out, ny_out, grad_x=None, grad_y=None, gmask=None, has_missing=False, missing=0.0,
                    monotonic=False, cell_methods=0, weight=None, cell_area=None, farea=None, target=False, dst_cell_area=None):
    """orc_conserve_apply_ex for one field-level"""
    L = oracle_lib()
    nx = np.array([t[0] for t in tiles], np.int32); ny = np.array([t[1] for t in tiles], np.int32)
    out = np.zeros(nx_out * ny_out)
    keep = []

    def ptr(a, dt=np.float64):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt); keep.append(a)
        return a.ctypes.data
    L.orc_conserve_apply_ex(order, x["area"].size, x["t_in"], x["i_in"], x["j_in"], x["i_out"], x["j_out"], x["area"],
                            ptr(x.get("di")) if order == 2 else None, ptr(x.get("dj")) if order == 2 else None, len(tiles), nx, ny,
                            np.ascontiguousarray(data, np.float64), ptr(grad_x), ptr(grad_y), ptr(gmask, np.int32), int(has_missing),
                            float(missing), int(monotonic), int(cell_methods), ptr(weight), ptr(cell_area), ptr(farea), int(target),
                            ptr(dst_cell_area), nx_out, ny_out, out)
    return out


def ref_apply_ex(handle, order, data, nout, grad_x=None, grad_y=None, gmask=None, has_missing=False, missing=0.0, monotonic=False,
                 cell_methods=0, weight=None, farea=None, area_missing=-1e20, target=False):
    L = ref_lib()
    out = np.zeros(nout)
    keep = []

    def ptr(a, dt=np.float64):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt); keep.append(a)
        return a.ctypes.data
    extra = (MONOTONIC if monotonic else 0) | (TARGET if target else 0)
    L.ref_regrid_apply_ex(handle, order, int(has_missing), float(missing), int(cell_methods), ptr(weight), ptr(farea), float(area_missing),
                          extra, np.ascontiguousarray(data, np.float64), ptr(grad_x), ptr(grad_y), ptr(gmask, np.int32), out)
    return out


NC_TYPES = {1: "b", 2: "c", 3: "h", 4: "i", 5: "f", 6: "d"}


def ref_store_file(name):
    """what the reference handed its I/O layer for file `name` (oracle/shim/io_stubs.c): dims [(name, size)], vars
    [(name, nc_type, dim indices, [(att, value)], data)] in definition order"""
    L = ref_lib()
    f = L.stub_find(name.encode())
    assert f >= 0, name
    dims = [(L.stub_dim_name(f, d).decode(), L.stub_dim_size(f, d)) for d in range(L.stub_ndims(f))]
    out = []
    for v in range(L.stub_nvars(f)):
        t = L.stub_var_type(f, v)
        dd = [L.stub_var_dim(f, v, k) for k in range(L.stub_var_ndim(f, v))]
        atts = [(L.stub_var_att_name(f, v, a).decode(), L.stub_var_att_value(f, v, a).decode()) for a in range(L.stub_var_natts(f, v))]
        n = L.stub_var_nelem(f, v)
        dt = {4: np.int32, 6: np.float64}[t]
        data = np.ctypeslib.as_array(C.cast(L.stub_var_data(f, v), C.POINTER(C.c_int if t == 4 else C.c_double)), shape=(n,)).astype(dt).copy()
        out.append((L.stub_var_name(f, v).decode(), t, dd, atts, data.reshape([dims[k][1] for k in dd])))
    return dims, out


def ref_store_load(name, path):
    """put the contents of a classic netCDF file (read with scipy, an implementation independent of csrc/nc3.c) into the
    reference's in-memory file store under `name`, so that its READ branch can be run on it"""
    from scipy.io import netcdf_file
    L = ref_lib()
    g = netcdf_file(path, "r", mmap=False)
    f = L.stub_new_file(name.encode())
    dn = list(g.dimensions)
    for d in dn:
        L.stub_add_dim(f, d.encode(), g.dimensions[d])
    for vn, var in g.variables.items():
        t = {"i": 4, "d": 6}[var.typecode()]
        dd = np.array([dn.index(d) for d in var.dimensions], np.int32)
        v = L.stub_add_var(f, vn.encode(), t, len(dd), dd)
        data = np.ascontiguousarray(var[:], np.int32 if t == 4 else np.float64)
        L.stub_set_var_data(f, v, data.ctypes.data)
    g.close()
