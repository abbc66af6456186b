"""GPU parity tests of the conservative apply path (through the C ABI): grad_c2l, do_scalar_conserve_interp
(order 1 / order 2 / missing values / monotone limiter) batched over field-levels, against the CPU oracle and the
golden vectors the unmodified reference produced.  Bar: remapped fields BIT-IDENTICAL (the per-destination sums run
in the reference's order); device-computed gradient metrics (asin/acos/atan2 on the device) to 1e-12 relative."""
import os

import numpy as np
import pytest

import xgtest

pytestmark = pytest.mark.gpu

FIELD_RTOL = 1e-12


class Case:
    def __init__(self, pkg, ni, nlon, nlat, order):
        self.ni, self.nlon, self.nlat, self.order = ni, nlon, nlat, order
        self.lonc, self.latc, self.lont, self.latt = pkg.cubed_sphere_grid(ni, centers=True)
        self.lon2, self.lat2 = pkg.latlon_grid(nlon, nlat)
        self.hm = xgtest.cubed_sphere_halo_map(self.lonc, self.latc)
        self.xt = xgtest.with_halo(self.lont.reshape(-1), self.hm); self.yt = xgtest.with_halo(self.latt.reshape(-1), self.hm)
        self.nh, self.nc = (ni + 2) ** 2, ni * ni
        self.tiles = [(ni, ni)] * 6
        self.plan = pkg.XgridPlan(0)
        self.plan.set_dst(self.lon2, self.lat2)
        self.plan.set_src(self.lonc, self.latc)
        self.plan.generate(order)
        self.x = self.plan.result_host()
        self.plan.apply_setup()
        self.metrics = None

    def oracle_metrics(self):
        if self.metrics is None:
            ni, nh = self.ni, self.nh
            self.metrics = [xgtest.c2l_metrics("oracle", ni, ni, self.xt[t * nh:(t + 1) * nh], self.yt[t * nh:(t + 1) * nh],
                                               self.lonc[t], self.latc[t]) for t in range(6)]
        return self.metrics

    def oracle_grad(self, fh, missing=None):
        ni, nh, nc = self.ni, self.nh, self.nc
        gx = np.zeros(6 * nc); gy = np.zeros(6 * nc); gm = np.zeros(6 * nc, np.int32)
        for t, m in enumerate(self.oracle_metrics()):
            gx[t * nc:(t + 1) * nc], gy[t * nc:(t + 1) * nc] = xgtest.grad_c2l("oracle", ni, ni, fh[t * nh:(t + 1) * nh], m)
            if missing is not None:
                gm[t * nc:(t + 1) * nc] = xgtest.grad_mask(ni, ni, fh[t * nh:(t + 1) * nh], missing)
        return gx, gy, gm

    def fields(self, nf, seed=1234, holes=0.0, missing=-999.0):
        rng = np.random.default_rng(seed)
        f = np.stack([xgtest.smooth_field(self.lont, self.latt, k, k // 3) if k % 2 == 0 else rng.uniform(0, 1, 6 * self.nc)
                      for k in range(nf)])
        if holes:
            f[np.random.default_rng(4321).uniform(size=f.shape) < holes] = missing
        return f


def test_grad_c2l_and_metrics(pkg):
    c = Case(pkg, 16, 90, 45, 2)
    p = c.plan
    for t, m in enumerate(c.oracle_metrics()):
        p.grad_set_metrics(t, m)
    nf = 11                                            # not a multiple of the per-thread field tile
    f = c.fields(nf, holes=0.05)
    fh = xgtest.with_halo(f, c.hm)
    gx, gy, gm = p.grad_c2l(fh.reshape(-1), nf, has_missing=True, missing=-999.0)
    for k in range(nf):
        ox, oy, om = c.oracle_grad(fh[k], -999.0)
        s = slice(k * 6 * c.nc, (k + 1) * 6 * c.nc)
        assert np.array_equal(gx[s], ox) and np.array_equal(gy[s], oy) and np.array_equal(gm[s], om), k
    # metrics computed on the device agree with the reference's to rounding
    p.grad_setup(c.xt, c.yt)
    for t, m in enumerate(c.oracle_metrics()):
        got = p.grad_get_metrics(t)
        for k in xgtest.METRICS:
            scale = np.max(np.abs(m[k]))
            if k in ("vlon", "vlat", "en_n", "en_e"):
                assert np.array_equal(got[k], m[k]), (t, k)     # reference's own sin / cos / sincos bits
            elif k == "area":
                # spherical excess of four acosl() angles: bit-identical but where the x87 fpatan is not correctly rounded
                assert np.mean(got[k] == m[k]) > 0.99 and np.max(np.abs(got[k] - m[k])) <= 1e-13 * scale, (t, np.mean(got[k] == m[k]))
            else:
                # dx, dy, edge weights: asin / atan2 in double-double, correctly rounded; glibc's own differ from that on about one
                # argument per thousand
                assert np.max(np.abs(got[k] - m[k])) <= 1e-15 * scale, (t, k, np.max(np.abs(got[k] - m[k])) / scale)
                assert np.mean(got[k] == m[k]) > 0.9, (t, k, np.mean(got[k] == m[k]))


@pytest.mark.parametrize("ni,nlon,nlat", [(8, 36, 18), (24, 144, 72)])
def test_apply_order1(pkg, ni, nlon, nlat):
    c = Case(pkg, ni, nlon, nlat, 1)
    nf = 10
    for holes in (0.0, 0.05):
        f = c.fields(nf, holes=holes)
        out = c.plan.apply(1, f.reshape(-1), nf, has_missing=holes > 0, missing=-999.0).reshape(nf, -1)
        for k in range(nf):
            want = xgtest.oracle_apply(c.x, 1, c.tiles, f[k], nlon, nlat, has_missing=holes > 0, missing=-999.0)
            assert np.array_equal(out[k], want), (holes, k)


@pytest.mark.parametrize("ni,nlon,nlat", [(8, 36, 18), (24, 144, 72), (24, 18, 9), (16, 250, 130)])
def test_apply_order2_variants(pkg, ni, nlon, nlat):
    """(24, 18, 9): destination far coarser than the source, every 32x8 destination patch references more source cells than
    the tiled kernel stages (direct-gather patches); (16, 250, 130): patch grid with ragged right/top edges"""
    c = Case(pkg, ni, nlon, nlat, 2)
    p = c.plan
    for t, m in enumerate(c.oracle_metrics()):
        p.grad_set_metrics(t, m)
    nf = 9
    for holes in (0.0, 0.05):
        hmiss = holes > 0
        f = c.fields(nf, holes=holes)
        fh = xgtest.with_halo(f, c.hm)
        gx, gy, gm = p.grad_c2l(fh.reshape(-1), nf, has_missing=hmiss, missing=-999.0)
        out = p.apply(2, fh.reshape(-1), nf, gx, gy, gm, has_missing=hmiss, missing=-999.0).reshape(nf, -1)
        fused = p.regrid(2, fh.reshape(-1), nf, has_missing=hmiss, missing=-999.0).reshape(nf, -1)
        mono = p.regrid(2 | xgtest.MONOTONIC, fh.reshape(-1), nf, has_missing=hmiss, missing=-999.0).reshape(nf, -1)
        for k in range(nf):
            ox, oy, om = c.oracle_grad(fh[k], -999.0 if hmiss else None)
            want = xgtest.oracle_apply(c.x, 2, c.tiles, fh[k], nlon, nlat, ox, oy, om, has_missing=hmiss, missing=-999.0)
            assert np.array_equal(out[k], want), (holes, k)
            assert np.array_equal(fused[k], want), (holes, k)
            want = xgtest.oracle_apply(c.x, 2, c.tiles, fh[k], nlon, nlat, ox, oy, om, has_missing=hmiss, missing=-999.0, monotonic=True)
            assert np.array_equal(mono[k], want), (holes, k)


def test_apply_matches_reference_golden(pkg):
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "apply_c8_36x18.npz"))
    ni, nlon, nlat = int(g["ni"]), int(g["nlon"]), int(g["nlat"])
    miss = float(g["missing"])
    sz = xgtest.metric_sizes(ni, ni)
    c1 = Case(pkg, ni, nlon, nlat, 1)
    out = c1.plan.apply(1, g["fields"].reshape(-1), 2).reshape(2, -1)
    assert np.array_equal(out, g["out_o1"])
    out = c1.plan.apply(1, g["fields_missing"].reshape(-1), 2, has_missing=True, missing=miss).reshape(2, -1)
    assert np.array_equal(out, g["out_o1_missing"])
    c2 = Case(pkg, ni, nlon, nlat, 2)
    for t in range(6):
        c2.plan.grad_set_metrics(t, {k: g["m_" + k][t * sz[k]:(t + 1) * sz[k]] for k in xgtest.METRICS})
    for name, src, hmf in (("", g["fields"], False), ("_missing", g["fields_missing"], True)):
        fh = xgtest.with_halo(src, c2.hm)
        gx, gy, gm = c2.plan.grad_c2l(fh.reshape(-1), 2, has_missing=hmf, missing=miss)
        assert np.array_equal(gx.reshape(2, -1), g["grad_x" + name]) and np.array_equal(gy.reshape(2, -1), g["grad_y" + name])
        assert np.array_equal(gm.reshape(2, -1), g["grad_mask" + name])
        out = c2.plan.regrid(2, fh.reshape(-1), 2, has_missing=hmf, missing=miss).reshape(2, -1)
        assert np.array_equal(out, g["out_o2" + name]), name
        out = c2.plan.regrid(2 | xgtest.MONOTONIC, fh.reshape(-1), 2, has_missing=hmf, missing=miss).reshape(2, -1)
        assert np.array_equal(out, g["out_o2_mono" + name]), name


def test_device_metrics_end_to_end_and_conservation(pkg):
    """config 2 shape at reduced size: C48 -> 360x180 order 2, 33 levels; metrics computed on the device.
    Remapped fields within 1e-12 relative of the oracle, and the global integral is conserved."""
    c = Case(pkg, 48, 360, 180, 2)
    p = c.plan
    p.grad_setup(c.xt, c.yt)
    nf = 33
    f = np.stack([xgtest.smooth_field(c.lont, c.latt, k, 1) for k in range(nf)])
    fh = xgtest.with_halo(f, c.hm)
    out = p.regrid(2, fh.reshape(-1), nf).reshape(nf, -1)
    src_area = p.src_area(); dst_area = p.dst_area()
    for k in (0, 7, 32):
        ox, oy, om = c.oracle_grad(fh[k])
        want = xgtest.oracle_apply(c.x, 2, c.tiles, fh[k], c.nlon, c.nlat, ox, oy, om)
        assert np.max(np.abs(out[k] - want) / np.abs(want)) <= FIELD_RTOL
    # conservation: sum(out * xgrid area per destination) == sum over exchange cells of the reconstructed field
    xa = np.zeros(c.nlon * c.nlat)
    np.add.at(xa, c.x["j_out"].astype(np.int64) * c.nlon + c.x["i_out"], c.x["area"])
    for k in (0, 32):
        tot_out = float(np.sum(out[k] * xa))
        tot_in = float(np.sum(f[k] * src_area))
        assert abs(tot_out - tot_in) / abs(tot_in) < 1e-6      # second-order reconstruction: equal up to the centroid terms
    # order-1 remap of the same exchange grid conserves to rounding
    o1 = p.apply(1, f.reshape(-1), nf).reshape(nf, -1)
    xs = np.zeros(6 * c.nc)
    np.add.at(xs, c.x["t_in"].astype(np.int64) * c.nc + c.x["j_in"].astype(np.int64) * c.ni + c.x["i_in"], c.x["area"])
    for k in (0, 32):
        assert abs(np.sum(o1[k] * xa) - np.sum(f[k] * xs)) / abs(np.sum(f[k] * xs)) < 1e-13


def test_set_xgrid_from_host_lists_and_device_buffers(pkg):
    """READ branch of setup_conserve_interp: lists handed in (here in a scrambled order, as a fregrid_parallel remap file
    has them); the sums follow the list order given, like the reference's loop would"""
    import torch
    c = Case(pkg, 12, 48, 24, 2)
    rng = np.random.default_rng(5)
    perm = rng.permutation(c.x["area"].size)
    x = {k: np.ascontiguousarray(c.x[k][perm]) for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "di", "dj")}
    p = pkg.XgridPlan(0)
    p.set_xgrid(c.tiles, c.nlon, c.nlat, x)
    for t, m in enumerate(c.oracle_metrics()):
        p.grad_set_metrics(t, m)
    nf = 3
    f = c.fields(nf)
    fh = xgtest.with_halo(f, c.hm)
    d_in = torch.from_numpy(fh.reshape(-1)).cuda()
    out = p.regrid(2, d_in, nf)                       # device-resident in, device-resident out
    torch.cuda.synchronize(); p.sync()
    out = out.cpu().numpy().reshape(nf, -1)
    for k in range(nf):
        ox, oy, om = c.oracle_grad(fh[k])
        want = xgtest.oracle_apply(x, 2, c.tiles, fh[k], c.nlon, c.nlat, ox, oy, om)
        assert np.array_equal(out[k], want), k
    o1 = p.apply(1, f.reshape(-1), nf).reshape(nf, -1)
    for k in range(nf):
        assert np.array_equal(o1[k], xgtest.oracle_apply(x, 1, c.tiles, f[k], c.nlon, c.nlat)), k


def test_reference_signature_setup_and_apply_through_real_structs(pkg):
    """the drop-in claim end to end: the reference's own driver code (oracle/ref_driver.c, compiled against the real headers)
    builds Grid_config / Field_config structs and calls libxgrid_b200's setup_conserve_interp and do_scalar_conserve_interp
    through function pointers; lists and remapped fields equal the reference's own, bit for bit"""
    import ctypes as C
    R = xgtest.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built")
    L = pkg.lib()
    setup_fn = C.cast(L.setup_conserve_interp, C.c_void_p)
    apply_fn = C.cast(L.do_scalar_conserve_interp, C.c_void_p)
    ni, nlon, nlat = 12, 48, 24
    lonc, latc, lont, latt = pkg.cubed_sphere_grid(ni, centers=True)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    nh, nc = (ni + 2) ** 2, ni * ni
    rng = np.random.default_rng(17)
    for order in (1, 2):
        ref = xgtest.ref_setup(lonc, latc, lon2, lat2, order, keep=True)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)      # the NOTE line
        try:
            h = R.ref_regrid_setup_through(ref["handle"], setup_fn)
        finally:
            os.dup2(saved, 1); os.close(saved); os.close(devnull)
        n = R.ref_regrid_nxgrid(h)
        assert n == ref["nxgrid"]
        out = xgtest._alloc(n, order)
        R.ref_regrid_get(h, out["t_in"], out["i_in"], out["j_in"], out["i_out"], out["j_out"], out["area"],
                         out["di"].ctypes.data if order == 2 else None, out["dj"].ctypes.data if order == 2 else None)
        for k in out:
            assert np.array_equal(out[k], ref[k]), (order, k)
        # fields: two levels in one call (nz = 2), tile-major with the levels inside each tile like Field_config.data
        f = rng.uniform(0, 1, (2, 6, nc))
        if order == 1:
            data = np.ascontiguousarray(f.transpose(1, 0, 2)).reshape(-1)
            gx = gy = gm = None
        else:
            fh = np.stack([xgtest.with_halo(f[k].reshape(-1), hm).reshape(6, nh) for k in range(2)])
            data = np.ascontiguousarray(fh.transpose(1, 0, 2)).reshape(-1)
            gx = rng.normal(size=(6, 2, nc)).reshape(-1); gy = rng.normal(size=(6, 2, nc)).reshape(-1)
            gm = np.zeros(6 * nc, np.int32)
        want = np.zeros(2 * nlon * nlat); got = np.zeros(2 * nlon * nlat)
        ptr = lambda a: None if a is None else a.ctypes.data
        R.ref_regrid_apply(ref["handle"], order, 0, 0.0, 0, 2, 0, data, ptr(gx), ptr(gy), ptr(gm), want)
        R.ref_regrid_apply_through(h, apply_fn, order, 0, 0.0, 2, 0, data, ptr(gx), ptr(gy), ptr(gm), got)
        assert np.array_equal(got, want), order
        # one level with missing values and the monotone limiter
        if order == 2:
            f1 = f[0].reshape(-1).copy(); f1[rng.uniform(size=f1.size) < 0.05] = -999.0
            d1 = xgtest.with_halo(f1, hm)
            p = pkg.XgridPlan(0)       # gradients from the library itself, reference metrics not needed for this comparison
            g1x = rng.normal(size=6 * nc); g1y = rng.normal(size=6 * nc)
            g1m = np.concatenate([xgtest.grad_mask(ni, ni, d1[t * nh:(t + 1) * nh], -999.0) for t in range(6)])
            for extra in (0, xgtest.MONOTONIC):
                want = np.zeros(nlon * nlat); got = np.zeros(nlon * nlat)
                R.ref_regrid_apply(ref["handle"], 2, 1, -999.0, 0, 1, extra, d1, g1x.ctypes.data, g1y.ctypes.data, g1m.ctypes.data, want)
                R.ref_regrid_apply_through(h, apply_fn, 2, 1, -999.0, 1, extra, d1, g1x.ctypes.data, g1y.ctypes.data, g1m.ctypes.data, got)
                assert np.array_equal(got, want), extra
            p.close()
        R.ref_regrid_free(ref["handle"])


def test_apply_variants_sum_measures_weight_target(pkg):
    """cell_methods sum, cell_measures, weight field, --target_grid and their combinations with missing values and the
    monotone limiter (conserve_interp.c:572-585, :821-865): GPU == oracle bit for bit; and once more through the
    reference-signature do_scalar_conserve_interp with the real structs"""
    import itertools
    ni, nlon, nlat = 12, 40, 24
    c1 = Case(pkg, ni, nlon, nlat, 1)
    c2 = Case(pkg, ni, nlon, nlat, 2)
    nc, nh = c1.nc, c1.nh
    rng = np.random.default_rng(5)
    f = rng.uniform(0.5, 1.5, 6 * nc); fm = f.copy(); fm[rng.uniform(size=f.size) < 0.08] = -999.0
    w = rng.uniform(0.2, 1.0, 6 * nc); fa = rng.uniform(1e9, 2e9, 6 * nc)
    ncase = 0
    for c in (c1, c2):
        order = c.order
        ca = c.plan.src_area(); da = c.plan.dst_area()
        for hmiss, cm, usew, usefa, tgt, mono in itertools.product((0, 1), repeat=6):
            if (cm and (usefa or tgt)) or (mono and order == 1):
                continue
            src = fm if hmiss else f
            data = src if order == 1 else xgtest.with_halo(src, c.hm)
            gx = rng.normal(size=6 * nc) * 0.1; gy = rng.normal(size=6 * nc) * 0.1
            gm = np.zeros(6 * nc, np.int32)
            if order == 2 and hmiss:
                gm = np.concatenate([xgtest.grad_mask(ni, ni, data[t * nh:(t + 1) * nh], -999.0) for t in range(6)])
            want = xgtest.oracle_apply_ex(c.x, order, c.tiles, data, nlon, nlat, gx, gy, gm, bool(hmiss), -999.0, bool(mono), cm,
                                          w if usew else None, ca, fa if usefa else None, bool(tgt), da)
            c.plan.apply_options(cm, w if usew else None, None, fa if usefa else None, -1e20, bool(tgt), None)
            op = order | (xgtest.MONOTONIC if mono else 0)
            got = c.plan.apply(op, data, 1, gx if order == 2 else None, gy if order == 2 else None, gm if order == 2 else None,
                               has_missing=bool(hmiss), missing=-999.0)
            assert np.array_equal(got, want), (order, hmiss, cm, usew, usefa, tgt, mono)
            ncase += 1
        c.plan.apply_options()
        # back to the plain mean
        data = f if order == 1 else xgtest.with_halo(f, c.hm)
        plain = c.plan.apply(order, data, 1, np.zeros(6 * nc) if order == 2 else None, np.zeros(6 * nc) if order == 2 else None)
        assert np.array_equal(plain, xgtest.oracle_apply(c.x, order, c.tiles, data, nlon, nlat, np.zeros(6 * nc), np.zeros(6 * nc)))
    assert ncase == 60
    # data present where the field area is missing is the reference's fatal error (:578): reported, not ignored
    bad = fa.copy(); bad[7] = -1e20
    c1.plan.apply_options(0, None, None, bad, -1e20, False, None)
    with pytest.raises(pkg.XgridError, match="area is missing"):
        c1.plan.apply(1, fm if fm[7] != -999.0 else f, 1, has_missing=True, missing=-999.0)
    c1.plan.apply_options()
    # reference signature with the real structs: sum + weight, and cell_measures + target
    R = xgtest.ref_lib()
    if R is not None:
        import ctypes as C
        L = pkg.lib()
        apply_fn = C.cast(L.do_scalar_conserve_interp, C.c_void_p)
        R.ref_regrid_apply_ex_through.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                                  C.c_uint, xgtest.dp, C.c_void_p, C.c_void_p, C.c_void_p, xgtest.dp]
        ref = xgtest.ref_setup(c1.lonc, c1.latc, c1.lon2, c1.lat2, 1, keep=True)
        for cm, usefa, tgt in ((1, 0, 0), (0, 1, 1)):
            want = xgtest.ref_apply_ex(ref["handle"], 1, fm, nlon * nlat, has_missing=True, missing=-999.0, cell_methods=cm, weight=w,
                                       farea=fa if usefa else None, target=bool(tgt))
            got = np.zeros(nlon * nlat)
            R.ref_regrid_apply_ex_through(ref["handle"], apply_fn, 1, 1, -999.0, cm, w.ctypes.data, fa.ctypes.data if usefa else None, -1e20,
                                          xgtest.TARGET if tgt else 0, np.ascontiguousarray(fm), None, None, None, got)
            assert np.array_equal(got, want), (cm, usefa, tgt)
        R.ref_regrid_free(ref["handle"])


def test_order2_over_several_output_tiles_and_mixed_order_variables(pkg):
    """(a) conserve_order2 onto a 6-tile output mosaic — the reference sums every output tile's exchange cells
    per source cell before the AREA_RATIO test and the centroid subtraction (conserve_interp.c:204-221, :319-358), so
    tile1_distance of a source cell that straddles two output tiles depends on all of them; (b) an order-1 variable
    followed by an order-2 variable on the same Interp_config (interp_method is per variable, conserve_interp.c:528).
    Both through the reference's own structs and libxgrid_b200's exported setup_conserve_interp /
    do_scalar_conserve_interp, against the compiled reference, bit for bit."""
    import ctypes as C
    R = xgtest.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built")
    L = pkg.lib()
    setup_fn = C.cast(L.setup_conserve_interp, C.c_void_p)
    apply_fn = C.cast(L.do_scalar_conserve_interp, C.c_void_p)
    ni, no = 12, 10
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lono, lato = pkg.cubed_sphere_grid(no)
    href, ref = xgtest.ref_multi_setup(lonc, latc, lono, lato, 2)
    hgot, got = xgtest.ref_multi_setup(lonc, latc, lono, lato, 2, setup_fn)
    straddlers = 0
    for n in range(6):
        assert got[n]["nxgrid"] == ref[n]["nxgrid"] > 0
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "di", "dj"):
            assert np.array_equal(got[n][k], ref[n][k]), (n, k)
    cells = [set(zip(r["t_in"].tolist(), r["j_in"].tolist(), r["i_in"].tolist())) for r in ref]
    for a in range(6):
        for b in range(a + 1, 6):
            straddlers += len(cells[a] & cells[b])
    assert straddlers > 100                      # the case the per-tile correction got wrong is really exercised
    # (b) mixed interp_method on one Interp_config
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    nc = ni * ni
    rng = np.random.default_rng(5)
    f = rng.uniform(0, 1, 6 * nc)
    fh = xgtest.with_halo(f, hm)
    gx = rng.normal(size=6 * nc); gy = rng.normal(size=6 * nc); gm = np.zeros(6 * nc, np.int32)
    nout = 6 * no * no
    for order, data in ((1, f), (2, fh), (1, f), (2, fh)):
        want = np.zeros(nout); out = np.zeros(nout)
        a = (gx.ctypes.data, gy.ctypes.data, gm.ctypes.data) if order == 2 else (None, None, None)
        R.ref_multi_apply(href, None, order, data, *a, want)
        R.ref_multi_apply(hgot, apply_fn, order, data, *a, out)
        assert np.array_equal(out, want), order


def test_exchange_grid_naming_cells_outside_the_output_tile_is_refused(pkg):
    """a stale remap file written for another output grid must raise, not write out of bounds"""
    ni, nlon, nlat = 8, 36, 18
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    x = xgtest.oracle_setup(lonc, latc, lon2, lat2, 1)
    p = pkg.XgridPlan(0)
    bad = dict(x); bad["j_out"] = x["j_out"].copy(); bad["j_out"][7] = nlat + 3
    with pytest.raises(pkg.XgridError, match="outside"):
        p.set_xgrid([(ni, ni)] * 6, nlon, nlat, bad)
    bad = dict(x); bad["i_out"] = x["i_out"].copy(); bad["i_out"][11] = -1
    with pytest.raises(pkg.XgridError, match="outside"):
        p.set_xgrid([(ni, ni)] * 6, nlon, nlat, bad)
    p.set_xgrid([(ni, ni)] * 6, nlon, nlat, x)          # the plan is still usable
    p.close()


def test_config1_full_size_396_field_levels_three_compared_with_the_oracle(pkg):
    """BASELINE configs[1] at full size — C96 -> 1440x720 order 2 with gradient terms, 33 levels x 12 times in ONE batched
    call (what bench.py's apply leg times): three of the 396 remapped field-levels (first, middle, last; a smooth and two
    random ones) against the CPU oracle's grad_c2l + do_scalar_conserve_interp, bit for bit (oracle metrics handed in)."""
    c = Case(pkg, 96, 1440, 720, 2)
    p = c.plan
    for t, m in enumerate(c.oracle_metrics()):
        p.grad_set_metrics(t, m)
    B = 33 * 12
    rng = np.random.default_rng(1234)
    base = xgtest.smooth_field(c.lont, c.latt)
    f = np.stack([(base + 0.01 * (b % 33) + 0.001 * (b // 33)).reshape(-1) if b % 2 == 0 else rng.uniform(0, 1, 6 * c.nc) for b in range(B)])
    fh = xgtest.with_halo(f, c.hm)
    out = p.regrid(2, fh.reshape(-1), B).reshape(B, -1)
    for k in (0, 197, B - 1):
        ox, oy, om = c.oracle_grad(fh[k])
        want = xgtest.oracle_apply(c.x, 2, c.tiles, fh[k], c.nlon, c.nlat, ox, oy, om)
        assert np.array_equal(out[k], want), k
    assert np.all(np.isfinite(out))
    p.close()


def test_shared_reciprocal_division_is_the_compilers_division(pkg):
    """csrc/shared_div.cuh: out = sum / out_area for every field-level of a destination cell shares one reciprocal.  Each
    quotient must be bit-identical to `a / b`: 2^24 pairs over the operand ranges the apply path sees (areas 1e3..1e13,
    sums of either sign), wide random exponents, and the edges (tiny, huge, zero, equal, powers of two)."""
    import ctypes as C
    L = pkg.lib()
    L.xgb_shared_div_check.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong)]
    rng = np.random.default_rng(2024)
    n = 1 << 22
    sets = []
    sets.append((rng.normal(0, 1, n) * 10.0 ** rng.uniform(0, 14, n), 10.0 ** rng.uniform(3, 13, n)))
    sets.append((rng.normal(0, 1, n) * 2.0 ** rng.integers(-1000, 1000, n), rng.uniform(1, 2, n) * 2.0 ** rng.integers(-1000, 1000, n)))
    m = rng.uniform(1, 2, n)
    sets.append((m * rng.integers(1, 1000, n), m))                                            # exact and nearly exact quotients
    edge = np.array([0.0, -0.0, 1.0, -1.0, 2.0 ** -1074, 2.0 ** -1022, 2.0 ** 1023, 1.7976931348623157e308, 3.0, 1.0 / 3.0, np.inf, np.nan, 5e-324])
    ea, eb = np.meshgrid(edge, edge)
    sets.append((ea.reshape(-1).copy(), eb.reshape(-1).copy()))
    with np.errstate(all="ignore"):
        for a, b in sets:
            a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
            bad = C.c_longlong(-1)
            assert L.xgb_shared_div_check(a.size, a.ctypes.data, b.ctypes.data, C.byref(bad)) == 0
            assert bad.value == 0, bad.value


def test_fused_regrid_on_tiles_of_any_shape_equals_the_two_step_path(pkg):
    """grad_c2l_rec_kernel deals its threads the interior cells of every tile first and the rim cells after (the rim needs
    a2b_ord2's edge formulas, gradient_c2l.c:169-196): a mosaic of a square, a wide, a two-row and a one-column tile — the thin
    ones are all rim — must give exactly what the storage-order gradient kernel followed by the apply kernel gives."""
    shapes = [(6, 6), (9, 4), (7, 2), (1, 5), (3, 3)]                  # (nx, ny)
    lon0 = 20.0
    lons, lats, xt, yt = [], [], [], []
    for nx, ny in shapes:                                               # regional lat-lon tiles side by side
        xe = np.deg2rad(lon0 + np.arange(nx + 1) * 1.5); ye = np.deg2rad(-10.0 + np.arange(ny + 1) * 1.25)
        lo, la = np.meshgrid(xe, ye)
        lons.append(lo); lats.append(la)
        xc = np.deg2rad(lon0 + (np.arange(-1, nx + 1) + 0.5) * 1.5); yc = np.deg2rad(-10.0 + (np.arange(-1, ny + 1) + 0.5) * 1.25)
        cx, cy = np.meshgrid(xc, yc)
        xt.append(cx.reshape(-1)); yt.append(cy.reshape(-1))
        lon0 += nx * 1.5 + 3.0
    lon2, lat2 = pkg.latlon_grid(90, 45)
    p = pkg.XgridPlan(0)
    p.set_dst(lon2, lat2)
    p.set_src(lons, lats)
    assert p.generate(2) > 0
    p.apply_setup()
    p.grad_setup(np.concatenate(xt), np.concatenate(yt))
    rng = np.random.default_rng(11)
    nf = 70                                                             # two 64-field-level chunks, ragged
    nh = sum((nx + 2) * (ny + 2) for nx, ny in shapes)
    for hmiss in (False, True):
        fh = rng.uniform(0.0, 1.0, (nf, nh))
        if hmiss:
            fh[rng.uniform(size=fh.shape) < 0.05] = -999.0
        gx, gy, gm = p.grad_c2l(fh.reshape(-1), nf, has_missing=hmiss, missing=-999.0)
        two_step = p.apply(2, fh.reshape(-1), nf, gx, gy, gm, has_missing=hmiss, missing=-999.0)
        fused = p.regrid(2, fh.reshape(-1), nf, has_missing=hmiss, missing=-999.0)
        assert np.array_equal(two_step, fused, equal_nan=True), hmiss
        assert (fused != 0).any()
    p.close()
