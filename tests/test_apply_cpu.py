"""CPU-side checks of the apply path's oracle (no GPU): orc_conserve_apply / orc_grad_c2l / orc_calc_c2l_grid_info
against golden vectors produced by the unmodified reference (tests/golden/apply_*.npz, made by make_golden.py)
and, where oracle/_ref exists, live against the compiled reference.  Everything here is bit-exact."""
import os

import numpy as np
import pytest

import xgtest


def _case():
    g = np.load(os.path.join(xgtest.GOLDEN_DIR, "apply_c8_36x18.npz"))
    x1 = np.load(os.path.join(xgtest.GOLDEN_DIR, "xgrid_c8_36x18_o1.npz"))
    x2 = np.load(os.path.join(xgtest.GOLDEN_DIR, "xgrid_c8_36x18_o2.npz"))
    return g, x1, x2


def _metrics(g, ni, t):
    sz = xgtest.metric_sizes(ni, ni)
    return {k: np.ascontiguousarray(g["m_" + k][t * sz[k]:(t + 1) * sz[k]]) for k in xgtest.METRICS}


def test_c2l_metrics_and_gradient_match_golden():
    g, x1, x2 = _case()
    ni = int(g["ni"]); nh = (ni + 2) ** 2; nc = ni * ni; nv = (ni + 1) ** 2
    lonc = x2["lon_in"].reshape(6, ni + 1, ni + 1); latc = x2["lat_in"].reshape(6, ni + 1, ni + 1)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    for t in range(6):
        m = xgtest.c2l_metrics("oracle", ni, ni, g["xt"][t * nh:(t + 1) * nh], g["yt"][t * nh:(t + 1) * nh], lonc[t], latc[t])
        want = _metrics(g, ni, t)
        for k in xgtest.METRICS:
            assert np.array_equal(m[k], want[k]), (t, k)
        for name, fields in (("", g["fields"]), ("_missing", g["fields_missing"])):
            for f in range(fields.shape[0]):
                fh = xgtest.with_halo(fields[f], hm)[t * nh:(t + 1) * nh]
                gx, gy = xgtest.grad_c2l("oracle", ni, ni, fh, want)
                assert np.array_equal(gx, g["grad_x" + name][f][t * nc:(t + 1) * nc])
                assert np.array_equal(gy, g["grad_y" + name][f][t * nc:(t + 1) * nc])
                if name:
                    assert np.array_equal(xgtest.grad_mask(ni, ni, fh, float(g["missing"])), g["grad_mask_missing"][f][t * nc:(t + 1) * nc])


def test_conserve_apply_matches_golden():
    g, x1, x2 = _case()
    ni, nlon, nlat = int(g["ni"]), int(g["nlon"]), int(g["nlat"])
    tiles = [(ni, ni)] * 6
    miss = float(g["missing"])
    lonc = x2["lon_in"].reshape(6, ni + 1, ni + 1); latc = x2["lat_in"].reshape(6, ni + 1, ni + 1)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    X1 = {k: np.ascontiguousarray(x1[k]) for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area")}
    X2 = {k: np.ascontiguousarray(x2[k]) for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "di", "dj")}
    for f in range(2):
        assert np.array_equal(xgtest.oracle_apply(X1, 1, tiles, g["fields"][f], nlon, nlat), g["out_o1"][f])
        assert np.array_equal(xgtest.oracle_apply(X1, 1, tiles, g["fields_missing"][f], nlon, nlat, has_missing=True, missing=miss),
                              g["out_o1_missing"][f])
        for name, src, hmf in (("", g["fields"], False), ("_missing", g["fields_missing"], True)):
            fh = xgtest.with_halo(src[f], hm)
            gx, gy, gm = g["grad_x" + name][f], g["grad_y" + name][f], g["grad_mask" + name][f]
            got = xgtest.oracle_apply(X2, 2, tiles, fh, nlon, nlat, gx, gy, gm, has_missing=hmf, missing=miss)
            assert np.array_equal(got, g["out_o2" + name][f]), name
            got = xgtest.oracle_apply(X2, 2, tiles, fh, nlon, nlat, gx, gy, gm, has_missing=hmf, missing=miss, monotonic=True)
            assert np.array_equal(got, g["out_o2_mono" + name][f]), name
    # the monotone limiter really acted on the random field, and every destination cell is covered
    assert not np.array_equal(g["out_o2"][1], g["out_o2_mono"][1])
    assert np.all(g["out_o2"] > -1e19)


def test_oracle_apply_equals_compiled_reference(reflib, pkg):
    """live: another grid pair, including destination cells no source cell covers (regional source)"""
    ni, nlon, nlat = 12, 40, 24
    lonc, latc, lont, latt = xgtest.ref_cubed_sphere(ni, centers=True)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    xt = xgtest.with_halo(lont.reshape(-1), hm); yt = xgtest.with_halo(latt.reshape(-1), hm)
    nh = (ni + 2) ** 2; nc = ni * ni
    tiles = [(ni, ni)] * 6
    rng = np.random.default_rng(7)
    f = rng.uniform(0, 1, 6 * nc)
    f[rng.uniform(size=f.size) < 0.1] = 1e20
    fh = xgtest.with_halo(f, hm)
    gx = np.zeros(6 * nc); gy = np.zeros(6 * nc); gm = np.zeros(6 * nc, np.int32)
    for t in range(6):
        m = xgtest.c2l_metrics("oracle", ni, ni, xt[t * nh:(t + 1) * nh], yt[t * nh:(t + 1) * nh], lonc[t], latc[t])
        r = xgtest.c2l_metrics("ref", ni, ni, xt[t * nh:(t + 1) * nh], yt[t * nh:(t + 1) * nh], lonc[t], latc[t])
        for k in xgtest.METRICS:
            assert np.array_equal(m[k], r[k]), k
        a, b = xgtest.grad_c2l("oracle", ni, ni, fh[t * nh:(t + 1) * nh], m)
        a2, b2 = xgtest.grad_c2l("ref", ni, ni, fh[t * nh:(t + 1) * nh], r)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
        gx[t * nc:(t + 1) * nc] = a; gy[t * nc:(t + 1) * nc] = b
        gm[t * nc:(t + 1) * nc] = xgtest.grad_mask(ni, ni, fh[t * nh:(t + 1) * nh], 1e20)
    for order in (1, 2):
        r = xgtest.ref_setup(lonc, latc, lon2, lat2, order, keep=True)
        for mono in ((False,) if order == 1 else (False, True)):
            data = f if order == 1 else fh
            want = xgtest.ref_apply(r["handle"], order, data, nlon * nlat, gx, gy, gm, has_missing=True, missing=1e20, monotonic=mono)
            got = xgtest.oracle_apply(r, order, tiles, data, nlon, nlat, gx, gy, gm, has_missing=True, missing=1e20, monotonic=mono)
            assert np.array_equal(got, want), (order, mono)
        reflib.ref_regrid_free(r["handle"])


def test_oracle_apply_variants_equal_compiled_reference(reflib, pkg):
    """weight field / cell_methods sum / cell_measures / --target_grid x missing x monotone x order: orc_conserve_apply_ex
    against the reference's do_scalar_conserve_interp, 60 combinations, bit for bit"""
    import itertools
    ni, nlon, nlat = 12, 40, 24
    lonc, latc, lont, latt = xgtest.ref_cubed_sphere(ni, centers=True)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    hm = xgtest.cubed_sphere_halo_map(lonc, latc)
    nc = ni * ni; nh = (ni + 2) ** 2
    tiles = [(ni, ni)] * 6
    rng = np.random.default_rng(5)
    f = rng.uniform(0.5, 1.5, 6 * nc); fm = f.copy(); fm[rng.uniform(size=f.size) < 0.08] = -999.0
    w = rng.uniform(0.2, 1.0, 6 * nc); fa = rng.uniform(1e9, 2e9, 6 * nc)
    n = 0
    for order in (1, 2):
        r = xgtest.ref_setup(lonc, latc, lon2, lat2, order, keep=True)
        ca = np.zeros(6 * nc); da = np.zeros(nlon * nlat)
        reflib.ref_regrid_cell_area(r["handle"], ca.ctypes.data, da.ctypes.data)
        for hmiss, cm, usew, usefa, tgt, mono in itertools.product((0, 1), repeat=6):
            if (cm and (usefa or tgt)) or (mono and order == 1):
                continue
            src = fm if hmiss else f
            data = src if order == 1 else xgtest.with_halo(src, hm)
            gx = rng.normal(size=6 * nc) * 0.1; gy = rng.normal(size=6 * nc) * 0.1
            gm = np.zeros(6 * nc, np.int32)
            if order == 2 and hmiss:
                gm = np.concatenate([xgtest.grad_mask(ni, ni, data[t * nh:(t + 1) * nh], -999.0) for t in range(6)])
            want = xgtest.ref_apply_ex(r["handle"], order, data, nlon * nlat, gx, gy, gm, bool(hmiss), -999.0, bool(mono), cm,
                                       w if usew else None, fa if usefa else None, -1e20, bool(tgt))
            got = xgtest.oracle_apply_ex(r, order, tiles, data, nlon, nlat, gx, gy, gm, bool(hmiss), -999.0, bool(mono), cm,
                                         w if usew else None, ca, fa if usefa else None, bool(tgt), da)
            assert np.array_equal(got, want), (order, hmiss, cm, usew, usefa, tgt, mono)
            n += 1
        reflib.ref_regrid_free(r["handle"])
    assert n == 60
