"""The C-ABI library loads on a CPU-only box and exports every symbol include/xgrid_b200.h declares;
without a GPU the entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np

import xgtest


def _declared_symbols():
    text = open(os.path.join(xgtest.ROOT, "include", "xgrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text))
    return sorted(n for n in names if n.startswith(("xgb_", "create_xgrid", "get_", "setup_conserve", "do_scalar_conserve")) and n != "xgb_plan")


def test_all_declared_symbols_are_exported(pkg):
    L = pkg.lib()
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/xgrid_b200.h but not exported"


def test_no_cpu_fallback(pkg):
    L = pkg.lib()
    if L.xgb_device_count() > 0:
        return                      # on a GPU box the real path is exercised by the -m gpu tests
    p = L.xgb_plan_create(0)
    assert not p
    assert b"no CUDA device" in L.xgb_last_error()
    try:
        pkg.XgridPlan(0)
    except pkg.XgridError as e:
        assert "no CPU path" in str(e)
    else:
        raise AssertionError("XgridPlan must raise without a GPU")


def test_host_side_grid_helpers(pkg):
    lon, lat = pkg.latlon_grid(8, 4)
    assert lon.shape == (5, 9) and abs(lon[0, -1] - 2 * np.pi) < 1e-15 and abs(lat[-1, 0] - np.pi / 2) < 1e-15
    lonc, latc = pkg.cubed_sphere_grid(4)
    assert lonc.shape == (6, 5, 5)
    # tile 3 (index 2) holds the north pole at its centre vertex, tile 6 the south pole (create_gnomonic_cubic_grid.c:1691-1743)
    assert latc[2, 2, 2] == np.pi / 2 and latc[5, 2, 2] == -np.pi / 2


def test_struct_layout_matches_reference(pkg, reflib):
    """csrc/fregrid_abi.h mirrors Grid_config / Interp_config / Field_config / Var_config: sizeof and the offsets of every member
    the library reads equal the compiled reference's (oracle/ref_driver.c ref_abi_layout, built from the real headers)"""
    L = pkg.lib()
    L.xgb_abi_layout.argtypes = [C.POINTER(C.c_size_t), C.c_int]
    a = (C.c_size_t * 64)(); b = (C.c_size_t * 64)()
    na = L.xgb_abi_layout(a, 64); nb = reflib.ref_abi_layout(b, 64)
    assert na == nb and na >= 30
    assert list(a)[:na] == list(b)[:nb]


def _moment_polygons(rng, count):
    """convex-ish polygons of 3..8 vertices at every latitude, with the special edges the routines branch on: meridian edges
    (dx == 0), parallels (dy == 0), a side through a pole (|dx| == pi), longitudes straddling +-pi of the centre"""
    for c in range(count):
        n = int(rng.integers(3, 9))
        lat0 = rng.uniform(-1.55, 1.55)
        lon0 = rng.uniform(-1.0, 7.0)
        r = 10.0 ** rng.uniform(-4, -0.5)
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        x = lon0 + r * np.cos(ang) / max(np.cos(lat0), 0.05)
        y = np.clip(lat0 + r * np.sin(ang), -np.pi / 2, np.pi / 2)
        kind = c % 8
        if kind == 1:
            x[1] = x[0]                      # meridian edge
        elif kind == 2:
            y[2 % n] = y[1]                  # parallel
        elif kind == 3:
            x[1] = x[0] + np.pi              # side through a pole
        elif kind == 4:
            x[n - 1] = x[0]; y[n - 1] = y[0] + 1e-11   # nearly flat, not moving
        clon = lon0 + (rng.uniform(-4, 4) if kind == 5 else rng.uniform(-r, r))
        yield n, np.ascontiguousarray(x), np.ascontiguousarray(y), clon


def test_poly_moments_site_host_build_equals_oracle(pkg):
    """csrc/xgrid_geom.cuh poly_moments_site (the clip kernel's one-pass area + centroid sums, host build) against the
    oracle's poly_area / poly_ctrlon / poly_ctrlat: bit for bit"""
    L = pkg.lib()
    O = xgtest.oracle_lib()
    rng = np.random.default_rng(2024)
    out = np.zeros(3)
    bad = 0
    for n, x, y, clon in _moment_polygons(rng, 20000):
        want = (O.orc_poly_area(x, y, n), O.orc_poly_ctrlon(x, y, n, clon), O.orc_poly_ctrlat(x, y, n))
        for order in (1, 2):
            L.xgb_poly_moments_site_host(order, n, x.ctypes.data, y.ctypes.data, clon, out.ctypes.data)
            m = 3 if order == 2 else 1
            if not all(np.float64(out[k]).view(np.uint64) == np.float64(want[k]).view(np.uint64) for k in range(m)):
                bad += 1
                assert bad < 1, (order, n, x, y, clon, out, want)
