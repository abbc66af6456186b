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
