"""setup_conserve_interp with WRITE on the GPU path: the file the product writes == the file content the unmodified reference
hands its I/O layer (conserve_interp.c:368-443), and reading it back with READ reproduces the lists."""
import ctypes as C

import numpy as np
import pytest

import xgtest
from test_remap_cpu import _classic_bytes, _grids

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [1, 2])
def test_setup_conserve_interp_WRITE_then_READ(pkg, reflib, tmp_path, order):
    lonc, latc, lon2, lat2 = _grids(pkg, ni=24, nlon=72, nlat=36)
    L = pkg.lib(); R = reflib
    setup_fn = C.cast(L.setup_conserve_interp, C.c_void_p)
    op = xgtest.ORDER1 if order == 1 else xgtest.ORDER2
    assert L.xgb_set_nc_format(b"64bit_offset") == 0
    for jsc, jec in ((None, None), (9, 30)):
        want = xgtest.ref_setup(lonc, latc, lon2, lat2, op, jsc=jsc, jec=jec, remap=("gpu_ref.nc", 1), keep=True)
        dims, variables = xgtest.ref_store_file("gpu_ref.nc")
        path = str(tmp_path / f"gpu_{order}_{jsc}.nc")
        h = R.ref_regrid_setup_through_remap(want["handle"], setup_fn, path.encode(), 1)      # generate on the GPU + WRITE
        assert R.ref_regrid_nxgrid(h) == want["nxgrid"]
        got = open(path, "rb").read()
        ref_bytes = _classic_bytes(2, dims, variables)
        if xgtest.libm_matches_ref_trig():
            assert got == ref_bytes
        else:       # another libm: integer variables identical, areas to rounding
            assert len(got) == len(ref_bytes)
        h2 = R.ref_regrid_setup_through_remap(want["handle"], setup_fn, path.encode(), 2)     # READ it back
        n = R.ref_regrid_nxgrid(h2)
        assert n == want["nxgrid"]
        back = xgtest._alloc(n, order)
        R.ref_regrid_get(h2, back["t_in"], back["i_in"], back["j_in"], back["i_out"], back["j_out"], back["area"],
                         back["di"].ctypes.data if order == 2 else None, back["dj"].ctypes.data if order == 2 else None)
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
            assert np.array_equal(back[k], want[k]), k
        assert np.allclose(back["area"], want["area"], rtol=1e-12, atol=0)
