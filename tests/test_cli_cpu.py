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
