"""fregrid_b200 (fre-nctools_b200/cli/fregrid_b200.c) end to end on the GPU: mosaic + supergrid + field files in classic netCDF
(written here with scipy, an implementation independent of csrc/nc3.c) -> remap file and remapped fields, checked against the
oracle on the same grids.  Mirrors the reference's own CLI tests (t/Test20-fregrid.sh: run fregrid on a cubed-sphere mosaic,
look at the files it leaves behind), with numbers instead of `ncdump` existence checks."""
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

import xgtest

pytestmark = pytest.mark.gpu

D2R = np.pi / 180.0
R2D = 180.0 / np.pi
NI = 16
MISSING = np.float32(-1.0e10)


def _exe(pkg):
    path = os.path.join(os.path.dirname(pkg.__file__), "bin", "fregrid_b200")
    assert os.path.exists(path), "fregrid_b200 is not built (python __graft_entry__.py)"
    return path


def _contacts(hmap, n):
    """the mosaic's contact list from geometry (what make_solo_mosaic writes): for every pair of tiles sharing an edge, the
    edge of each in supergrid indices, reversed on the second side when the two edges run in opposite directions"""
    S = 2 * n
    def side(t, which):
        idx = {"W": hmap[t, 1:-1, 0], "E": hmap[t, 1:-1, n + 1], "S": hmap[t, 0, 1:-1], "N": hmap[t, n + 1, 1:-1]}[which]
        tt = idx // (n * n); jj, ii = np.divmod(idx % (n * n), n)
        assert np.all(tt == tt[0])
        if np.all(ii == 0): other, along = "W", jj
        elif np.all(ii == n - 1): other, along = "E", jj
        elif np.all(jj == 0): other, along = "S", ii
        else:
            assert np.all(jj == n - 1); other, along = "N", ii
        return int(tt[0]), other, bool(along[0] > along[-1])
    def rng(which, rev):
        a = f"{S}:1" if rev else f"1:{S}"
        return {"W": f"1:1,{a}", "E": f"{S}:{S},{a}", "S": f"{a},1:1", "N": f"{a},{S}:{S}"}[which]
    names, index, seen = [], [], set()
    for t in range(6):
        for which in "WESN":
            tt, other, rev = side(t, which)
            if (tt, other, t, which) in seen:
                continue
            seen.add((t, which, tt, other))
            names.append(f"C{n}_mosaic:tile{t + 1}::C{n}_mosaic:tile{tt + 1}")
            index.append(f"{rng(which, False)}::{rng(other, rev)}")
    assert len(names) == 12
    return names, index


def _strings(g, name, dim, values):
    v = g.createVariable(name, "c", (dim, "string"))
    for k, s in enumerate(values):
        v[k] = np.frombuffer(s.encode().ljust(255, b"\0"), "S1")


def _write_mosaic(pkg, d, n):
    """C<n>_mosaic.nc + C<n>_grid.tile[1-6].nc as make_hgrid / make_solo_mosaic leave them; returns what fregrid sees"""
    lonc, latc, lont, latt = pkg.cubed_sphere_grid(n, centers=True)
    hmap = xgtest.cubed_sphere_halo_map(lonc, latc)
    names, index = _contacts(hmap, n)
    g = netcdf_file(os.path.join(d, f"C{n}_mosaic.nc"), "w", version=2)
    g.createDimension("ntiles", 6); g.createDimension("ncontact", 12); g.createDimension("string", 255)
    m = g.createVariable("mosaic", "c", ("string",)); m[:] = np.frombuffer(f"C{n}_mosaic".encode().ljust(255, b"\0"), "S1")
    _strings(g, "gridfiles", "ntiles", [f"C{n}_grid.tile{t + 1}.nc" for t in range(6)])
    _strings(g, "gridtiles", "ntiles", [f"tile{t + 1}" for t in range(6)])
    _strings(g, "contacts", "ncontact", names)
    _strings(g, "contact_index", "ncontact", index)
    g.close()
    xdeg = np.zeros((6, 2 * n + 1, 2 * n + 1)); ydeg = np.zeros_like(xdeg)
    xdeg[:, ::2, ::2] = lonc * R2D; ydeg[:, ::2, ::2] = latc * R2D
    xdeg[:, 1::2, 1::2] = lont * R2D; ydeg[:, 1::2, 1::2] = latt * R2D
    for t in range(6):
        g = netcdf_file(os.path.join(d, f"C{n}_grid.tile{t + 1}.nc"), "w", version=2)
        g.createDimension("nx", 2 * n); g.createDimension("ny", 2 * n); g.createDimension("nxp", 2 * n + 1); g.createDimension("nyp", 2 * n + 1)
        x = g.createVariable("x", "d", ("nyp", "nxp")); y = g.createVariable("y", "d", ("nyp", "nxp"))
        x[:] = xdeg[t]; y[:] = ydeg[t]
        g.close()
    # what the tool sees after degrees -> radians (fregrid_util.c:227-241)
    return dict(lonc=xdeg[:, ::2, ::2] * D2R, latc=ydeg[:, ::2, ::2] * D2R, lont=xdeg[:, 1::2, 1::2] * D2R,
                latt=ydeg[:, 1::2, 1::2] * D2R, hmap=hmap)


@pytest.fixture(scope="module")
def dataset(pkg, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("cli"))
    n = NI
    grid = _write_mosaic(pkg, d, n)
    hmap = grid["hmap"]
    lonc_r, latc_r, lont_r, latt_r = grid["lonc"], grid["latc"], grid["lont"], grid["latt"]
    nt, nz = 2, 3
    rng = np.random.default_rng(7)
    temp = np.stack([[xgtest.smooth_field(lont_r, latt_r, k, t).reshape(6, n, n) for k in range(nz)] for t in range(nt)]).astype(np.float32)
    hole = rng.uniform(size=temp.shape) < 0.05
    temp[hole] = MISSING
    ps = np.stack([xgtest.smooth_field(lont_r, latt_r, 0, t).reshape(6, n, n) * 1000.0 for t in range(nt)])
    orog = rng.uniform(0, 3000, (6, n, n))
    for t in range(6):
        g = netcdf_file(os.path.join(d, f"atmos.tile{t + 1}.nc"), "w", version=2)
        g.createDimension("time", None); g.createDimension("pfull", nz); g.createDimension("grid_yt", n); g.createDimension("grid_xt", n)
        g.title = "synthetic"
        v = g.createVariable("time", "d", ("time",)); v.units = "days since 2000-01-01"; v.cartesian_axis = "T"
        p = g.createVariable("pfull", "d", ("pfull",)); p.units = "mb"; p.cartesian_axis = "Z"; p[:] = [100.0, 500.0, 900.0]
        yy = g.createVariable("grid_yt", "d", ("grid_yt",)); yy.cartesian_axis = "Y"; yy.units = "degrees_N"; yy[:] = np.arange(1.0, n + 1)
        xx = g.createVariable("grid_xt", "d", ("grid_xt",)); xx.cartesian_axis = "X"; xx.units = "degrees_E"; xx[:] = np.arange(1.0, n + 1)
        a = g.createVariable("temp", "f", ("time", "pfull", "grid_yt", "grid_xt")); a.missing_value = MISSING; a.units = "K"; a.long_name = "temperature"
        b = g.createVariable("ps", "d", ("time", "grid_yt", "grid_xt")); b.units = "Pa"
        c = g.createVariable("orog", "d", ("grid_yt", "grid_xt")); c.units = "m"
        c[:] = orog[t]
        for k in range(nt):
            v[k] = 10.0 + k; a[k] = temp[k, :, t]; b[k] = ps[k, t]
        g.close()
    return dict(dir=d, n=n, lonc=lonc_r, latc=latc_r, lont=lont_r, latt=latt_r, hmap=hmap, temp=temp, ps=ps, orog=orog, nt=nt, nz=nz)


def _run(pkg, ds, *args, ok=True):
    r = subprocess.run([_exe(pkg)] + list(args), cwd=ds["dir"], capture_output=True, text=True, timeout=600)
    assert (r.returncode == 0) == ok, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return r


def _remap_lists(pkg, path, order):
    from test_remap_cpu import _read
    return _read(pkg, path, order)


def test_weights_only_writes_the_reference_remap_file(pkg, dataset):
    """configs[0]: cubed sphere -> lat-lon, order-1 conservative, remap file written"""
    ds = dataset
    r = _run(pkg, ds, "--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", "60", "--nlat", "30", "--remap_file", "remap_o1", "--check_conserve")
    assert "only weight information is calculated" in r.stdout and "****remap_o1.nc" in r.stdout
    assert "The maximum ratio change" in r.stdout
    lon2, lat2 = pkg.latlon_grid(60, 30)
    want = xgtest.oracle_setup(ds["lonc"], ds["latc"], lon2, lat2, xgtest.ORDER1)
    got = _remap_lists(pkg, os.path.join(ds["dir"], "remap_o1.nc"), 1)
    xgtest.assert_xgrid_equal(got, want, 1, area_tol=1e-12, exact=False)     # areas read back are (a / 4 pi R^2) * 4 pi R^2
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(got[k], want[k])
    g = netcdf_file(os.path.join(ds["dir"], "remap_o1.nc"), "r", mmap=False)
    assert list(g.variables) == ["tile1", "tile1_cell", "tile2_cell", "xgrid_area"] and g.version_byte == 2
    g.close()


def _expected(pkg, ds, order, nlon, nlat):
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    x = xgtest.oracle_setup(ds["lonc"], ds["latc"], lon2, lat2, xgtest.ORDER2 if order == 2 else xgtest.ORDER1)
    n = ds["n"]; tiles = [(n, n)] * 6
    metrics = None
    if order == 2:
        nh = (n + 2) ** 2
        xt = xgtest.with_halo(ds["lont"].reshape(-1), ds["hmap"]); yt = xgtest.with_halo(ds["latt"].reshape(-1), ds["hmap"])
        metrics = [xgtest.c2l_metrics("oracle", n, n, xt[t * nh:(t + 1) * nh], yt[t * nh:(t + 1) * nh], ds["lonc"][t], ds["latc"][t]) for t in range(6)]

    def remap(field, has_missing, missing):
        flat = np.asarray(field, np.float64).reshape(-1)
        if order == 1:
            return xgtest.oracle_apply(x, 1, tiles, flat, nlon, nlat, has_missing=has_missing, missing=missing)
        nc, nh = n * n, (n + 2) ** 2
        fh = xgtest.with_halo(flat, ds["hmap"])
        gx = np.zeros(6 * nc); gy = np.zeros(6 * nc); gm = np.zeros(6 * nc, np.int32)
        for t in range(6):
            gx[t * nc:(t + 1) * nc], gy[t * nc:(t + 1) * nc] = xgtest.grad_c2l("oracle", n, n, fh[t * nh:(t + 1) * nh], metrics[t])
            if has_missing:
                gm[t * nc:(t + 1) * nc] = xgtest.grad_mask(n, n, fh[t * nh:(t + 1) * nh], missing)
        return xgtest.oracle_apply(x, 2, tiles, fh, nlon, nlat, grad_x=gx, grad_y=gy, gmask=gm, has_missing=has_missing, missing=missing)
    return remap


def test_order1_fields_written_then_remap_file_read_back(pkg, dataset):
    ds = dataset
    nlon, nlat = 48, 24
    args = ["--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", str(nlon), "--nlat", str(nlat), "--input_file", "atmos",
            "--scalar_field", "temp,ps,orog", "--output_file", "out1.nc", "--remap_file", "remap_f1.nc"]
    r1 = _run(pkg, ds, *args)
    assert "done calculating index and weight" in r1.stdout and "****out1.nc" in r1.stdout
    out = os.path.join(ds["dir"], "out1.nc")
    g = netcdf_file(out, "r", mmap=False)
    assert g.dimensions["grid_xt"] == nlon and g.dimensions["grid_yt"] == nlat and g.dimensions["pfull"] == ds["nz"] and g.dimensions["bnds"] == 2
    assert g.variables["time"].shape == (ds["nt"],) and np.array_equal(g.variables["time"][:], [10.0, 11.0])
    assert np.allclose(g.variables["grid_xt"][:], (np.arange(nlon) + 0.5) * 360.0 / nlon, rtol=0, atol=1e-12)
    assert np.allclose(g.variables["grid_yt"][:], -90 + (np.arange(nlat) + 0.5) * 180.0 / nlat, rtol=0, atol=1e-12)
    assert g.variables["grid_xt"].bounds == b"grid_xt_bnds" and np.allclose(g.variables["grid_xt_bnds"][:, 1], (np.arange(nlon) + 1) * 360.0 / nlon, atol=1e-12)
    assert g.variables["temp"].interp_method == b"conserve_order1" and g.variables["temp"].units == b"K" and g.title == b"synthetic"
    assert b"--scalar_field temp,ps,orog" in g.history
    remap = _expected(pkg, ds, 1, nlon, nlat)
    for t in range(ds["nt"]):
        want = remap(ds["ps"][t], False, 0.0)
        assert np.array_equal(g.variables["ps"][t].reshape(-1), want)                     # double field: bit for bit
        for k in range(ds["nz"]):
            want = remap(ds["temp"][t, k].astype(np.float64), True, float(MISSING))
            assert np.array_equal(g.variables["temp"][t, k].reshape(-1), want.astype(np.float32))
    assert np.array_equal(g.variables["orog"][:].reshape(-1), remap(ds["orog"], False, 0.0))
    ps_first = g.variables["ps"][:].copy(); temp_first = g.variables["temp"][:].copy()
    g.close()
    size_first = os.path.getsize(out)
    r2 = _run(pkg, ds, *args)                                  # the remap file exists now: READ branch
    assert "Finish reading index and weight" in r2.stdout
    g = netcdf_file(out, "r", mmap=False)
    # (areas read back are (a / 4 pi R^2) * 4 pi R^2 like the reference's, so values agree to rounding, not bit for bit)
    assert np.allclose(g.variables["ps"][:], ps_first, rtol=1e-14, atol=0)
    assert np.allclose(g.variables["temp"][:], temp_first, rtol=2e-7, atol=0)
    g.close()
    assert size_first == os.path.getsize(out)


def test_order2_fields_with_missing_values(pkg, dataset):
    """configs[1] in small: order-2 remap with gradient terms of a 3-D field over levels and times; halos come from the mosaic
    contacts, gradient metrics from the device (rounding-level differences from the oracle's)"""
    ds = dataset
    nlon, nlat = 64, 32
    _run(pkg, ds, "--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", str(nlon), "--nlat", str(nlat), "--input_file", "atmos.nc",
         "--scalar_field", "ps,temp", "--output_file", "out2", "--interp_method", "conserve_order2", "--format", "classic")
    g = netcdf_file(os.path.join(ds["dir"], "out2.nc"), "r", mmap=False)
    assert g.version_byte == 1 and g.variables["ps"].interp_method == b"conserve_order2"
    remap = _expected(pkg, ds, 2, nlon, nlat)
    for t in range(ds["nt"]):
        want = remap(ds["ps"][t], False, 0.0)
        got = g.variables["ps"][t].reshape(-1)
        assert np.allclose(got, want, rtol=1e-13, atol=0), np.abs(got / want - 1).max()
        for k in range(ds["nz"]):
            want = remap(ds["temp"][t, k].astype(np.float64), True, float(MISSING))
            got = g.variables["temp"][t, k].reshape(-1).astype(np.float64)
            same_holes = (got == float(MISSING)) == (want == float(MISSING))
            assert same_holes.all()
            ok = want != float(MISSING)
            assert np.allclose(got[ok], want[ok], rtol=3e-7, atol=0)        # float output
    g.close()


def test_klevel_lstep_windows_and_standard_dimension(pkg, dataset):
    ds = dataset
    _run(pkg, ds, "--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", "36", "--nlat", "18", "--input_file", "atmos", "--scalar_field", "temp",
         "--output_file", "out3", "--KlevelBegin", "2", "--KlevelEnd", "3", "--LstepBegin", "2", "--LstepEnd", "2", "--standard_dimension")
    g = netcdf_file(os.path.join(ds["dir"], "out3.nc"), "r", mmap=False)
    assert g.dimensions["lon"] == 36 and g.dimensions["lat"] == 18 and g.dimensions["pfull"] == 2
    assert g.variables["lon"].units == b"degrees_E" and g.variables["lon"].bounds == b"lon_bnds" and g.variables["lat_bnds"].shape == (18, 2)
    assert np.array_equal(g.variables["pfull"][:], [500.0, 900.0]) and np.array_equal(g.variables["time"][:], [11.0])
    remap = _expected(pkg, ds, 1, 36, 18)
    want = remap(ds["temp"][1, 2].astype(np.float64), True, float(MISSING))
    assert np.array_equal(g.variables["temp"][0, 1].reshape(-1), want.astype(np.float32))
    g.close()


def test_reference_error_messages(pkg, dataset):
    ds = dataset
    mosaic = f"C{ds['n']}_mosaic.nc"
    cases = [
        (["--nlon", "10", "--nlat", "5"], "input_mosaic is not specified"),
        (["--input_mosaic", mosaic], "nlon and nlat should be specified"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5"], "remap_file must be specified to save weight information"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--interp_method", "nearest"], "interp_method must be"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5", "--input_file", "atmos"], "both scalar_field and vector_field are not specified"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5", "--input_file", "atmos", "--scalar_field", "nope"], "variable nope"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5", "--remap_file", "r", "--format", "netcdf4"], "HDF5"),
        (["--input_mosaic", mosaic, "--nlon", "10", "--nlat", "5", "--input_file", "atmos", "--u_field", "u", "--v_field", "v"],
         "conservative interpolation of vector fields is not supported"),
    ]
    for args, msg in cases:
        r = _run(pkg, ds, *args, ok=False)
        assert r.returncode == 1 and "FATAL Error" in r.stderr and msg in r.stderr, (args, r.stderr)


def test_output_mosaic_with_several_tiles(pkg, dataset):
    """--output_mosaic: cubed sphere onto another cubed sphere; one remap file and one output file per destination tile, named
    after the mosaic's gridtiles (set_remap_file, set_mosaic_data_file); a curvilinear destination takes the pyramid search"""
    ds = dataset
    m = 10
    out = _write_mosaic(pkg, ds["dir"], m)
    args = ["--input_mosaic", f"C{ds['n']}_mosaic.nc", "--output_mosaic", f"C{m}_mosaic.nc", "--input_file", "atmos", "--scalar_field", "ps",
            "--output_file", "cs_out", "--remap_file", "cs_remap"]
    r = _run(pkg, ds, *args)
    for t in range(6):
        assert f"****cs_out.tile{t + 1}.nc" in r.stdout
    n = ds["n"]; tiles = [(n, n)] * 6
    for t in (0, 2, 5):
        x = xgtest.oracle_setup(ds["lonc"], ds["latc"], out["lonc"][t], out["latc"][t], xgtest.ORDER1)
        got = _remap_lists(pkg, os.path.join(ds["dir"], f"cs_remap.tile{t + 1}.nc"), 1)
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
            assert np.array_equal(got[k], x[k]), (t, k)
        g = netcdf_file(os.path.join(ds["dir"], f"cs_out.tile{t + 1}.nc"), "r", mmap=False)
        assert g.dimensions["grid_xt"] == m and g.dimensions["grid_yt"] == m and "grid_xt_bnds" not in g.variables
        for k in range(ds["nt"]):
            want = xgtest.oracle_apply(x, 1, tiles, ds["ps"][k].reshape(-1), m, m)
            assert np.array_equal(g.variables["ps"][k].reshape(-1), want), (t, k)
        g.close()
    r2 = _run(pkg, ds, *args)                                  # all six remap files exist now: READ
    assert "Finish reading index and weight" in r2.stdout


def test_great_circle_grid_file_selects_the_great_circle_generator(pkg, reflib, tmp_path):
    """configs[2] in small: a tripolar ocean grid whose grid file carries great_circle_algorithm = "TRUE" (get_great_circle_algorithm,
    fregrid.c:757-765) onto a lat-lon grid: the remap file holds the great-circle exchange grid; order 2 is refused like the
    reference does"""
    d = str(tmp_path)
    tl, ta = xgtest.tripolar_grid(120, 80)                     # model grid 60 x 40
    ny, nx = tl.shape[0] - 1, tl.shape[1] - 1
    g = netcdf_file(os.path.join(d, "ocean_mosaic.nc"), "w", version=1)
    g.createDimension("ntiles", 1); g.createDimension("string", 255)
    _strings(g, "gridfiles", "ntiles", ["ocean_hgrid.nc"]); _strings(g, "gridtiles", "ntiles", ["tile1"])
    g.close()
    xdeg = np.zeros((2 * ny + 1, 2 * nx + 1)); ydeg = np.zeros_like(xdeg)
    xdeg[::2, ::2] = tl * R2D; ydeg[::2, ::2] = ta * R2D
    g = netcdf_file(os.path.join(d, "ocean_hgrid.nc"), "w", version=1)
    g.createDimension("nx", 2 * nx); g.createDimension("ny", 2 * ny); g.createDimension("nxp", 2 * nx + 1); g.createDimension("nyp", 2 * ny + 1)
    g.great_circle_algorithm = "TRUE"
    x = g.createVariable("x", "d", ("nyp", "nxp")); y = g.createVariable("y", "d", ("nyp", "nxp"))
    x[:] = xdeg; y[:] = ydeg
    g.close()
    ds = {"dir": d}
    _run(pkg, ds, "--input_mosaic", "ocean_mosaic.nc", "--nlon", "36", "--nlat", "18", "--remap_file", "gc_remap")
    lon2, lat2 = pkg.latlon_grid(36, 18)
    want = xgtest.oracle_setup([xdeg[::2, ::2] * D2R], [ydeg[::2, ::2] * D2R], lon2, lat2, xgtest.ORDER1 | xgtest.GREAT_CIRCLE)
    got = _remap_lists(pkg, os.path.join(d, "gc_remap.nc"), 1)
    assert got["nxgrid"] == want["nxgrid"] and want["nxgrid"] > 0
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(got[k], want[k]), k
    assert np.allclose(got["area"], want["area"], rtol=0, atol=8e-15 * 6371000.0 ** 2)     # DESIGN.md section 2, great circle
    r = _run(pkg, ds, "--input_mosaic", "ocean_mosaic.nc", "--nlon", "36", "--nlat", "18", "--remap_file", "gc2", "--interp_method",
             "conserve_order2", ok=False)
    assert "can not be conserve_order2" in r.stderr


def test_gpus_option_writes_the_same_remap_file(pkg, dataset):
    """fregrid_b200 --gpus / --gpu_list (one process, several devices, csrc/multi_gpu.cu; replaces mpirun fregrid_parallel):
    the remap file is byte-identical to the one-device file, order 1 and order 2, lat-lon and mosaic destinations.  On a
    one-GPU box the same device is named twice — two plans, two host threads, the whole window / offset / gather logic."""
    import torch
    ds = dataset
    ndev = torch.cuda.device_count()
    lists = ["0,0", "0,0,0"] + (["0,1"] if ndev > 1 else [])
    for method, tag in (("conserve_order1", "o1"), ("conserve_order2", "o2")):
        base = ["--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", "72", "--nlat", "36", "--interp_method", method]
        _run(pkg, ds, *base, "--remap_file", f"one_{tag}")
        want = open(os.path.join(ds["dir"], f"one_{tag}.nc"), "rb").read()
        for k, gl in enumerate(lists):
            _run(pkg, ds, *base, "--remap_file", f"multi_{tag}_{k}", "--gpu_list", gl)
            assert open(os.path.join(ds["dir"], f"multi_{tag}_{k}.nc"), "rb").read() == want, (tag, gl)
        if ndev > 1:
            _run(pkg, ds, *base, "--remap_file", f"multi_{tag}_n", "--gpus", str(min(ndev, 8)))
            assert open(os.path.join(ds["dir"], f"multi_{tag}_n.nc"), "rb").read() == want, tag
    # a curvilinear destination tile by tile
    m = 10
    _write_mosaic(pkg, ds["dir"], m)
    base = ["--input_mosaic", f"C{ds['n']}_mosaic.nc", "--output_mosaic", f"C{m}_mosaic.nc"]
    _run(pkg, ds, *base, "--remap_file", "cs_one")
    _run(pkg, ds, *base, "--remap_file", "cs_two", "--gpu_list", "0,0")
    for t in range(6):
        a = open(os.path.join(ds["dir"], f"cs_one.tile{t + 1}.nc"), "rb").read()
        assert a == open(os.path.join(ds["dir"], f"cs_two.tile{t + 1}.nc"), "rb").read(), t


def test_order2_onto_several_output_tiles_equals_the_reference(pkg, dataset):
    """conserve_order2 with --output_mosaic: tile1_distance of every output tile's remap file equals the reference's
    setup_conserve_interp over all six output tiles at once (per-source-cell sums across the tiles, conserve_interp.c:204-221)"""
    if xgtest.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    ds = dataset
    m = 10
    out = _write_mosaic(pkg, ds["dir"], m)
    _run(pkg, ds, "--input_mosaic", f"C{ds['n']}_mosaic.nc", "--output_mosaic", f"C{m}_mosaic.nc", "--remap_file", "cs2", "--interp_method", "conserve_order2")
    h, ref = xgtest.ref_multi_setup(ds["lonc"], ds["latc"], out["lonc"], out["latc"], 2)
    for t in range(6):
        got = _remap_lists(pkg, os.path.join(ds["dir"], f"cs2.tile{t + 1}.nc"), 2)
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "di", "dj"):
            assert np.array_equal(got[k], ref[t][k]), (t, k)
        assert np.allclose(got["area"], ref[t]["area"], rtol=1e-15, atol=0), t     # the reader rescales: (a / 4 pi R^2) * 4 pi R^2


def test_fregrid_b200_against_the_unmodified_reference_fregrid(pkg, dataset):
    """The drop-in claim at the command line: oracle/_ref/fregrid_ref — the reference's own fregrid (main, option parsing, mosaic
    and field readers, mpp_io, conserve_interp, writers) compiled unmodified over the netCDF-C shim — and fregrid_b200 run with
    the same arguments on the same files.  Remap files: same cells, same order, bit-identical areas and distances (byte-
    identical when the container versions agree).  Output files: same variables, dimensions and attributes; order 1 values
    identical; order 2 (the product computes the c2l metrics on the device: bit-identical but for asin / atan2 on about one
    argument per thousand) to 1e-13."""
    ref = os.path.join(xgtest.ORACLE_DIR, "_ref", "fregrid_ref")
    xgtest.ref_lib()
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/fregrid_ref not built")
    ds = dataset
    for method, order, tol in (("conserve_order1", 1, 0.0), ("conserve_order2", 2, 1e-13)):
        common = ["--input_mosaic", f"C{ds['n']}_mosaic.nc", "--nlon", "48", "--nlat", "24", "--input_file", "atmos",
                  "--scalar_field", "temp,ps,orog", "--interp_method", method]
        r = subprocess.run([ref] + common + ["--output_file", f"cmp_ref_{order}.nc", "--remap_file", f"cmp_ref_remap_{order}.nc"],
                           cwd=ds["dir"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1000:])
        _run(pkg, ds, *common, "--output_file", f"cmp_b200_{order}.nc", "--remap_file", f"cmp_b200_remap_{order}.nc")
        a = open(os.path.join(ds["dir"], f"cmp_ref_remap_{order}.nc"), "rb").read()
        b = open(os.path.join(ds["dir"], f"cmp_b200_remap_{order}.nc"), "rb").read()
        if a[3] == b[3]:
            assert a == b, f"remap files differ (order {order})"
        else:
            x = _remap_lists(pkg, os.path.join(ds["dir"], f"cmp_ref_remap_{order}.nc"), order)
            y = _remap_lists(pkg, os.path.join(ds["dir"], f"cmp_b200_remap_{order}.nc"), order)
            for k in x:
                assert np.array_equal(x[k], y[k]), (order, k)
        ga = netcdf_file(os.path.join(ds["dir"], f"cmp_ref_{order}.nc"), "r", mmap=False)
        gb = netcdf_file(os.path.join(ds["dir"], f"cmp_b200_{order}.nc"), "r", mmap=False)
        assert list(ga.variables) == list(gb.variables), (list(ga.variables), list(gb.variables))
        assert dict(ga.dimensions) == dict(gb.dimensions)
        for name, va in ga.variables.items():
            vb = gb.variables[name]
            assert va.dimensions == vb.dimensions and va.typecode() == vb.typecode(), name
            assert {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in va._attributes.items()} == \
                   {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in vb._attributes.items()}, name
            xa, xb = np.asarray(va[:], np.float64), np.asarray(vb[:], np.float64)
            if tol == 0.0 or name not in ("temp", "ps", "orog"):
                assert np.array_equal(xa, xb), (order, name)
            else:
                m = (xa != MISSING) | (xb != MISSING)
                assert np.array_equal(xa == MISSING, xb == MISSING), name
                assert np.max(np.abs(xa[m] - xb[m]) / np.maximum(np.abs(xa[m]), 1e-30), initial=0.0) <= tol, (order, name)
        ga.close(); gb.close()


def test_netcdf4_inputs_give_the_same_output_files(pkg, dataset, tmp_path):
    """The reference's default file format (NC_FORMAT_NETCDF4_CLASSIC, mpp_io.c:52): the mosaic, the six supergrids and the six
    field files re-expressed as netCDF-4 by tests/h5_writer.py (both libnetcdf layouts, the field files chunked, shuffled and
    deflated like FRE history files) -> the remap file byte-identical and the output file equal (all but the time in its history) to the run on the classic files; then the remap
    file itself as netCDF-4 through the READ branch."""
    import h5_writer
    ds = dataset
    n = ds["n"]
    d4 = str(tmp_path)
    names = [f"C{n}_mosaic.nc"] + [f"C{n}_grid.tile{t + 1}.nc" for t in range(6)] + [f"atmos.tile{t + 1}.nc" for t in range(6)]
    for k, name in enumerate(names):
        field = name.startswith("atmos")
        h5_writer.from_classic(os.path.join(ds["dir"], name), os.path.join(d4, name), style=("v18", "earliest")[k % 2],
                               chunk=16 if field else None, deflate=3 if field else 0, shuffle=field)
        assert open(os.path.join(d4, name), "rb").read(4) == b"\x89HDF"
    args = ["--input_mosaic", f"C{n}_mosaic.nc", "--nlon", "48", "--nlat", "24", "--input_file", "atmos", "--scalar_field", "temp,ps,orog",
            "--interp_method", "conserve_order2", "--output_file", "o_h5.nc", "--remap_file", "r_h5.nc"]

    def run(cwd, a):
        r = subprocess.run([_exe(pkg)] + a, cwd=cwd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
        return r.stdout
    run(ds["dir"], args)
    run(d4, args)
    assert open(os.path.join(ds["dir"], "r_h5.nc"), "rb").read() == open(os.path.join(d4, "r_h5.nc"), "rb").read()
    # the output file carries the time of the run in its history attribute: everything else must be equal
    g3 = netcdf_file(os.path.join(ds["dir"], "o_h5.nc"), "r", mmap=False); g4 = netcdf_file(os.path.join(d4, "o_h5.nc"), "r", mmap=False)
    assert g3.dimensions == g4.dimensions and list(g3.variables) == list(g4.variables) and g3.title == g4.title
    for name, v3 in g3.variables.items():
        v4 = g4.variables[name]
        assert v3.dimensions == v4.dimensions and v3._attributes == v4._attributes and v3.data.dtype == v4.data.dtype, name
        assert np.array_equal(v3[...], v4[...]), name
    g3.close(); g4.close()
    assert os.path.getsize(os.path.join(ds["dir"], "o_h5.nc")) == os.path.getsize(os.path.join(d4, "o_h5.nc"))
    # READ branch: the same remap file as netCDF-4 against the classic one
    h5_writer.from_classic(os.path.join(d4, "r_h5.nc"), os.path.join(d4, "r4.nc"), style="v18", chunk=4096, deflate=1)
    a_read = args[:-3] + ["o_read3.nc", "--remap_file", "r_h5.nc"]
    b_read = args[:-3] + ["o_read4.nc", "--remap_file", "r4.nc"]
    assert "Finish reading index and weight" in run(d4, a_read)
    assert "Finish reading index and weight" in run(d4, b_read)
    g3 = netcdf_file(os.path.join(d4, "o_read3.nc"), "r", mmap=False); g4 = netcdf_file(os.path.join(d4, "o_read4.nc"), "r", mmap=False)
    for v in ("temp", "ps", "orog"):
        assert np.array_equal(g3.variables[v][:], g4.variables[v][:]), v
    g3.close(); g4.close()
