"""make_coupler_mosaic's exchange grids on the GPU (csrc/coupler.cu, xgb_make_coupler_xgrid) against
the UNMODIFIED reference tool (oracle/_ref/make_coupler_mosaic_ref = tools/make_coupler_mosaic/make_coupler_mosaic.c compiled
over the netCDF-C shim) run on the same mosaic files: every exchange-grid file it writes (parent cells in order, xgrid_area,
tile1_distance, tile2_distance) and its land_mask / ocean_mask files, bit for bit."""
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

import xgtest
from test_cli_gpu import _write_mosaic, _strings

pytestmark = pytest.mark.gpu

D2R = np.pi / 180.0          # constant.h: D2R (M_PI/180.)


def _ref_tool():
    xgtest.ref_lib()
    path = os.path.join(xgtest.ORACLE_DIR, "_ref", "make_coupler_mosaic_ref")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/make_coupler_mosaic_ref not built")
    return path


def _write_ocean(d, nx, ny, lat0, seed, frac_land=0.4, area_frac=False):
    """a one-tile ocean mosaic the way make_hgrid / make_solo_mosaic / make_topog leave it: a sheared, latitude-stretched
    supergrid from -280 to 80 degrees east (MOM's range: every cell is fix_lon'd by +2 pi first), starting at lat0 > -90 so the
    tool adds its artificial southern row, and a depth field with land points"""
    g = netcdf_file(os.path.join(d, "ocean_mosaic.nc"), "w", version=2)
    g.createDimension("ntiles", 1); g.createDimension("string", 255)
    m = g.createVariable("mosaic", "c", ("string",)); m[:] = np.frombuffer(b"ocean_mosaic".ljust(255, b"\0"), "S1")
    _strings(g, "gridfiles", "ntiles", ["ocean_hgrid.nc"])
    _strings(g, "gridtiles", "ntiles", ["tile1"])
    g.close()
    xs = np.linspace(-280.0, 80.0, 2 * nx + 1)
    t = np.linspace(0.0, 1.0, 2 * ny + 1)
    ys = lat0 + (90.0 - lat0) * (0.65 * t + 0.35 * t * t)
    ys[-1] = 90.0
    x = xs[None, :] + 1.7 * np.sin(ys * D2R * 2.0)[:, None]
    y = np.repeat(ys[:, None], 2 * nx + 1, axis=1)
    g = netcdf_file(os.path.join(d, "ocean_hgrid.nc"), "w", version=2)
    g.createDimension("nx", 2 * nx); g.createDimension("ny", 2 * ny); g.createDimension("nxp", 2 * nx + 1); g.createDimension("nyp", 2 * ny + 1)
    vx = g.createVariable("x", "d", ("nyp", "nxp")); vy = g.createVariable("y", "d", ("nyp", "nxp"))
    vx[:] = x; vy[:] = y
    g.close()
    rng = np.random.default_rng(seed)
    depth = np.where(rng.uniform(size=(ny, nx)) < frac_land, 0.0, rng.uniform(10.0, 5000.0, (ny, nx)))
    g = netcdf_file(os.path.join(d, "topog.nc"), "w", version=2)
    g.createDimension("nx", nx); g.createDimension("ny", ny)
    v = g.createVariable("depth", "d", ("ny", "nx")); v[:] = depth
    frac = (depth > 0.0).astype(np.float64)
    if area_frac:
        # partly wet cells (make_coupler_mosaic.c:941-947): both the sea and the land branch run on them; values at and
        # around MIN_AREA_FRAC
        u = rng.uniform(size=(ny, nx))
        frac = np.where(u < 0.25, rng.uniform(0.0, 1.0, (ny, nx)), frac)
        frac = np.where((u >= 0.25) & (u < 0.30), 1.0e-4, frac)
        frac = np.where((u >= 0.30) & (u < 0.35), 1.0 - 1.0e-4, frac)
        frac = np.where((u >= 0.35) & (u < 0.40), 5.0e-5, frac)
        w = g.createVariable("area_frac", "d", ("ny", "nx")); w[:] = frac
    g.close()
    # what the tool holds (make_coupler_mosaic.c:834-876, :939-955): model-grid vertices in radians, the extra southern row,
    # omask = depth > sea_level
    lon = x[::2, ::2] * D2R
    lat = y[::2, ::2] * D2R
    ext = 1 if ys[0] * D2R > -90.0 * D2R + 1.0e-7 else 0
    if ext:
        lon = np.vstack([lon[:1], lon])
        lat = np.vstack([np.full((1, nx + 1), -90.0 * D2R), lat])
    omask = np.zeros((ny + ext, nx))
    omask[ext:] = frac
    return dict(lon=lon, lat=lat, omask=omask, ext=ext, nx=nx, ny=ny)


def _tiles(grid):
    return [(grid["lonc"][t], grid["latc"][t]) for t in range(6)]


def _read_xgrid(path, order):
    g = netcdf_file(path, "r", mmap=False)
    d = dict(c1=np.array(g.variables["tile1_cell"][:]), c2=np.array(g.variables["tile2_cell"][:]), area=np.array(g.variables["xgrid_area"][:]))
    if order == 2:
        d["d1"] = np.array(g.variables["tile1_distance"][:]); d["d2"] = np.array(g.variables["tile2_distance"][:])
    g.close()
    return d


def _compare_lists(d, lst, name1, name2, order, ext2=0):
    """every file <name1>_tileAX<name2>_tileB.nc of the tool equals the (A, B) sub-list, and there is no other file"""
    seen, checked = set(), 0
    for t1 in np.unique(lst["t1"]):
        for t2 in np.unique(lst["t2"]):
            m = (lst["t1"] == t1) & (lst["t2"] == t2)
            path = os.path.join(d, f"{name1}_tile{t1 + 1}X{name2}_tile{t2 + 1}.nc")
            if not m.any():
                assert not os.path.exists(path), path
                continue
            seen.add(os.path.basename(path))
            ref = _read_xgrid(path, order)
            assert ref["area"].shape[0] == int(m.sum()), (path, ref["area"].shape[0], int(m.sum()))
            assert np.array_equal(ref["c1"][:, 0], lst["i1"][m] + 1) and np.array_equal(ref["c1"][:, 1], lst["j1"][m] + 1), path
            assert np.array_equal(ref["c2"][:, 0], lst["i2"][m] + 1) and np.array_equal(ref["c2"][:, 1], lst["j2"][m] + 1 - ext2), path
            assert np.array_equal(ref["area"], lst["area"][m]), (path, np.max(np.abs(ref["area"] - lst["area"][m]) / ref["area"]))
            if order == 2:
                for k, col, r in (("d1i", 0, "d1"), ("d1j", 1, "d1"), ("d2i", 0, "d2"), ("d2j", 1, "d2")):
                    assert np.array_equal(ref[r][:, col], lst[k][m]), (path, k, np.max(np.abs(ref[r][:, col] - lst[k][m])))
            checked += int(m.sum())
    others = {f for f in os.listdir(d) if f.startswith(f"{name1}_tile") and f"X{name2}_tile" in f}
    assert others == seen, (others - seen, seen - others)
    return checked


@pytest.mark.parametrize("order", [2, 1])
def test_three_mosaics_against_the_reference_tool(pkg, tmp_path, order):
    """C8 atmosphere, C6 land, 30 x 20 sheared ocean with 40 % land points and the artificial southern row"""
    tool = _ref_tool()
    d = str(tmp_path)
    atm = _write_mosaic(pkg, d, 8)
    lnd = _write_mosaic(pkg, d, 6)
    ocn = _write_ocean(d, 30, 20, -78.0, seed=3)
    r = subprocess.run([tool, "--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C6_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                        "--ocean_topog", "topog.nc", "--interp_order", str(order), "--mosaic_name", "grid_spec"],
                       cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert ocn["ext"] == 1 and "one row is added to the south end" in r.stdout
    x = pkg.make_coupler_xgrid(_tiles(atm), [(ocn["lon"], ocn["lat"])], [ocn["omask"]], lnd=_tiles(lnd), interp_order=order)
    n = _compare_lists(d, x["atmxlnd"], "C8_mosaic", "C6_mosaic", order)
    n += _compare_lists(d, x["atmxocn"], "C8_mosaic", "ocean_mosaic", order, ext2=1)
    n += _compare_lists(d, x["lndxocn"], "C6_mosaic", "ocean_mosaic", order, ext2=1)
    assert n > 2000
    # ocean_mask.nc (make_coupler_mosaic.c:2020-2070): the artificial row is not written
    g = netcdf_file(os.path.join(d, "ocean_mask.nc"), "r", mmap=False)
    nx, ny = ocn["nx"], ocn["ny"]
    ox = x["ocn_xarea"].reshape(ny + 1, nx)[1:]
    oa = x["area_ocn"].reshape(ny + 1, nx)[1:]
    assert np.array_equal(np.array(g.variables["areaX"][:]), ox)
    assert np.array_equal(np.array(g.variables["areaO"][:]), oa)
    assert np.array_equal(np.array(g.variables["mask"][:]), ox / oa)
    g.close()
    # land_mask_tile<n>.nc (:2072-2120)
    for t in range(6):
        g = netcdf_file(os.path.join(d, f"land_mask_tile{t + 1}.nc"), "r", mmap=False)
        lx = x["lnd_xarea"].reshape(6, 6, 6)[t]
        la = x["area_lnd"].reshape(6, 6, 6)[t]
        assert np.array_equal(np.array(g.variables["l_area"][:]), lx), t
        assert np.array_equal(np.array(g.variables["area_lnd"][:]), la), t
        assert np.array_equal(np.array(g.variables["mask"][:]), lx / la), t
        g.close()


def test_land_on_the_atmosphere_mosaic(pkg, tmp_path):
    """the usual coupled-model set-up: --land_mosaic is the atmosphere's; an atmosphere cell meets its own land cell only and
    there is no land x ocean grid (make_coupler_mosaic.c:1118-1140, :1474-1486, :2484)"""
    tool = _ref_tool()
    d = str(tmp_path)
    atm = _write_mosaic(pkg, d, 8)
    ocn = _write_ocean(d, 36, 24, -81.0, seed=11, frac_land=0.3, area_frac=True)
    r = subprocess.run([tool, "--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C8_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                        "--ocean_topog", "topog.nc", "--interp_order", "2", "--mosaic_name", "grid_spec"],
                       cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    x = pkg.make_coupler_xgrid(_tiles(atm), [(ocn["lon"], ocn["lat"])], [ocn["omask"]], lnd=None, interp_order=2)
    assert x["lndxocn"]["area"].size == 0
    n = _compare_lists(d, x["atmxlnd"], "C8_mosaic", "C8_mosaic", 2)
    n += _compare_lists(d, x["atmxocn"], "C8_mosaic", "ocean_mosaic", 2, ext2=1)
    assert n > 800
    assert np.array_equal(x["atmxlnd"]["t1"], x["atmxlnd"]["t2"]) and np.array_equal(x["atmxlnd"]["i1"], x["atmxlnd"]["i2"])
    # every atmosphere cell is covered: land share + ocean share of its exchange cells add up to its area (but for the shares
    # below MIN_AREA_FRAC = 1e-4 the tool drops, and the lon-lat straight edges of the clip)
    cover = np.zeros(6 * 64)
    for lst in (x["atmxlnd"], x["atmxocn"]):
        np.add.at(cover, lst["t1"] * 64 + lst["j1"] * 8 + lst["i1"], lst["area"])
    assert np.max(np.abs(cover - x["area_atm"]) / x["area_atm"]) < 2e-3


def test_finer_grids_with_partly_wet_ocean_cells(pkg, tmp_path):
    """C24 atmosphere, C16 land, 120 x 80 ocean with area_frac: ~10^5 exchange cells, atmosphere cells with dozens of ocean
    cells under them (the sequential land-share sums are long), land cells with many atmosphere parents"""
    tool = _ref_tool()
    d = str(tmp_path)
    atm = _write_mosaic(pkg, d, 24)
    lnd = _write_mosaic(pkg, d, 16)
    ocn = _write_ocean(d, 120, 80, -79.5, seed=5, frac_land=0.35, area_frac=True)
    r = subprocess.run([tool, "--atmos_mosaic", "C24_mosaic.nc", "--land_mosaic", "C16_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                        "--ocean_topog", "topog.nc", "--interp_order", "2", "--mosaic_name", "grid_spec"],
                       cwd=d, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    x = pkg.make_coupler_xgrid(_tiles(atm), [(ocn["lon"], ocn["lat"])], [ocn["omask"]], lnd=_tiles(lnd), interp_order=2)
    n = _compare_lists(d, x["atmxlnd"], "C24_mosaic", "C16_mosaic", 2)
    n += _compare_lists(d, x["atmxocn"], "C24_mosaic", "ocean_mosaic", 2, ext2=1)
    n += _compare_lists(d, x["lndxocn"], "C16_mosaic", "ocean_mosaic", 2, ext2=1)
    assert n > 30000, n
    g = netcdf_file(os.path.join(d, "ocean_mask.nc"), "r", mmap=False)
    assert np.array_equal(np.array(g.variables["areaX"][:]), x["ocn_xarea"].reshape(81, 120)[1:])
    g.close()


def test_c48_one_degree_trio(pkg, tmp_path):
    """the size the coupled models run at: C48 atmosphere, a land mosaic of its own (C32), one-degree 360 x 200 ocean with
    partly wet cells, order 2: ~240 000 exchange cells in three families, every file of the reference tool bit for bit"""
    tool = _ref_tool()
    d = str(tmp_path)
    atm = _write_mosaic(pkg, d, 48)
    lnd = _write_mosaic(pkg, d, 32)
    ocn = _write_ocean(d, 360, 200, -80.0, seed=1, frac_land=0.3, area_frac=True)
    r = subprocess.run([tool, "--atmos_mosaic", "C48_mosaic.nc", "--land_mosaic", "C32_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                        "--ocean_topog", "topog.nc", "--interp_order", "2", "--mosaic_name", "grid_spec"],
                       cwd=d, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    x = pkg.make_coupler_xgrid(_tiles(atm), [(ocn["lon"], ocn["lat"])], [ocn["omask"]], lnd=_tiles(lnd), interp_order=2)
    n = _compare_lists(d, x["atmxlnd"], "C48_mosaic", "C32_mosaic", 2)
    n += _compare_lists(d, x["atmxocn"], "C48_mosaic", "ocean_mosaic", 2, ext2=1)
    n += _compare_lists(d, x["lndxocn"], "C32_mosaic", "ocean_mosaic", 2, ext2=1)
    assert n > 200000, n
    g = netcdf_file(os.path.join(d, "ocean_mask.nc"), "r", mmap=False)
    assert np.array_equal(np.array(g.variables["areaX"][:]), x["ocn_xarea"].reshape(201, 360)[1:])
    g.close()
    for t in range(6):
        g = netcdf_file(os.path.join(d, f"land_mask_tile{t + 1}.nc"), "r", mmap=False)
        assert np.array_equal(np.array(g.variables["l_area"][:]), x["lnd_xarea"].reshape(6, 32, 32)[t]), t
        g.close()


def test_all_three_models_on_one_mosaic(pkg, tmp_path):
    """aquaplanet-style set-up: atmosphere, land and ocean on the same cubed-sphere mosaic, a six-tile topography file with
    area_frac.  The tool then visits ocean tile n for atmosphere tile n only (make_coupler_mosaic.c:1332-1345), names its files
    atm_..Xlnd_.. / atm_..Xocn_.. (:2151, :2278) and adds no southern row (:840, one-tile oceans only)"""
    tool = _ref_tool()
    d = str(tmp_path)
    atm = _write_mosaic(pkg, d, 8)
    rng = np.random.default_rng(21)
    g = netcdf_file(os.path.join(d, "topog.nc"), "w", version=2)
    g.createDimension("ntiles", 6)
    omask = []
    for t in range(6):
        g.createDimension(f"nx_tile{t + 1}", 8); g.createDimension(f"ny_tile{t + 1}", 8)
    for t in range(6):
        u = rng.uniform(size=(8, 8))
        frac = np.where(u < 0.3, 0.0, np.where(u < 0.6, 1.0, rng.uniform(0.0, 1.0, (8, 8))))
        v = g.createVariable(f"depth_tile{t + 1}", "d", (f"ny_tile{t + 1}", f"nx_tile{t + 1}")); v[:] = 100.0 * frac
        w = g.createVariable(f"area_frac_tile{t + 1}", "d", (f"ny_tile{t + 1}", f"nx_tile{t + 1}")); w[:] = frac
        omask.append(frac)
    g.close()
    r = subprocess.run([tool, "--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C8_mosaic.nc", "--ocean_mosaic", "C8_mosaic.nc",
                        "--ocean_topog", "topog.nc", "--interp_order", "2", "--mosaic_name", "grid_spec"],
                       cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    x = pkg.make_coupler_xgrid(_tiles(atm), _tiles(atm), omask, lnd=None, interp_order=2, ocn_same_as_atm=True)
    n = _compare_lists(d, x["atmxlnd"], "atm_C8_mosaic", "lnd_C8_mosaic", 2)
    n += _compare_lists(d, x["atmxocn"], "atm_C8_mosaic", "ocn_C8_mosaic", 2)
    assert n > 400 and x["lndxocn"]["area"].size == 0
    assert np.array_equal(x["atmxocn"]["t1"], x["atmxocn"]["t2"])
    for t in range(6):
        g = netcdf_file(os.path.join(d, f"ocean_mask_tile{t + 1}.nc"), "r", mmap=False)
        assert np.array_equal(np.array(g.variables["areaX"][:]), x["ocn_xarea"].reshape(6, 8, 8)[t]), t
        g.close()


def _same_files(da, db, inputs):
    """every output file of directory da exists in db with the same dimensions, variables, attributes and values; global
    attributes only by name (they carry version, host, time and argv[0])"""
    fa = sorted(f for f in os.listdir(da) if f.endswith(".nc") and f not in inputs)
    fb = sorted(f for f in os.listdir(db) if f.endswith(".nc") and f not in inputs)
    assert fa == fb, (set(fa) ^ set(fb))
    for name in fa:
        ga = netcdf_file(os.path.join(da, name), "r", mmap=False); gb = netcdf_file(os.path.join(db, name), "r", mmap=False)
        assert dict(ga.dimensions) == dict(gb.dimensions), name
        assert list(ga.variables) == list(gb.variables), (name, list(ga.variables), list(gb.variables))
        assert set(ga._attributes) - {"great_circle_algorithm"} == set(gb._attributes), (name, set(ga._attributes) ^ set(gb._attributes))
        assert ga._attributes["grid_version"] == gb._attributes["grid_version"]
        for v, va in ga.variables.items():
            vb = gb.variables[v]
            assert va.dimensions == vb.dimensions and va.typecode() == vb.typecode(), (name, v)
            assert va._attributes == vb._attributes, (name, v, va._attributes, vb._attributes)
            xa, xb = np.array(va[:]), np.array(vb[:])
            if name.startswith("land_mask") and v == "area_atm":
                continue          # the reference indexes the atmosphere tile with the land cell number (make_coupler_mosaic.c:2083)
            assert np.array_equal(xa, xb), (name, v)
        ga.close(); gb.close()
    return len(fa)


@pytest.mark.parametrize("case", ["own_land_order2", "atm_land_order1", "one_mosaic"])
def test_make_coupler_mosaic_b200_against_the_reference_tool(pkg, tmp_path, case):
    """the command line: make_coupler_mosaic_b200 and the unmodified reference make_coupler_mosaic with the same arguments on
    the same files -> the same files (exchange grids, land / ocean masks, the coupler mosaic file with its file lists)"""
    tool = _ref_tool()
    exe = os.path.join(os.path.dirname(pkg.__file__), "bin", "make_coupler_mosaic_b200")
    assert os.path.exists(exe), "make_coupler_mosaic_b200 is not built (python __graft_entry__.py)"
    dirs = [str(tmp_path / "ref"), str(tmp_path / "b200")]
    for d in dirs:
        os.makedirs(d)
        _write_mosaic(pkg, d, 8)
        if case == "own_land_order2":
            _write_mosaic(pkg, d, 6)
        if case == "one_mosaic":
            rng = np.random.default_rng(33)
            g = netcdf_file(os.path.join(d, "topog.nc"), "w", version=2)
            g.createDimension("ntiles", 6)
            for t in range(6):
                g.createDimension(f"nx_tile{t + 1}", 8); g.createDimension(f"ny_tile{t + 1}", 8)
            for t in range(6):
                v = g.createVariable(f"depth_tile{t + 1}", "d", (f"ny_tile{t + 1}", f"nx_tile{t + 1}"))
                v[:] = np.where(rng.uniform(size=(8, 8)) < 0.4, 0.0, 50.0)
            g.close()
        else:
            _write_ocean(d, 30, 20, -78.0, seed=3, area_frac=(case == "atm_land_order1"))
    args = {"own_land_order2": ["--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C6_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                                "--ocean_topog", "topog.nc", "--interp_order", "2", "--mosaic_name", "grid_spec"],
            "atm_land_order1": ["--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C8_mosaic.nc", "--ocean_mosaic", "ocean_mosaic.nc",
                                "--ocean_topog", "topog.nc", "--interp_order", "1"],
            "one_mosaic": ["--atmos_mosaic", "C8_mosaic.nc", "--land_mosaic", "C8_mosaic.nc", "--ocean_mosaic", "C8_mosaic.nc",
                           "--ocean_topog", "topog.nc", "--sea_level", "10"]}[case]
    inputs = set(os.listdir(dirs[0]))
    for cmd, d in ((tool, dirs[0]), (exe, dirs[1])):
        r = subprocess.run([cmd] + args, cwd=d, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (cmd, r.stdout[-1500:], r.stderr[-1500:])
    n = _same_files(dirs[0], dirs[1], inputs)
    assert n >= (14 if case != "own_land_order2" else 30), n


def test_make_coupler_mosaic_b200_messages(pkg, tmp_path):
    exe = os.path.join(os.path.dirname(pkg.__file__), "bin", "make_coupler_mosaic_b200")
    d = str(tmp_path)
    _write_mosaic(pkg, d, 8)
    r = subprocess.run([exe, "--atmos_mosaic", "C8_mosaic.nc"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "ocean_mosaic is not specified" in r.stderr
    r = subprocess.run([exe, "--atmos_mosaic", "C8_mosaic.nc", "--ocean_mosaic", "C8_mosaic.nc", "--ocean_topog", "t.nc", "--wave_mosaic", "w.nc"],
                       cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "--wave_mosaic is not built" in r.stderr


def test_bad_arguments_are_refused(pkg):
    lon, lat = np.meshgrid(np.linspace(0, 1, 3), np.linspace(0, 1, 3))
    with pytest.raises(pkg.XgridError):
        pkg.make_coupler_xgrid([(lon, lat)], [(lon, lat)], [np.ones((2, 2))], lnd=None, interp_order=3)
