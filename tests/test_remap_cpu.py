"""Remap files (include/xgrid_b200.h Part 4; csrc/remap_file.c over csrc/nc3.c) against the UNMODIFIED reference's WRITE and
READ branches of setup_conserve_interp (conserve_interp.c:62-125, :368-443), whose netCDF calls are recorded by the in-memory
store of oracle/shim/io_stubs.c, and against scipy.io.netcdf_file as an independent implementation of the classic format.
No GPU: the lists come from the compiled reference."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

import xgtest

ORDER = {1: xgtest.ORDER1, 2: xgtest.ORDER2}


def _grids(pkg, ni=12, nlon=40, nlat=20):
    lonc, latc = pkg.cubed_sphere_grid(ni)
    lon2, lat2 = pkg.latlon_grid(nlon, nlat)
    return lonc, latc, lon2, lat2


def _write(pkg, path, order, x, isc=0, jsc=0):
    L = pkg.lib()
    p = lambda a: a.ctypes.data
    rc = L.xgb_remap_write(path.encode(), order, x["nxgrid"], p(x["t_in"]), p(x["i_in"]), p(x["j_in"]), p(x["i_out"]), p(x["j_out"]),
                           isc, jsc, p(x["area"]), p(x["di"]) if order == 2 else None, p(x["dj"]) if order == 2 else None)
    assert rc == 0, L.xgb_last_error()


def _read(pkg, path, order):
    L = pkg.lib()
    n = L.xgb_remap_size(path.encode())
    assert n >= 0, L.xgb_last_error()
    out = {k: np.zeros(n, np.int32) for k in ("t_in", "i_in", "j_in", "i_out", "j_out")}
    out["area"] = np.zeros(n)
    if order == 2:
        out["di"] = np.zeros(n); out["dj"] = np.zeros(n)
    p = lambda k: out[k].ctypes.data
    rc = L.xgb_remap_read(path.encode(), order, n, p("t_in"), p("i_in"), p("j_in"), p("i_out"), p("j_out"), p("area"),
                          p("di") if order == 2 else None, p("dj") if order == 2 else None)
    assert rc == 0, L.xgb_last_error()
    out["nxgrid"] = n
    return out


def _classic_bytes(version, dims, variables):
    """the netCDF classic format specification, written out independently of csrc/nc3.c: header (magic, numrecs = 0, dim list,
    absent global attributes, var list with text attributes) followed by the fixed-size variables in definition order, as
    libnetcdf lays out a file without record variables (scipy's writer cannot serve here: it reorders variables by shape)"""
    import struct
    def name(s_):
        b = s_.encode()
        return struct.pack(">i", len(b)) + b + bytes(-len(b) % 4)
    ts = {4: 4, 6: 8}
    def header(begins):
        h = b"CDF" + bytes([version]) + struct.pack(">i", 0)
        h += struct.pack(">ii", 0x0A, len(dims)) + b"".join(name(n) + struct.pack(">i", sz) for n, sz in dims)
        h += struct.pack(">ii", 0, 0)
        h += struct.pack(">ii", 0x0B, len(variables))
        for (vn, t, dd, atts, data), beg in zip(variables, begins):
            h += name(vn) + struct.pack(">i", len(dd)) + b"".join(struct.pack(">i", k) for k in dd)
            h += struct.pack(">ii", 0x0C, len(atts)) if atts else struct.pack(">ii", 0, 0)
            for an, av in atts:
                h += name(an) + struct.pack(">ii", 2, len(av)) + av.encode() + bytes(-len(av) % 4)
            vsize = data.size * ts[t]
            h += struct.pack(">ii", t, vsize + (-vsize % 4)) + struct.pack(">i" if version == 1 else ">q", beg)
        return h
    n0 = len(header([0] * len(variables)))
    begins, at = [], n0
    for vn, t, dd, atts, data in variables:
        begins.append(at); at += data.size * ts[t]
    body = b"".join(np.ascontiguousarray(data).astype(">i4" if t == 4 else ">f8").tobytes() for vn, t, dd, atts, data in variables)
    return header(begins) + body


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("fmt", ["classic", "64bit_offset"])
def test_written_file_equals_what_the_reference_writes(pkg, reflib, tmp_path, order, fmt):
    """the reference's WRITE branch, recorded call by call and rendered per the classic-format specification == the product's
    file, byte for byte (and field for field through scipy's reader); a destination row window (jsc > 0) exercises the index
    shift of tile2_cell"""
    lonc, latc, lon2, lat2 = _grids(pkg)
    for jsc, jec in ((None, None), (5, 13)):
        ref = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[order], jsc=jsc, jec=jec, remap=("remap_ref.nc", 1))
        dims, variables = xgtest.ref_store_file("remap_ref.nc")
        assert [d[0] for d in dims] == ["string", "ncells", "two"] and dims[0][1] == 255 and dims[1][1] == ref["nxgrid"]
        names = [v[0] for v in variables]
        assert names == ["tile1", "tile1_cell", "tile2_cell", "xgrid_area"] + (["tile1_distance"] if order == 2 else [])
        got = str(tmp_path / "got.nc")
        assert pkg.lib().xgb_set_nc_format(fmt.encode()) == 0
        _write(pkg, got, order, ref, isc=0, jsc=jsc or 0)
        a, b = _classic_bytes(1 if fmt == "classic" else 2, dims, variables), open(got, "rb").read()
        assert a == b, (len(a), len(b))
        # and, through an independent reader, name for name and value for value in the reference's order
        g = netcdf_file(got, "r", mmap=False)
        assert list(g.dimensions.items()) == dims and list(g.variables) == names
        for vn, t, dd, atts, data in variables:
            v = g.variables[vn]
            assert v.typecode() == xgtest.NC_TYPES[t] and v.dimensions == tuple(dims[k][0] for k in dd)
            assert {k: val.decode() for k, val in v._attributes.items()} == dict(atts)
            assert np.array_equal(v[:], data)
        g.close()
    pkg.lib().xgb_set_nc_format(b"64bit_offset")


@pytest.mark.parametrize("order", [1, 2])
def test_read_equals_the_reference_read_branch(pkg, reflib, tmp_path, order):
    """a file the product wrote, loaded into the reference's store with scipy, read by the reference's READ branch ==
    xgb_remap_read (area rescale (a / 4 pi R^2) * 4 pi R^2 included), bit for bit"""
    lonc, latc, lon2, lat2 = _grids(pkg)
    ref = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[order])
    path = str(tmp_path / "r.nc")
    _write(pkg, path, order, ref)
    xgtest.ref_store_load("remap_in.nc", path)
    back = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[order], remap=("remap_in.nc", 2))
    mine = _read(pkg, path, order)
    assert back["nxgrid"] == mine["nxgrid"] == ref["nxgrid"]
    for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
        assert np.array_equal(back[k], mine[k]) and np.array_equal(mine[k], ref[k]), k
    for k in ("area",) + (("di", "dj") if order == 2 else ()):
        assert np.array_equal(back[k].view(np.uint64), mine[k].view(np.uint64)), k
    assert np.allclose(mine["area"], ref["area"], rtol=4e-16, atol=0)
    if order == 2:
        assert np.array_equal(mine["di"], ref["di"]) and np.array_equal(mine["dj"], ref["dj"])


@pytest.mark.parametrize("order", [1, 2])
def test_setup_conserve_interp_READ_through_the_reference_structs(pkg, reflib, tmp_path, order):
    """the product's reference-signature setup_conserve_interp with READ set and a destination row window, called from the
    reference's driver code on its own Grid_config / Interp_config, == the reference's READ branch on the same file"""
    lonc, latc, lon2, lat2 = _grids(pkg)
    full = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[order])
    path = str(tmp_path / "w.nc")
    _write(pkg, path, order, full)
    xgtest.ref_store_load("remap_win.nc", path)
    L = pkg.lib(); R = reflib
    setup_fn = C.cast(L.setup_conserve_interp, C.c_void_p)
    for jsc, jec in ((0, 19), (4, 11)):
        want = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[order], jsc=jsc, jec=jec, remap=("remap_win.nc", 2), keep=True)
        h = R.ref_regrid_setup_through_remap(want["handle"], setup_fn, path.encode(), 2)
        n = R.ref_regrid_nxgrid(h)
        assert n == want["nxgrid"] and n > 0
        got = xgtest._alloc(n, order)
        R.ref_regrid_get(h, got["t_in"], got["i_in"], got["j_in"], got["i_out"], got["j_out"], got["area"],
                         got["di"].ctypes.data if order == 2 else None, got["dj"].ctypes.data if order == 2 else None)
        for k in got:
            assert np.array_equal(got[k].view(np.uint32 if got[k].dtype == np.int32 else np.uint64),
                                  want[k].view(np.uint32 if got[k].dtype == np.int32 else np.uint64)), (k, jsc)
        assert got["j_out"].min() == 0 or jsc == 0          # indices are relative to the window


def test_scipy_reads_what_the_product_writes_and_errors_are_loud(pkg, reflib, tmp_path):
    lonc, latc, lon2, lat2 = _grids(pkg)
    ref = xgtest.ref_setup(lonc, latc, lon2, lat2, ORDER[2])
    L = pkg.lib()
    for fmt, vb in (("classic", 1), ("64bit_offset", 2)):
        assert L.xgb_set_nc_format(fmt.encode()) == 0
        path = str(tmp_path / f"{fmt}.nc")
        _write(pkg, path, 2, ref)
        g = netcdf_file(path, "r", mmap=False)
        assert g.version_byte == vb
        assert np.array_equal(g.variables["tile1"][:], ref["t_in"] + 1)
        assert np.array_equal(g.variables["tile1_cell"][:, 0], ref["i_in"] + 1) and np.array_equal(g.variables["tile1_cell"][:, 1], ref["j_in"] + 1)
        assert np.array_equal(g.variables["tile2_cell"][:, 0], ref["i_out"] + 1) and np.array_equal(g.variables["tile2_cell"][:, 1], ref["j_out"] + 1)
        assert np.array_equal(g.variables["xgrid_area"][:], ref["area"])
        assert np.array_equal(g.variables["tile1_distance"][:, 0], ref["di"]) and np.array_equal(g.variables["tile1_distance"][:, 1], ref["dj"])
        assert g.variables["xgrid_area"].units == b"m2" and g.variables["xgrid_area"].standard_name == b"exchange_grid_area"
        assert g.variables["tile1_distance"].standard_name == b"distance_from_parent1_cell_centroid"
        g.close()
    assert L.xgb_set_nc_format(b"cdf5") == 0
    path = str(tmp_path / "five.nc")
    _write(pkg, path, 2, ref)
    back = _read(pkg, path, 2)
    assert open(path, "rb").read(4) == b"CDF\x05" and np.array_equal(back["i_out"], ref["i_out"]) and np.array_equal(back["dj"], ref["dj"])
    L.xgb_set_nc_format(b"64bit_offset")
    # netCDF-4 is refused by name on write; on read it goes through csrc/h5r.c (the reference's default format, mpp_io.c:52)
    assert L.xgb_set_nc_format(b"netcdf4") != 0 and b"HDF5" in L.xgb_last_error()
    assert L.xgb_set_nc_format(b"bogus") != 0 and b"not a valid option" in L.xgb_last_error()
    h5 = str(tmp_path / "h5.nc")
    open(h5, "wb").write(b"\x89HDF\r\n\x1a\n" + bytes(64))
    assert L.xgb_remap_size(h5.encode()) < 0 and b"h5r" in L.xgb_last_error()
    import h5_writer
    for style in ("v18", "earliest"):
        for o in (1, 2):
            classic = str(tmp_path / f"c{o}.nc")
            _write(pkg, classic, o, ref)
            h5_writer.from_classic(classic, h5, style=style, chunk=1000, deflate=1, shuffle=True)
            assert open(h5, "rb").read(4) == b"\x89HDF"
            a, b = _read(pkg, classic, o), _read(pkg, h5, o)
            assert sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a)
    assert L.xgb_remap_size(str(tmp_path / "missing.nc").encode()) < 0 and b"cannot open" in L.xgb_last_error()


@pytest.mark.parametrize("order", [1, 2])
def test_golden_remap_files_round_trip(pkg, order):
    """tests/golden/remap_c8_20x10_order*.nc (the reference's WRITE branch, tests/golden/make_remap_golden.py; no reference needed
    here): read with xgb_remap_read, written again with xgb_remap_write -> the same bytes; scipy reads the same values"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"remap_c8_20x10_order{order}.nc")
    gold = open(path, "rb").read()
    L = pkg.lib()
    x = _read(pkg, path, order)
    g = netcdf_file(path, "r", mmap=False)
    assert g.dimensions["ncells"] == x["nxgrid"] and np.array_equal(g.variables["tile1_cell"][:, 0], x["i_in"] + 1)
    area_disk = np.ascontiguousarray(g.variables["xgrid_area"][:], dtype=np.float64)     # native byte order
    g.close()
    assert np.allclose(x["area"], area_disk, rtol=4e-16, atol=0)       # (a / 4 pi R^2) * 4 pi R^2 on read, like the reference
    x["area"] = area_disk
    assert L.xgb_set_nc_format(b"64bit_offset") == 0
    out = path + ".tmp"
    try:
        _write(pkg, out, order, x)
        assert open(out, "rb").read() == gold
    finally:
        if os.path.exists(out):
            os.remove(out)


def test_remap_files_with_another_layout_are_refused(pkg, tmp_path):
    """xgb_remap_read sizes its buffers from the `ncells` dimension; a file whose variables have another shape (three
    columns, a missing `two`, a text variable) or whose header claims an absurd attribute length must be refused, not read
    past the buffers"""
    L = pkg.lib()
    n = 5

    def write(path, mutate):
        g = netcdf_file(path, "w", version=2)
        g.createDimension("string", 255); g.createDimension("ncells", n); g.createDimension("two", 2); g.createDimension("three", 3)
        shapes = {"tile1": ("ncells",), "tile1_cell": ("ncells", "two"), "tile2_cell": ("ncells", "two"), "xgrid_area": ("ncells",)}
        types = {"tile1": "i", "tile1_cell": "i", "tile2_cell": "i", "xgrid_area": "d"}
        mutate(shapes, types)
        for k, dims in shapes.items():
            v = g.createVariable(k, types[k], dims)
            v[:] = np.ones([{"ncells": n, "two": 2, "three": 3, "string": 255}[d] for d in dims], dtype=v.data.dtype) if types[k] != "c" else b"x"
        g.close()

    def read(path):
        out = [np.zeros(n, np.int32) for _ in range(5)] + [np.zeros(n)]
        return L.xgb_remap_read(path.encode(), 1, n, *[a.ctypes.data for a in out], None, None)

    good = str(tmp_path / "good.nc"); write(good, lambda s, t: None)
    assert read(good) == 0, L.xgb_last_error()
    for tag, mut in (("three", lambda s, t: s.update(tile1_cell=("ncells", "three"))),
                     ("flat", lambda s, t: s.update(tile2_cell=("ncells",))),
                     ("pairs", lambda s, t: s.update(xgrid_area=("ncells", "two"))),
                     ("other_dim", lambda s, t: s.update(tile1=("three",))),
                     ("text", lambda s, t: (s.update(tile1=("ncells",)), t.update(tile1="c")))):
        path = str(tmp_path / f"{tag}.nc"); write(path, mut)
        assert read(path) != 0, tag
        assert b"is not a numeric" in L.xgb_last_error(), (tag, L.xgb_last_error())
    # a CDF-5 header whose one global attribute claims 2^60 doubles
    import struct
    hdr = b"CDF\x05" + struct.pack(">q", 0) + struct.pack(">iq", 0, 0) + struct.pack(">iq", 12, 1) + \
        struct.pack(">q", 4) + b"name" + struct.pack(">iq", 6, 1 << 60) + bytes(64)
    bad = str(tmp_path / "huge_att.nc"); open(bad, "wb").write(hdr)
    assert L.xgb_remap_size(bad.encode()) < 0 and b"malformed header" in L.xgb_last_error(), L.xgb_last_error()
