"""csrc/h5r.c + the netCDF-4 layer of csrc/nc3.c: netCDF-4 (HDF5) files read through the same nc3_* calls as classic files.

The reference reads every file through libnetcdf and writes NC_FORMAT_NETCDF4_CLASSIC by default (mpp_io.c:52, :109-140,
:163-169), so stock mosaics, supergrids, field files and remap files are HDF5.  No HDF5 library exists in this image and
the reference tree has no sample file: the files here come from tests/h5_writer.py, an independent writer of the same
specification in the two layouts libnetcdf produces (and a plain HDF5 1.6-style one) — parity of the READER with real
libhdf5 output is therefore unpinned, and DESIGN.md says so."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

import h5_writer as W

STYLES = ("v18", "earliest", "plain")


@pytest.fixture(scope="module")
def L(pkg):
    # XGB_NC3_TEST_LIB: a sanitizer build of nc3.c + h5r.c alone (developer runs)
    L = C.CDLL(os.environ.get("XGB_NC3_TEST_LIB") or os.path.join(os.path.dirname(pkg.__file__), "libxgrid_b200.so"))
    vp = C.c_void_p
    L.nc3_open.restype = vp; L.nc3_open.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.nc3_dim_len.restype = C.c_longlong; L.nc3_dim_len.argtypes = [vp, C.c_int]
    L.nc3_dim_name.restype = C.c_char_p; L.nc3_dim_name.argtypes = [vp, C.c_int]
    L.nc3_var_name.restype = C.c_char_p; L.nc3_var_name.argtypes = [vp, C.c_int]
    L.nc3_att_name.restype = C.c_char_p; L.nc3_att_name.argtypes = [vp, C.c_int, C.c_int]
    L.nc3_strerror.restype = C.c_char_p; L.nc3_strerror.argtypes = [vp]
    L.nc3_var_dimids.restype = C.POINTER(C.c_int); L.nc3_var_dimids.argtypes = [vp, C.c_int]
    for n in ("nc3_ndims", "nc3_nvars", "nc3_unlimdim", "nc3_close", "nc3_format"):
        getattr(L, n).argtypes = [vp]
    for n in ("nc3_var_type", "nc3_var_ndims", "nc3_var_natts"):
        getattr(L, n).argtypes = [vp, C.c_int]
    for n in ("nc3_get_var_double", "nc3_get_var_int"):
        getattr(L, n).argtypes = [vp, C.c_int, vp]
    sz = C.POINTER(C.c_size_t)
    for n in ("nc3_get_vara_double", "nc3_get_vara_int", "nc3_get_vara_text"):
        getattr(L, n).argtypes = [vp, C.c_int, sz, sz, vp]
    L.nc3_get_att_text.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    L.nc3_get_att_double.argtypes = [vp, C.c_int, C.c_char_p, vp, C.c_int]
    L.nc3_att_inq.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
    L.nc3_var_id.argtypes = [vp, C.c_char_p]; L.nc3_dim_id.argtypes = [vp, C.c_char_p]
    return L


def _open(L, path):
    err = C.create_string_buffer(512)
    f = L.nc3_open(str(path).encode(), err, 512)
    return f, err.value


def _sz(*v):
    return (C.c_size_t * len(v))(*v)


def _describe(L, f):
    """Everything nc3 exposes of an open file: dimensions, variables (type, dimensions, attributes, values), global attributes."""
    def atts(v):
        out = {}
        for k in range(L.nc3_var_natts(f, v)):
            name = L.nc3_att_name(f, v, k)
            t, n = C.c_int(), C.c_longlong()
            assert L.nc3_att_inq(f, v, name, C.byref(t), C.byref(n)) == 0
            if t.value == 2:
                buf = C.create_string_buffer(int(n.value) + 1)
                assert L.nc3_get_att_text(f, v, name, buf, int(n.value) + 1) == 0
                out[name] = (2, buf.raw[:n.value])
            else:
                d = np.zeros(max(int(n.value), 1))
                assert L.nc3_get_att_double(f, v, name, d.ctypes.data, int(n.value)) >= 0
                out[name] = (t.value, d[:n.value].tolist())
        return out
    dims = [(L.nc3_dim_name(f, i), L.nc3_dim_len(f, i)) for i in range(L.nc3_ndims(f))]
    res = {"dims": dims, "unlim": L.nc3_unlimdim(f), "gatts": atts(-1), "vars": {}}
    for v in range(L.nc3_nvars(f)):
        nd = L.nc3_var_ndims(f, v)
        di = [L.nc3_var_dimids(f, v)[k] for k in range(nd)]
        shape = tuple(dims[k][1] for k in di)
        t = L.nc3_var_type(f, v)
        n = int(np.prod(shape)) if shape else 1
        if t == 2:
            buf = C.create_string_buffer(max(n, 1))
            assert L.nc3_get_vara_text(f, v, _sz(*([0] * nd)), _sz(*shape), buf) == 0, L.nc3_strerror(f)
            val = buf.raw[:n]
        else:
            a = np.zeros(shape)
            assert L.nc3_get_var_double(f, v, a.ctypes.data) == 0, L.nc3_strerror(f)
            val = a
        res["vars"][L.nc3_var_name(f, v)] = (t, di, atts(v), val)
    return res


def _sample(style, rng):
    x = rng.standard_normal((7, 13))
    t = np.arange(5.0)
    fld = rng.standard_normal((5, 7, 13)).astype("f4")
    name = np.frombuffer(b"hello".ljust(16, b"\0"), "S1")
    cnt = np.arange(13, dtype="i4")
    atts = {"missing_value": np.float32(1e20), "_FillValue": np.float32(1e20), "long_name": "f", "a1": 1.0, "a2": np.int32(2),
            "a3": np.int16(3), "a4": "x", "a5": np.arange(3.0), "a6": "", "a7": "seven", "a8": np.int8(-4)}
    vs = [W.Var("time", ("time",), t, {"units": "days", "cartesian_axis": "T"}),
          W.Var("x", ("ny", "nx"), x, {"standard_name": "geographic_longitude", "units": "degree_east"}, chunks=(3, 5), deflate=4,
                shuffle=True, fletcher=True),
          W.Var("fld", ("time", "ny", "nx"), fld, atts, fill=np.float32(1e20)),
          W.Var("name", ("string",), name),
          W.Var("nx", ("nx",), cnt, big_endian=True),
          W.Var("ny", ("time", "ny"), np.arange(35, dtype="i2").reshape(5, 7)),     # shares a dimension's name: _nc4_non_coord_
          W.Var("scalar", (), np.float64(3.5), layout="compact"),
          W.Var("u1", ("nx",), np.arange(200, 213, dtype="u1")), W.Var("i8", ("nx",), np.arange(13, dtype="i8") * 2 ** 33)]
    extra = [W.Var("v%d" % i, ("nx",), np.arange(13.0) + i) for i in range(40 if style != "plain" else 3)]
    dims = {"time": None, "ny": 7, "nx": 13, "string": 16}
    return dims, vs + extra, {"grid_version": "0.2", "n": np.int32(4)}, dict(x=x, t=t, fld=fld, cnt=cnt)


@pytest.mark.parametrize("style", STYLES)
def test_every_structure_the_two_libnetcdf_layouts_use(L, tmp_path, style):
    """Dense links and attributes (fractal heap with a root indirect block), continuation blocks, chunk B-trees of two levels
    with deflate + shuffle + fletcher32, an unlimited dimension, a big-endian variable, compact and contiguous data."""
    rng = np.random.default_rng(1)
    dims, vs, gatts, ref = _sample(style, rng)
    p = tmp_path / f"t_{style}.nc"
    W.write_netcdf4(str(p), dims, vs, gatts, style=style, leaf=3)
    f, err = _open(L, p)
    assert f, err
    assert L.nc3_format(f) == 4
    got = _describe(L, f)
    if style == "plain":
        assert all(n.startswith(b"phony_dim_") for n, _ in got["dims"]) and got["unlim"] == -1
    else:
        assert got["dims"] == [(b"time", 5), (b"ny", 7), (b"nx", 13), (b"string", 16)] and got["unlim"] == 0
        assert list(got["vars"])[:9] == [b"time", b"x", b"fld", b"name", b"nx", b"ny", b"scalar", b"u1", b"i8"]   # definition order
        assert got["vars"][b"fld"][1] == [0, 1, 2] and got["vars"][b"x"][1] == [1, 2] and got["vars"][b"ny"][1] == [0, 1]
    assert got["gatts"] == {b"grid_version": (2, b"0.2"), b"n": (4, [4.0])}
    V = got["vars"]
    assert len(V) == len(vs)
    assert np.array_equal(V[b"x"][3], ref["x"]) and V[b"x"][0] == 6
    assert np.array_equal(V[b"fld"][3], ref["fld"].astype("f8")) and V[b"fld"][0] == 5
    assert np.array_equal(V[b"time"][3], ref["t"]) and np.array_equal(V[b"nx"][3], ref["cnt"]) and V[b"nx"][0] == 4
    assert np.array_equal(V[b"ny"][3], np.arange(35.0).reshape(5, 7)) and V[b"ny"][0] == 3
    assert V[b"scalar"][3] == 3.5 and V[b"scalar"][1] == []
    assert np.array_equal(V[b"u1"][3], np.arange(200.0, 213.0)) and np.array_equal(V[b"i8"][3], np.arange(13.0) * 2.0 ** 33)
    assert V[b"name"][3].rstrip(b"\0") == b"hello" and V[b"name"][0] == 2
    assert np.array_equal(V[b"v2"][3], np.arange(13.0) + 2)
    a = V[b"fld"][2]
    assert a[b"missing_value"] == (5, [float(np.float32(1e20))]) and a[b"a1"] == (6, [1.0]) and a[b"a2"] == (4, [2.0])
    assert a[b"a3"] == (3, [3.0]) and a[b"a4"] == (2, b"x") and a[b"a5"] == (6, [0.0, 1.0, 2.0]) and a[b"a7"] == (2, b"seven")
    assert a[b"a8"] == (1, [-4.0]) and a[b"a6"][0] == 2 and a[b"a6"][1].rstrip(b"\0") == b""
    assert list(a) == [b"missing_value", b"_FillValue", b"long_name", b"a1", b"a2", b"a3", b"a4", b"a5", b"a6", b"a7", b"a8"]
    assert V[b"time"][2] == {b"units": (2, b"days"), b"cartesian_axis": (2, b"T")}      # CLASS / NAME / _Netcdf4Dimid are hidden
    # hyperslabs across chunk boundaries, integer access, and the errors
    out = np.zeros((2, 3, 4)); v = L.nc3_var_id(f, b"fld")
    assert L.nc3_get_vara_double(f, v, _sz(1, 2, 3), _sz(2, 3, 4), out.ctypes.data) == 0 and np.array_equal(out, ref["fld"][1:3, 2:5, 3:7])
    out = np.zeros((4, 6)); v = L.nc3_var_id(f, b"x")
    assert L.nc3_get_vara_double(f, v, _sz(2, 4), _sz(4, 6), out.ctypes.data) == 0 and np.array_equal(out, ref["x"][2:6, 4:10])
    io = np.zeros(5, "i4"); v = L.nc3_var_id(f, b"nx")
    assert L.nc3_get_vara_int(f, v, _sz(8), _sz(5), io.ctypes.data) == 0 and np.array_equal(io, ref["cnt"][8:])
    assert L.nc3_get_vara_double(f, L.nc3_var_id(f, b"x"), _sz(5, 10), _sz(3, 4), out.ctypes.data) != 0 and b"exceeds" in L.nc3_strerror(f)
    assert L.nc3_get_vara_double(f, L.nc3_var_id(f, b"name"), _sz(0), _sz(1), out.ctypes.data) != 0
    L.nc3_close(f)


@pytest.mark.parametrize("style", ("v18", "earliest"))
def test_a_classic_file_and_its_netcdf4_twin_read_the_same(L, tmp_path, style):
    """A field file like the CLI tests' (record variables, float data with missing values, coordinate variables) written by
    scipy, re-expressed as netCDF-4 (chunked + deflated as FRE history files are): nc3 describes both identically."""
    rng = np.random.default_rng(3)
    p3 = str(tmp_path / "c.nc")
    g = netcdf_file(p3, "w", version=2)
    g.createDimension("time", None); g.createDimension("pfull", 3); g.createDimension("grid_yt", 12); g.createDimension("grid_xt", 20)
    g.createDimension("string", 255)
    g.title = "synthetic"
    v = g.createVariable("time", "d", ("time",)); v.units = "days since 2000-01-01"; v.cartesian_axis = "T"
    pf = g.createVariable("pfull", "d", ("pfull",)); pf.units = "mb"; pf[:] = [100.0, 500.0, 900.0]
    a = g.createVariable("temp", "f", ("time", "pfull", "grid_yt", "grid_xt")); a.missing_value = np.float32(1e20); a.units = "K"
    b = g.createVariable("ps", "d", ("time", "grid_yt", "grid_xt")); b.units = "Pa"
    c = g.createVariable("orog", "d", ("grid_yt", "grid_xt"))
    c[:] = rng.uniform(0, 3000, (12, 20))
    s = g.createVariable("gridfiles", "c", ("pfull", "string"))
    for k in range(3):
        s[k] = np.frombuffer(b"C48_grid.tile%d.nc" % (k + 1) + b"\0" * (255 - 17), "S1")
    for k in range(4):
        v[k] = 10.0 + k; a[k] = rng.standard_normal((3, 12, 20)).astype("f4"); b[k] = rng.standard_normal((12, 20))
    g.close()
    p4 = str(tmp_path / "h.nc")
    W.from_classic(p3, p4, style=style, chunk=7, deflate=2, shuffle=True)
    assert open(p4, "rb").read(8) == b"\x89HDF\r\n\x1a\n"
    f3, e3 = _open(L, p3); f4, e4 = _open(L, p4)
    assert f3 and f4, (e3, e4)
    d3, d4 = _describe(L, f3), _describe(L, f4)
    assert d3["dims"] == d4["dims"] and d3["unlim"] == d4["unlim"] and d3["gatts"] == d4["gatts"]
    assert list(d3["vars"]) == list(d4["vars"])
    for name in d3["vars"]:
        t3, di3, a3, v3 = d3["vars"][name]; t4, di4, a4, v4 = d4["vars"][name]
        assert (t3, di3, a3) == (t4, di4, a4), name
        assert np.array_equal(v3, v4) if isinstance(v3, np.ndarray) else v3 == v4, name
    L.nc3_close(f3); L.nc3_close(f4)


def test_unwritten_chunks_read_as_the_fill_value_and_a_user_block_is_skipped(L, tmp_path):
    x = np.arange(48.0).reshape(6, 8)
    v = W.Var("x", ("ny", "nx"), x, chunks=(4, 3), fill=-7.5)
    v.skip_chunks = {(0, 1), (1, 2)}
    p = tmp_path / "u.nc"
    W.write_netcdf4(str(p), {"ny": 6, "nx": 8}, [v], style="v18", user_block=1024)
    f, err = _open(L, p)
    assert f, err
    out = np.zeros((6, 8))
    assert L.nc3_get_var_double(f, L.nc3_var_id(f, b"x"), out.ctypes.data) == 0, L.nc3_strerror(f)
    want = x.copy(); want[0:4, 3:6] = -7.5; want[4:6, 6:8] = -7.5
    assert np.array_equal(out, want)
    L.nc3_close(f)


def test_damaged_files_give_errors_not_crashes(L, tmp_path):
    """Every structure is bounds-checked: truncations and byte flips either still open (damage in data) or fail with a message."""
    rng = np.random.default_rng(5)
    dims, vs, gatts, _ = _sample("v18", rng)
    good = tmp_path / "g.nc"
    W.write_netcdf4(str(good), dims, vs, gatts, style="v18", leaf=3)
    raw = open(good, "rb").read()
    bad = tmp_path / "b.nc"
    refused = 0
    cases = [raw[:n] for n in (9, 40, 60, 200, 1000, len(raw) // 2, len(raw) - 40)]
    for _ in range(150):
        b = bytearray(raw)
        for _k in range(int(rng.integers(1, 6))):
            b[int(rng.integers(8, len(b)))] = int(rng.integers(0, 256))
        cases.append(bytes(b))
    for data in cases:
        open(bad, "wb").write(data)
        f, err = _open(L, bad)
        if not f:
            refused += 1
            assert err
            continue
        for v in range(L.nc3_nvars(f)):
            nd = L.nc3_var_ndims(f, v)
            n = 1
            for k in range(nd):
                n *= max(L.nc3_dim_len(f, L.nc3_var_dimids(f, v)[k]), 0)
            if n > 10 ** 6 or L.nc3_var_type(f, v) == 2:
                continue
            a = np.zeros(max(n, 1))
            L.nc3_get_var_double(f, v, a.ctypes.data)
        L.nc3_close(f)
    assert refused >= 7
    open(bad, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 100)
    f, err = _open(L, bad)
    assert not f and b"h5r" in err
