"""csrc/nc3.c (classic netCDF CDF-1 / CDF-2 / CDF-5 reader and writer) against scipy.io.netcdf_file, an independent implementation
of the same on-disk format: files written by either side are read by the other; fixed and record variables, every classic type,
strided hyperslabs ((n, 2) column writes like the reference's remap-file writer, conserve_interp.c:405-437), attributes, a
byte-for-byte comparison where scipy keeps the definition order, and the hand-over of HDF5 files to csrc/h5r.c."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.io import netcdf_file


@pytest.fixture(scope="module")
def L(pkg):
    L = C.CDLL(os.path.join(os.path.dirname(pkg.__file__), "libxgrid_b200.so"))
    vp = C.c_void_p
    L.nc3_open.restype = vp; L.nc3_open.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.nc3_create.restype = vp; L.nc3_create.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    L.nc3_strerror.restype = C.c_char_p; L.nc3_strerror.argtypes = [vp]
    L.nc3_dim_len.restype = C.c_longlong; L.nc3_dim_len.argtypes = [vp, C.c_int]
    for n in ("nc3_dim_id", "nc3_var_id"):
        getattr(L, n).argtypes = [vp, C.c_char_p]
    L.nc3_def_dim.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.nc3_def_var.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.nc3_put_att_text.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p]
    L.nc3_put_att_double.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.c_int, vp]
    L.nc3_enddef.argtypes = [vp]; L.nc3_close.argtypes = [vp]
    sz = C.POINTER(C.c_size_t)
    for n in ("nc3_put_vara_double", "nc3_put_vara_int", "nc3_put_vara_text", "nc3_get_vara_double", "nc3_get_vara_int", "nc3_get_vara_text"):
        getattr(L, n).argtypes = [vp, C.c_int, sz, sz, vp]
    for n in ("nc3_put_var_double", "nc3_put_var_int", "nc3_get_var_double", "nc3_get_var_int"):
        getattr(L, n).argtypes = [vp, C.c_int, vp]
    L.nc3_get_att_text.argtypes = [vp, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    L.nc3_get_att_double.argtypes = [vp, C.c_int, C.c_char_p, vp, C.c_int]
    L.nc3_copy_atts.argtypes = [vp, C.c_int, vp, C.c_int]
    return L


def arr(*v):
    return (C.c_size_t * len(v))(*v)


def ints(*v):
    return (C.c_int * len(v))(*v)


def test_written_files_are_read_by_scipy_and_back(L, tmp_path):
    d = str(tmp_path)
    for fmt in (1,2,5):
        p=os.path.join(d,'w%d.nc'%fmt).encode()
        err=C.create_string_buffer(256)
        f=L.nc3_create(p,fmt,err,256); assert f, err.value
        dt=L.nc3_def_dim(f,b'time',0); dn=L.nc3_def_dim(f,b'ncells',1000); d2=L.nc3_def_dim(f,b'two',2); ds=L.nc3_def_dim(f,b'string',255)
        v1=L.nc3_def_var(f,b'tile1',4,1,ints(dn)); L.nc3_put_att_text(f,v1,b'standard_name',b'tile_number_in_mosaic1')
        v2=L.nc3_def_var(f,b'tile1_cell',4,2,ints(dn,d2))
        v3=L.nc3_def_var(f,b'xgrid_area',6,1,ints(dn)); L.nc3_put_att_text(f,v3,b'units',b'm2')
        v4=L.nc3_def_var(f,b'fld',5,3,ints(dt,dn,d2)); mv=np.array([1e20]); L.nc3_put_att_double(f,v4,b'missing_value',5,1,mv.ctypes.data)
        v5=L.nc3_def_var(f,b'time',6,1,ints(dt))
        v6=L.nc3_def_var(f,b'name',2,1,ints(ds))
        v7=L.nc3_def_var(f,b'sh',3,2,ints(dt,dn))
        L.nc3_put_att_text(f,-1,b'history',b'test')
        assert L.nc3_enddef(f)==0, L.nc3_strerror(f)
        rng=np.random.default_rng(fmt)
        t1=rng.integers(1,7,1000).astype(np.int32); ii=rng.integers(1,100,1000).astype(np.int32); jj=rng.integers(1,100,1000).astype(np.int32)
        ar=rng.uniform(0,1e9,1000)
        assert L.nc3_put_var_int(f,v1,t1.ctypes.data)==0
        assert L.nc3_put_vara_int(f,v2,arr(0,0),arr(1000,1),ii.ctypes.data)==0, L.nc3_strerror(f)
        assert L.nc3_put_vara_int(f,v2,arr(0,1),arr(1000,1),jj.ctypes.data)==0
        assert L.nc3_put_var_double(f,v3,ar.ctypes.data)==0
        fld=rng.uniform(-1,1,(3,1000,2))
        for t in range(3):
            assert L.nc3_put_vara_double(f,v4,arr(t,0,0),arr(1,1000,2),fld[t].ctypes.data)==0, L.nc3_strerror(f)
            tv=np.array([float(t)+0.5]); assert L.nc3_put_vara_double(f,v5,arr(t),arr(1),tv.ctypes.data)==0
            sh=(np.arange(1000)+t).astype(np.float64); assert L.nc3_put_vara_double(f,v7,arr(t,0),arr(1,1000),sh.ctypes.data)==0
        nm=b'hello'.ljust(255,b'\0'); assert L.nc3_put_vara_text(f,v6,arr(0),arr(255),nm)==0
        assert L.nc3_close(f)==0
        if fmt!=5:
            g=netcdf_file(p.decode(),'r',mmap=False)
            assert g.version_byte==fmt
            assert np.array_equal(g.variables['tile1'][:],t1)
            assert np.array_equal(g.variables['tile1_cell'][:,0],ii) and np.array_equal(g.variables['tile1_cell'][:,1],jj)
            assert np.array_equal(g.variables['xgrid_area'][:],ar)
            assert np.array_equal(g.variables['fld'][:],fld.astype(np.float32))
            assert np.array_equal(g.variables['time'][:],[0.5,1.5,2.5])
            assert g.variables['sh'][2,5]==7
            assert g.variables['tile1'].standard_name==b'tile_number_in_mosaic1'
            assert g.history==b'test'
            assert abs(g.variables['fld'].missing_value-1e20)<1e14
            g.close()
        # read back with nc3
        f=L.nc3_open(p,err,256); assert f, err.value
        assert L.nc3_dim_len(f,L.nc3_dim_id(f,b'time'))==3
        out=np.empty((1000,),np.int32); assert L.nc3_get_vara_int(f,L.nc3_var_id(f,b'tile1_cell'),arr(0,1),arr(1000,1),out.ctypes.data)==0; assert np.array_equal(out,jj)
        o2=np.empty((2,500,1)); assert L.nc3_get_vara_double(f,L.nc3_var_id(f,b'fld'),arr(1,250,1),arr(2,500,1),o2.ctypes.data)==0, L.nc3_strerror(f)
        assert np.array_equal(o2[...,0], fld.astype(np.float32)[1:3,250:750,1])
        buf=C.create_string_buffer(64); assert L.nc3_get_att_text(f,L.nc3_var_id(f,b'xgrid_area'),b'units',buf,64)==0 and buf.value==b'm2'
        L.nc3_close(f)


def test_files_written_by_scipy_are_read(L, tmp_path):
    d = str(tmp_path)
    # file written by scipy -> nc3 reader
    p=os.path.join(d,'s.nc')
    g=netcdf_file(p,'w',version=2)
    g.createDimension('time',None); g.createDimension('y',7); g.createDimension('x',5)
    v=g.createVariable('t','f',('time','y','x')); v.missing_value=np.float32(-1e10); v.scale_factor=2.0
    w=g.createVariable('x','d',('x',)); w[:]=np.arange(5.)
    s=g.createVariable('s','h',('time','x'))
    data=np.random.default_rng(0).uniform(0,1,(4,7,5)).astype(np.float32)
    for t in range(4): v[t]=data[t]; s[t]=np.arange(5)+t
    g.close()
    err=C.create_string_buffer(256)
    f=L.nc3_open(p.encode(),err,256); assert f, err.value
    o=np.empty((4,7,5)); assert L.nc3_get_var_double(f,L.nc3_var_id(f,b't'),o.ctypes.data)==0, L.nc3_strerror(f)
    assert np.array_equal(o,data.astype(np.float64))
    o=np.empty((4,5),np.int32); assert L.nc3_get_var_int(f,L.nc3_var_id(f,b's'),o.ctypes.data)==0
    assert o[3,4]==7
    sf=np.zeros(1); assert L.nc3_get_att_double(f,L.nc3_var_id(f,b't'),b'scale_factor',sf.ctypes.data,1)==1 and sf[0]==2.0
    L.nc3_close(f)


def test_fixed_size_file_is_byte_identical_to_scipys(L, tmp_path):
    d = str(tmp_path)
    err = C.create_string_buffer(256)
    # byte identity with scipy for a no-record file (same schema order)
    p1=os.path.join(d,'a.nc'); p2=os.path.join(d,'b.nc')
    g=netcdf_file(p1,'w',version=1); g.createDimension('n',10); g.createDimension('two',2)
    a=g.createVariable('a','i',('n','two')); a.standard_name='x'; a[:]=np.arange(20).reshape(10,2)
    b=g.createVariable('b','d',('n',)); b.units='m2'; b[:]=np.arange(10.)*1.5
    g.close()
    f=L.nc3_create(p2.encode(),1,err,256)
    dn=L.nc3_def_dim(f,b'n',10); d2=L.nc3_def_dim(f,b'two',2)
    va=L.nc3_def_var(f,b'a',4,2,ints(dn,d2)); L.nc3_put_att_text(f,va,b'standard_name',b'x')
    vb=L.nc3_def_var(f,b'b',6,1,ints(dn)); L.nc3_put_att_text(f,vb,b'units',b'm2')
    L.nc3_enddef(f)
    x=np.arange(20,dtype=np.int32); L.nc3_put_var_int(f,va,x.ctypes.data); y=np.arange(10.)*1.5; L.nc3_put_var_double(f,vb,y.ctypes.data)
    L.nc3_close(f)
    assert open(p1, 'rb').read() == open(p2, 'rb').read()


def test_netcdf4_goes_to_the_hdf5_reader(L, tmp_path):
    """an HDF5 signature hands the file to csrc/h5r.c (tests/test_h5_cpu.py); an empty shell is refused with its message"""
    d = str(tmp_path)
    err = C.create_string_buffer(256)
    open(os.path.join(d,'h.nc'),'wb').write(b'\x89HDF\r\n\x1a\n'+b'\0'*100)
    assert not L.nc3_open(os.path.join(d, 'h.nc').encode(), err, 256)
    assert b'h5r' in err.value
    open(os.path.join(d,'x.nc'),'wb').write(b'GRIB'+b'\0'*100)
    assert not L.nc3_open(os.path.join(d, 'x.nc').encode(), err, 256)
    assert b'not a classic netCDF file' in err.value


def test_corrupt_headers_are_refused(L, tmp_path):
    """truncated files, absurd list counts and unknown types end in an error message, not in a crash"""
    import struct
    err = C.create_string_buffer(256)
    good = str(tmp_path / "good.nc")
    g = netcdf_file(good, "w", version=1); g.createDimension("n", 3); v = g.createVariable("v", "d", ("n",)); v[:] = [1.0, 2.0, 3.0]; g.close()
    raw = open(good, "rb").read()
    cases = {"short": raw[:20], "dims": raw[:8] + struct.pack(">ii", 0x0A, 0x7fffffff) + raw[16:],
             "magic": b"CDF\x07" + raw[4:], "empty": b""}
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".nc"))
        open(p, "wb").write(blob)
        assert not L.nc3_open(p.encode(), err, 256), name
        assert err.value, name
    f = L.nc3_open(good.encode(), err, 256)
    assert f
    out = np.zeros(3)
    assert L.nc3_get_var_double(f, L.nc3_var_id(f, b"v"), out.ctypes.data) == 0 and out.tolist() == [1.0, 2.0, 3.0]
    assert L.nc3_get_vara_double(f, 0, arr(2), arr(5), out.ctypes.data) != 0 and b"exceeds" in L.nc3_strerror(f)
    L.nc3_close(f)
