"""Golden vectors from the reference's OWN embedded polygon harness (create_xgrid.c:2355-3164, -Dtest_create_xgrid):
26 hand-built cases — poles, tripolar fold, identical boxes, containment, sides through the south pole, twin poles.
The harness only prints (%g); this script records the same cases in full precision.

Run in the build container only (needs /root/reference).  It extracts the text of the harness's `switch (n)` block from the
reference source into a throw-away C file under /tmp (nothing of it is committed), compiles it, runs every case to obtain
the INPUT polygons / grids exactly as the harness sets them up (degrees -> radians with the harness's D2R), and then calls
the unmodified reference functions (oracle/_ref/libfrenc_ref.so) the way the harness does:
   cases  1-10: latlon2xyz + clip_2dx2d_great_circle(.., n1 = 4, ..)                 (create_xgrid.c:3122-3127)
   cases 11-14: create_xgrid_great_circle on the small grids                          (:3046-3048)
   cases 15-26: clip_2dx2d, then fix_lon + poly_area of both inputs and the output    (:3089-3098)
plus, for every case whose two polygons are quadrilaterals, the reference's create_xgrid_2dx2d_order2 and
create_xgrid_great_circle on the pair taken as two 1x1 grids (what the product's C ABI can be asked directly).

    python tests/golden/make_polycases_golden.py      ->  tests/golden/ref_polycases.npz
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import xgtest  # noqa: E402

SRC = "/root/reference/tools/libfrencutils/create_xgrid.c"
MAXPOINT = 1000
D2R = np.pi / 180          # the harness: #define D2R (M_PI/180)


def build_case_lib():
    text = open(SRC).read()
    a = text.index("switch (n) {", text.index("#ifdef test_create_xgrid"))
    b = text.index("default:", a)
    body = text[a + len("switch (n) {"):b]
    code = """
#include <math.h>
#include <string.h>
#define MAXPOINT %d
int run_case(int n, double *lon1, double *lat1, double *lon2, double *lat2, int *dims)
{
  double lon1_in[MAXPOINT], lat1_in[MAXPOINT], lon2_in[MAXPOINT], lat2_in[MAXPOINT];
  int n1_in = 0, n2_in = 0, i, j;
  int nlon1 = 0, nlat1 = 0, nlon2 = 0, nlat2 = 0;
  memset(lon1_in, 0, sizeof(lon1_in)); memset(lat1_in, 0, sizeof(lat1_in));
  memset(lon2_in, 0, sizeof(lon2_in)); memset(lat2_in, 0, sizeof(lat2_in));
  switch (n) {
%s
  default: return -1;
  }
  memcpy(lon1, lon1_in, sizeof(lon1_in)); memcpy(lat1, lat1_in, sizeof(lat1_in));
  memcpy(lon2, lon2_in, sizeof(lon2_in)); memcpy(lat2, lat2_in, sizeof(lat2_in));
  dims[0] = n1_in; dims[1] = n2_in; dims[2] = nlon1; dims[3] = nlat1; dims[4] = nlon2; dims[5] = nlat2;
  (void)i; (void)j;
  return 0;
}
""" % (MAXPOINT, body)
    d = tempfile.mkdtemp(prefix="polycases_")
    cfile = os.path.join(d, "cases.c"); so = os.path.join(d, "cases.so")
    open(cfile, "w").write(code)
    subprocess.run(["gcc", "-O0", "-w", "-fPIC", "-shared", "-o", so, cfile, "-lm"], check=True)
    return C.CDLL(so)


def main():
    R = xgtest.ref_lib()
    assert R is not None, "reference not built"
    L = build_case_lib()
    dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    L.run_case.argtypes = [C.c_int, dp, dp, dp, dp, ip]
    R.latlon2xyz.argtypes = [C.c_int, dp, dp, dp, dp, dp]
    R.create_xgrid_great_circle.restype = C.c_int
    R.create_xgrid_2dx2d_order2.restype = C.c_int
    ci = lambda v: C.byref(C.c_int(v))
    pv = lambda a: a.ctypes.data_as(C.c_void_p)
    out = {}
    for n in range(1, 27):
        lon1 = np.zeros(MAXPOINT); lat1 = np.zeros(MAXPOINT); lon2 = np.zeros(MAXPOINT); lat2 = np.zeros(MAXPOINT)
        dims = np.zeros(6, np.int32)
        assert L.run_case(n, lon1, lat1, lon2, lat2, dims) == 0, n
        n1, n2, nlon1, nlat1, nlon2, nlat2 = (int(v) for v in dims)
        x1 = lon1[:n1] * D2R; y1 = lat1[:n1] * D2R; x2 = lon2[:n2] * D2R; y2 = lat2[:n2] * D2R     # create_xgrid.c:3017-3022
        k = f"c{n:02d}_"
        out[k + "lon1"], out[k + "lat1"], out[k + "lon2"], out[k + "lat2"] = x1, y1, x2, y2
        out[k + "dims"] = dims
        if n <= 10:
            m1 = 4                                           # the harness passes the literal 4 (:3125)
            a = [np.zeros(max(n1, 4)) for _ in range(3)]; b = [np.zeros(n2) for _ in range(3)]
            xx1 = np.zeros(max(n1, 4)); yy1 = np.zeros(max(n1, 4)); xx1[:n1] = x1; yy1[:n1] = y1
            R.latlon2xyz(max(n1, 4), xx1, yy1, *a)
            R.latlon2xyz(n2, np.ascontiguousarray(x2), np.ascontiguousarray(y2), *b)
            o = [np.zeros(50) for _ in range(3)]
            no = R.clip_2dx2d_great_circle(a[0], a[1], a[2], m1, b[0], b[1], b[2], n2, o[0], o[1], o[2])
            out[k + "gc_n"] = np.int32(no)
            out[k + "gc_xyz"] = np.stack([v[:no] for v in o])
            out[k + "xyz1"] = np.stack([v[:m1] for v in a]); out[k + "xyz2"] = np.stack(b)
        elif n <= 14:
            cap = 4096
            bi = [np.zeros(cap, np.int32) for _ in range(4)]
            xa = np.zeros(cap); xc = np.zeros(cap); yc = np.zeros(cap)
            mask = np.ones(nlon1 * nlat1)
            x1c, y1c, x2c, y2c = (np.ascontiguousarray(v) for v in (x1, y1, x2, y2))
            nx = R.create_xgrid_great_circle(ci(nlon1), ci(nlat1), ci(nlon2), ci(nlat2), pv(x1c), pv(y1c), pv(x2c), pv(y2c), pv(mask),
                                             pv(bi[0]), pv(bi[1]), pv(bi[2]), pv(bi[3]), pv(xa), pv(xc), pv(yc))
            out[k + "gcx_n"] = np.int32(nx)
            out[k + "gcx_idx"] = np.stack([v[:nx] for v in bi]); out[k + "gcx_area"] = xa[:nx].copy()
        else:
            lo = np.zeros(50); la = np.zeros(50)
            a1 = np.zeros(50); b1 = np.zeros(50); a2 = np.zeros(50); b2 = np.zeros(50)
            a1[:n1] = x1; b1[:n1] = y1; a2[:n2] = x2; b2[:n2] = y2
            no = R.clip_2dx2d(a1, b1, n1, a2, b2, n2, lo, la)
            out[k + "clip_n"] = np.int32(no)
            out[k + "clip_lon"] = lo[:no].copy(); out[k + "clip_lat"] = la[:no].copy()
            f1 = R.fix_lon(a1, b1, n1, np.pi); f2 = R.fix_lon(a2, b2, n2, np.pi); fo = R.fix_lon(lo, la, no, np.pi)
            out[k + "fix_n"] = np.array([f1, f2, fo], np.int32)
            out[k + "fix1"] = np.stack([a1[:f1], b1[:f1]]); out[k + "fix2"] = np.stack([a2[:f2], b2[:f2]])
            out[k + "fixo"] = np.stack([lo[:fo], la[:fo]])
            out[k + "areas"] = np.array([R.poly_area(a1, b1, f1), R.poly_area(a2, b2, f2), R.poly_area(lo, la, fo)])
        # quadrilateral pairs as two 1x1 grids (vertex order of a cell: (i,j) (i+1,j) (i+1,j+1) (i,j+1))
        if n1 == 4 and n2 == 4 and not (11 <= n <= 14):
            g = lambda v: np.ascontiguousarray(np.array([[v[0], v[1]], [v[3], v[2]]]))
            gl1, ga1, gl2, ga2 = g(x1), g(y1), g(x2), g(y2)
            bi = [np.zeros(64, np.int32) for _ in range(4)]
            xa = np.zeros(64); xc = np.zeros(64); yc = np.zeros(64)
            mask = np.ones(1)
            nx = R.create_xgrid_2dx2d_order2(ci(1), ci(1), ci(1), ci(1), pv(gl1), pv(ga1), pv(gl2), pv(ga2), pv(mask),
                                             pv(bi[0]), pv(bi[1]), pv(bi[2]), pv(bi[3]), pv(xa), pv(xc), pv(yc))
            out[k + "cell_o2"] = np.concatenate([[nx], xa[:nx], xc[:nx], yc[:nx]])
            nx = R.create_xgrid_great_circle(ci(1), ci(1), ci(1), ci(1), pv(gl1), pv(ga1), pv(gl2), pv(ga2), pv(mask),
                                             pv(bi[0]), pv(bi[1]), pv(bi[2]), pv(bi[3]), pv(xa), pv(xc), pv(yc))
            out[k + "cell_gc"] = np.concatenate([[nx], xa[:nx]])
    path = os.path.join(HERE, "ref_polycases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")
    for n in range(1, 27):
        k = f"c{n:02d}_"
        tag = "gc_n" if n <= 10 else "gcx_n" if n <= 14 else "clip_n"
        print(n, "dims", out[k + "dims"].tolist(), tag, int(out[k + tag]), "cell_o2" in "".join(out.keys()) and out.get(k + "cell_o2", [None])[0])


if __name__ == "__main__":
    main()
