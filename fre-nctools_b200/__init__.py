"""fre-nctools_b200 — B200-native conservative regridding (fregrid's exchange-grid path).

Host-side Python mirror of the reference's C interface for this path.  Everything here is a
thin ctypes binding over ``libxgrid_b200.so`` (include/xgrid_b200.h); there is no Python or
CPU implementation of the math — if the CUDA library or a GPU is missing the calls raise.

Mirrored reference entry points (tools/libfrencutils/create_xgrid.h:39-80):
    create_xgrid_2dx2d_order1 / create_xgrid_2dx2d_order2 / get_grid_area / get_maxxgrid
Batched interface (what tools/fregrid/conserve_interp.c:42 setup_conserve_interp does per
output tile, all source tiles at once, results left in HBM):
    XgridPlan

The directory name contains a hyphen, so import it with ``load_package()`` from
``__graft_entry__`` (or importlib) rather than a plain ``import``.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

CONSERVE_ORDER1 = 1       # globals.h:46
CONSERVE_ORDER2 = 2       # globals.h:47
GREAT_CIRCLE = 4096       # globals.h:58
MONOTONIC = 16384         # globals.h:60

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class XgridError(RuntimeError):
    pass


class _View(C.Structure):
    _fields_ = [("nxgrid", C.c_longlong),
                ("t_in", C.c_void_p), ("i_in", C.c_void_p), ("j_in", C.c_void_p),
                ("i_out", C.c_void_p), ("j_out", C.c_void_p),
                ("area", C.c_void_p), ("di", C.c_void_p), ("dj", C.c_void_p),
                ("xgrid_clon", C.c_void_p), ("xgrid_clat", C.c_void_p)]


def lib():
    """Load (building in-tree if stale and nvcc is present) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libxgrid_b200.so")
    dev = os.environ.get("XGRID_B200_LIB")        # developer override: a kernel-variant build (scripts/clip_variants.py)
    if dev:
        path = dev
    elif _build.needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if os.path.exists(nvcc):
            _build.build()
        elif not os.path.exists(path):
            raise XgridError("libxgrid_b200.so is not built and nvcc is unavailable; there is no CPU fallback")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.xgb_last_error.restype = C.c_char_p
    L.xgb_plan_create.restype = vp
    L.xgb_plan_create.argtypes = [C.c_int]
    L.xgb_plan_destroy.argtypes = [vp]
    L.xgb_plan_stream.restype = vp
    L.xgb_plan_stream.argtypes = [vp]
    L.xgb_plan_sync.argtypes = [vp]
    L.xgb_plan_set_dst.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int]
    L.xgb_plan_set_dst_latlon.argtypes = [vp, C.c_int, C.c_int] + [C.c_double] * 4
    L.xgb_plan_set_src.argtypes = [vp, C.c_int, _ip, _ip, vp, vp, vp, C.c_int]
    L.xgb_plan_set_src_window.argtypes = [vp, C.c_longlong, C.c_longlong]
    L.xgb_plan_set_src_sharded.argtypes = [vp, C.c_int, _ip, _ip, vp, vp, vp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.xgb_plan_partition.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong)]
    L.xgb_plan_partition_shares.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.xgb_plan_set_src_windows.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.xgb_plan_window_counts.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.xgb_plan_generate_async.argtypes = [vp, C.c_uint]
    L.xgb_plan_generate_finish.restype = C.c_longlong
    L.xgb_plan_generate_finish.argtypes = [vp]
    L.xgb_plan_window_counts_device.argtypes = [vp, vp]
    L.xgb_plan_generate.restype = C.c_longlong
    L.xgb_plan_generate.argtypes = [vp, C.c_uint]
    L.xgb_plan_generate_to_host.restype = C.c_longlong
    L.xgb_plan_generate_to_host.argtypes = [vp, C.c_uint, C.c_int, C.c_longlong] + [vp] * 10
    L.xgb_plan_last_npairs.restype = C.c_longlong
    L.xgb_plan_last_npairs.argtypes = [vp]
    L.xgb_plan_result_device.argtypes = [vp, C.POINTER(_View)]
    L.xgb_plan_result_host.argtypes = [vp] + [vp] * 8
    L.xgb_plan_result_centroids_host.argtypes = [vp, vp, vp]
    L.xgb_plan_src_area_host.argtypes = [vp, vp]
    L.xgb_plan_dst_area_host.argtypes = [vp, vp]
    L.xgb_plan_apply_setup.argtypes = [vp]
    L.xgb_plan_set_xgrid.argtypes = [vp, C.c_int, _ip, _ip, C.c_int, C.c_int, C.c_longlong] + [vp] * 8 + [C.c_int]
    L.xgb_plan_apply_nxgrid.restype = C.c_longlong
    L.xgb_plan_apply_nxgrid.argtypes = [vp]
    L.xgb_plan_grad_setup.argtypes = [vp, vp, vp, C.c_int]
    L.xgb_plan_grad_set_metrics.argtypes = [vp, C.c_int] + [vp] * 11 + [C.c_int]
    L.xgb_plan_grad_get_metrics.argtypes = [vp, C.c_int] + [vp] * 11
    L.xgb_plan_grad_c2l.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_double, C.c_int]
    L.xgb_plan_apply.argtypes = [vp, C.c_uint, C.c_int, vp, vp, vp, vp, C.c_int, C.c_double, vp, C.c_int]
    L.xgb_plan_regrid.argtypes = [vp, C.c_uint, C.c_int, vp, C.c_int, C.c_double, vp, C.c_int]
    L.xgb_plan_apply_options.argtypes = [vp, C.c_int, vp, vp, vp, C.c_double, C.c_int, vp, C.c_int]
    L.xgb_plan_great_circle_area_host.argtypes = [vp, C.c_int, vp]
    L.xgb_gc_clip_host.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp]
    L.create_xgrid_great_circle.restype = C.c_int
    L.xgb_cubed_sphere_grid.argtypes = [C.c_int, vp, vp, vp, vp]
    L.xgb_latlon_grid.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp]
    L.xgb_plan_phase_ms.argtypes = [vp, vp, vp, vp]
    L.xgb_plan_reset_phase_ms.restype = None
    L.xgb_plan_reset_phase_ms.argtypes = [vp]
    L.xgb_kernel_launches.restype = C.c_longlong
    L.xgb_fp64_peak_tflops.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.xgb_ref_trig_host.restype = None
    L.xgb_ref_trig_host.argtypes = [C.c_longlong] + [vp] * 5
    L.xgb_ref_trig_device.argtypes = [C.c_longlong] + [vp] * 5
    L.xgb_ref_trig_site_host.restype = None
    L.xgb_ref_trig_site_host.argtypes = [C.c_longlong] + [vp] * 5
    L.xgb_ref_trig_site_device.argtypes = [C.c_longlong] + [vp] * 5
    L.xgb_set_nc_format.argtypes = [C.c_char_p]
    L.xgb_remap_write.argtypes = [C.c_char_p, C.c_int, C.c_longlong] + [vp] * 5 + [C.c_int, C.c_int] + [vp] * 3
    L.xgb_remap_size.restype = C.c_longlong
    L.xgb_remap_size.argtypes = [C.c_char_p]
    L.xgb_remap_read.argtypes = [C.c_char_p, C.c_int, C.c_longlong] + [vp] * 8
    L.xgb_poly_moments_site_host.restype = None
    L.xgb_poly_moments_site_host.argtypes = [C.c_int, C.c_int, vp, vp, C.c_double, vp]
    L.get_maxxgrid.restype = C.c_int
    for name in ("create_xgrid_2dx2d_order1", "create_xgrid_2dx2d_order2", "create_xgrid_1dx2d_order1", "create_xgrid_1dx2d_order2",
                 "create_xgrid_2dx1d_order1", "create_xgrid_2dx1d_order2"):
        getattr(L, name).restype = C.c_int
    _LIB = L
    return L


def _err():
    return lib().xgb_last_error().decode(errors="replace")


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _f64_ptr(a):
    """-> (pointer int, on_device flag, keepalive) for a numpy array or a CUDA torch tensor."""
    if a is None:
        return None, 0, None
    if _is_torch(a):
        import torch
        if a.dtype != torch.float64:
            raise TypeError("expected float64 tensor")
        a = a.contiguous()
        return a.data_ptr(), (1 if a.is_cuda else 0), a
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr.ctypes.data, 0, arr


# ---------------------------------------------------------------------------------------------
# grid synthesis (host side, input preparation only)
# ---------------------------------------------------------------------------------------------
def cubed_sphere_grid(ni, centers=False):
    """make_hgrid 'gnomonic_ed' C<ni> grid as fregrid reads it: (lonc, latc) each [6, ni+1, ni+1] radians."""
    lonc = np.empty((6, ni + 1, ni + 1)); latc = np.empty_like(lonc)
    lont = np.empty((6, ni, ni)) if centers else None
    latt = np.empty((6, ni, ni)) if centers else None
    rc = lib().xgb_cubed_sphere_grid(ni, lonc.ctypes.data, latc.ctypes.data,
                                     lont.ctypes.data if centers else None, latt.ctypes.data if centers else None)
    if rc:
        raise XgridError("xgb_cubed_sphere_grid failed")
    return (lonc, latc, lont, latt) if centers else (lonc, latc)


def latlon_grid(nlon, nlat, lonbegin=0.0, lonend=360.0, latbegin=-90.0, latend=90.0):
    """fregrid --nlon/--nlat output grid (fregrid_util.c:588-603): (lonc, latc) each [nlat+1, nlon+1] radians."""
    lonc = np.empty((nlat + 1, nlon + 1)); latc = np.empty_like(lonc)
    if lib().xgb_latlon_grid(nlon, nlat, lonbegin, lonend, latbegin, latend, lonc.ctypes.data, latc.ctypes.data):
        raise XgridError("xgb_latlon_grid failed")
    return lonc, latc


# ---------------------------------------------------------------------------------------------
# batched plan
# ---------------------------------------------------------------------------------------------
class _CudaArray:
    """Zero-copy device array descriptor (CUDA array interface v3) for torch.as_tensor."""

    def __init__(self, ptr, n, typestr, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
        self._owner = owner


class XgridPlan:
    """Exchange-grid generation for one destination tile and a source mosaic on one GPU."""

    def __init__(self, device=0):
        self._L = lib()
        self._p = self._L.xgb_plan_create(int(device))
        if not self._p:
            raise XgridError(_err())
        self.device = int(device)
        self.ncell_src = 0
        self.nxgrid = -1
        self.order = 0

    def close(self):
        if getattr(self, "_p", None):
            self._L.xgb_plan_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise XgridError(_err())

    @property
    def stream(self):
        return self._L.xgb_plan_stream(self._p)

    def sync(self):
        self._ck(self._L.xgb_plan_sync(self._p))

    def set_dst(self, lon, lat):
        """lon/lat: [ny+1, nx+1] vertex arrays (numpy, or float64 CUDA tensors)."""
        ny, nx = lon.shape[0] - 1, lon.shape[1] - 1
        pl, dl, kl = _f64_ptr(lon); pa, da, ka = _f64_ptr(lat)
        if dl != da:
            raise TypeError("lon and lat must live on the same side")
        self._ck(self._L.xgb_plan_set_dst(self._p, nx, ny, pl, pa, dl))
        self.nx_dst, self.ny_dst = nx, ny

    def set_dst_latlon(self, nlon, nlat, lonbegin=0.0, lonend=360.0, latbegin=-90.0, latend=90.0):
        """fregrid's --nlon/--nlat output grid, built on the device (no upload); degrees."""
        self._ck(self._L.xgb_plan_set_dst_latlon(self._p, int(nlon), int(nlat), float(lonbegin), float(lonend), float(latbegin), float(latend)))
        self.nx_dst, self.ny_dst = int(nlon), int(nlat)

    def set_src(self, lons, lats, mask=None):
        """lons/lats: list of per-tile [ny+1, nx+1] arrays, or one [ntiles, ny+1, nx+1] array."""
        if not isinstance(lons, (list, tuple)):
            if lons.ndim == 3 and not _is_torch(lons) and lons.flags["C_CONTIGUOUS"] and lats.flags["C_CONTIGUOUS"]:
                nt, nyp, nxp = lons.shape          # already concatenated in memory: no staging copy
                return self.set_src_flat([nxp - 1] * nt, [nyp - 1] * nt, lons.reshape(-1), lats.reshape(-1), mask)
            lons = [lons[t] for t in range(lons.shape[0])] if lons.ndim == 3 else [lons]
            lats = [lats[t] for t in range(lats.shape[0])] if lats.ndim == 3 else [lats]
        nx = [int(a.shape[1] - 1) for a in lons]
        ny = [int(a.shape[0] - 1) for a in lons]
        if _is_torch(lons[0]):
            import torch
            lon = torch.cat([a.reshape(-1) for a in lons]); lat = torch.cat([a.reshape(-1) for a in lats])
        else:
            lon = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in lons])
            lat = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in lats])
        return self.set_src_flat(nx, ny, lon, lat, mask)

    def set_src_flat(self, nx, ny, lon, lat, mask=None):
        """tiles already concatenated: lon/lat flat arrays of sum((nx+1)*(ny+1)) vertices (host numpy, pinned or not, or CUDA tensors)."""
        nxa = (C.c_int * len(nx))(*[int(v) for v in nx])
        nya = (C.c_int * len(ny))(*[int(v) for v in ny])
        pl, dl, kl = _f64_ptr(lon); pa, da, ka = _f64_ptr(lat)
        pm, dm, km = _f64_ptr(mask)
        if dl != da or (mask is not None and dm != dl):
            raise TypeError("lon, lat and mask must live on the same side")
        self._ck(self._L.xgb_plan_set_src(self._p, len(nx), nxa, nya, pl, pa, pm, dl))
        self.tiles = [(int(a), int(b)) for a, b in zip(nx, ny)]
        self.ncell_src = sum(a * b for a, b in self.tiles)

    def set_src_sharded(self, nx, ny, lon, lat, windows, mask=None):
        """the mosaic for a rank that only generates `windows` [(begin, end), ...]: uploads the vertex rows they touch and
        precomputes their cells only (xgb_plan_set_src_sharded); lon/lat: flat host arrays of the WHOLE mosaic (pinned for
        asynchronous copies).  -> bytes copied to the device"""
        nxa = (C.c_int * len(nx))(*[int(v) for v in nx]); nya = (C.c_int * len(ny))(*[int(v) for v in ny])
        pl, dl, kl = _f64_ptr(lon); pa, da, ka = _f64_ptr(lat); pm, dm, km = _f64_ptr(mask)
        if dl or da or dm:
            raise TypeError("xgb_plan_set_src_sharded takes host arrays")
        n = len(windows)
        b = (C.c_longlong * n)(*[int(w[0]) for w in windows]); e = (C.c_longlong * n)(*[int(w[1]) for w in windows])
        self._ck(self._L.xgb_plan_set_src_sharded(self._p, len(nx), nxa, nya, pl, pa, pm, n, b, e))
        self.tiles = [(int(a), int(c)) for a, c in zip(nx, ny)]
        self.ncell_src = sum(a * c for a, c in self.tiles)
        self._nwin = n
        nbytes, off = 0, 0
        for a, c in self.tiles:                          # what the call copied: vertex rows j0 .. j1+1 of every (window, tile) overlap
            for wb, we in windows:
                lo, hi = max(wb, off), min(we, off + a * c)
                if hi > lo:
                    nbytes += ((hi - 1 - off) // a - (lo - off) // a + 2) * (a + 1) * 16 + (8 * (hi - lo) if mask is not None else 0)
            off += a * c
        return nbytes

    def set_src_window(self, begin, end):
        self._ck(self._L.xgb_plan_set_src_window(self._p, int(begin), int(end)))
        self._nwin = 1

    def set_src_windows(self, windows):
        """several source-cell windows [(begin, end), ...] (at most 64), generated one after the other in one call"""
        n = len(windows)
        b = (C.c_longlong * n)(*[int(w[0]) for w in windows]); e = (C.c_longlong * n)(*[int(w[1]) for w in windows])
        self._ck(self._L.xgb_plan_set_src_windows(self._p, n, b, e))
        self._nwin = n

    def window_counts(self):
        """exchange cells each window of the last generate produced"""
        n = getattr(self, "_nwin", 1)
        c = (C.c_longlong * n)()
        self._ck(self._L.xgb_plan_window_counts(self._p, c))
        return [int(v) for v in c]

    def partition(self, nparts, shares=None):
        """bounds of nparts contiguous source-cell windows: equal candidate-pair counts, or pair counts in the proportions `shares`"""
        b = (C.c_longlong * (nparts + 1))()
        if shares is None:
            self._ck(self._L.xgb_plan_partition(self._p, nparts, b))
        else:
            if len(shares) != nparts:
                raise ValueError("partition: one share per window")
            sh = (C.c_double * nparts)(*[float(v) for v in shares])
            self._ck(self._L.xgb_plan_partition_shares(self._p, nparts, sh, b))
        return [int(v) for v in b]

    def generate(self, opcode):
        n = self._L.xgb_plan_generate(self._p, int(opcode))
        if n < 0:
            raise XgridError(_err())
        self.nxgrid = int(n)
        self.order = 2 if (opcode & CONSERVE_ORDER2) else 1
        return self.nxgrid

    def generate_async(self, opcode):
        """enqueue a generate on the plan's stream without waiting (buffers as sized by the last generate())"""
        self._ck(self._L.xgb_plan_generate_async(self._p, int(opcode)))
        self._pending_order = 2 if (opcode & CONSERVE_ORDER2) else 1

    def generate_finish(self):
        n = self._L.xgb_plan_generate_finish(self._p)
        if n < 0:
            raise XgridError(_err())
        self.nxgrid = int(n)
        self.order = self._pending_order
        return self.nxgrid

    def window_counts_device(self, counts):
        """per-window exchange-cell counts of the window last enqueued -> int64 CUDA tensor (nwin entries), on the plan's stream"""
        self._ck(self._L.xgb_plan_window_counts_device(self._p, counts.data_ptr()))

    def generate_to_host(self, opcode, bufs, nchunks=8):
        """generate + download overlapped (xgb_plan_generate_to_host): bufs = dict of preallocated host arrays (numpy or
        pinned torch CPU tensors) keyed t_in, i_in, j_in, i_out, j_out, area[, di, dj, xgrid_clon, xgrid_clat]; returns nxgrid"""
        def ptr(k):
            b = bufs.get(k)
            if b is None:
                return None
            return b.data_ptr() if _is_torch(b) else b.ctypes.data
        cap = min(int(bufs[k].shape[0]) for k in bufs if bufs[k] is not None)
        n = self._L.xgb_plan_generate_to_host(self._p, int(opcode), int(nchunks), cap, ptr("t_in"), ptr("i_in"), ptr("j_in"),
                                              ptr("i_out"), ptr("j_out"), ptr("area"), ptr("di"), ptr("dj"),
                                              ptr("xgrid_clon"), ptr("xgrid_clat"))
        if n < 0:
            raise XgridError(_err())
        self.nxgrid = int(n)
        self.order = 2 if (opcode & CONSERVE_ORDER2) else 1
        return self.nxgrid

    @property
    def npairs(self):
        return int(self._L.xgb_plan_last_npairs(self._p))

    PHASES = ("candidate_count", "candidate_fill", "clip", "scan", "scatter_finalize")

    def phase_ms(self):
        """(last generate's per-phase ms, accumulated per-phase ms, number of generates accumulated)"""
        last = (C.c_float * 5)(); acc = (C.c_double * 5)(); n = C.c_longlong(0)
        self._ck(self._L.xgb_plan_phase_ms(self._p, last, acc, C.byref(n)))
        return dict(zip(self.PHASES, list(last))), dict(zip(self.PHASES, list(acc))), int(n.value)

    def reset_phase_ms(self):
        self._L.xgb_plan_reset_phase_ms(self._p)

    def result_host_into(self, bufs):
        """D2H into caller-provided (e.g. pinned) buffers: dict of numpy arrays / torch CPU tensors keyed like result_host()."""
        def ptr(k):
            b = bufs.get(k)
            if b is None:
                return None
            return b.data_ptr() if _is_torch(b) else b.ctypes.data
        self._ck(self._L.xgb_plan_result_host(self._p, ptr("t_in"), ptr("i_in"), ptr("j_in"), ptr("i_out"), ptr("j_out"),
                                              ptr("area"), ptr("di"), ptr("dj")))

    def result_host(self):
        n = self.nxgrid
        out = {k: np.empty(n, np.int32) for k in ("t_in", "i_in", "j_in", "i_out", "j_out")}
        out["area"] = np.empty(n)
        if self.order == 2:
            out["di"] = np.empty(n); out["dj"] = np.empty(n)
        args = [out[k].ctypes.data for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area")]
        args += [out["di"].ctypes.data, out["dj"].ctypes.data] if self.order == 2 else [None, None]
        self._ck(self._L.xgb_plan_result_host(self._p, *args))
        if self.order == 2:
            out["xgrid_clon"] = np.empty(n); out["xgrid_clat"] = np.empty(n)
            self._ck(self._L.xgb_plan_result_centroids_host(self._p, out["xgrid_clon"].ctypes.data, out["xgrid_clat"].ctypes.data))
        return out

    def result_device(self):
        """Zero-copy torch views of the device-resident result (valid until the next generate)."""
        import torch
        v = _View()
        self._ck(self._L.xgb_plan_result_device(self._p, C.byref(v)))
        n = int(v.nxgrid)
        dev = torch.device("cuda", self.device)
        out = {}
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out"):
            out[k] = torch.as_tensor(_CudaArray(getattr(v, k), n, "<i4", self), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev)
        for k in ("area", "di", "dj", "xgrid_clon", "xgrid_clat"):
            ptr = getattr(v, k)
            if ptr:
                out[k] = torch.as_tensor(_CudaArray(ptr, n, "<f8", self), device=dev) if n else torch.empty(0, dtype=torch.float64, device=dev)
        return out

    # ---- conservative apply (do_scalar_conserve_interp / grad_c2l), batched over field-levels -------------
    METRICS = ("dx", "dy", "area", "edge_w", "edge_e", "edge_s", "edge_n", "en_n", "en_e", "vlon", "vlat")

    def apply_setup(self):
        """regroup the last generate()'s exchange grid by destination cell (one-time)"""
        self._ck(self._L.xgb_plan_apply_setup(self._p))
        self._apply_tiles = list(self.tiles)
        self._apply_dst = (self.nx_dst, self.ny_dst)

    def set_xgrid(self, tiles, nx_out, ny_out, x):
        """exchange-grid lists from elsewhere (remap file, other GPUs): tiles = [(nx, ny), ...]; x = dict with
        t_in, i_in, j_in, i_out, j_out (int32), area and optionally di, dj (float64); numpy or CUDA tensors"""
        nxa = (C.c_int * len(tiles))(*[int(t[0]) for t in tiles])
        nya = (C.c_int * len(tiles))(*[int(t[1]) for t in tiles])
        n = int(x["area"].shape[0])
        dev = 1 if _is_torch(x["area"]) and x["area"].is_cuda else 0
        keep, ptrs = [], []
        for k in ("t_in", "i_in", "j_in", "i_out", "j_out", "area", "di", "dj"):
            v = x.get(k)
            if v is None:
                ptrs.append(None)
                continue
            if _is_torch(v):
                v = v.contiguous()
                ptrs.append(v.data_ptr())
            else:
                v = np.ascontiguousarray(v, dtype=np.float64 if k in ("area", "di", "dj") else np.int32)
                ptrs.append(v.ctypes.data)
            keep.append(v)
        self._ck(self._L.xgb_plan_set_xgrid(self._p, len(tiles), nxa, nya, int(nx_out), int(ny_out), n, *ptrs, dev))
        self._apply_tiles = [(int(a), int(b)) for a, b in tiles]
        self._apply_dst = (int(nx_out), int(ny_out))

    def _field_sizes(self):
        ncell = sum(a * b for a, b in self._apply_tiles)
        nhalo = sum((a + 2) * (b + 2) for a, b in self._apply_tiles)
        return ncell, nhalo, self._apply_dst[0] * self._apply_dst[1]

    def grad_setup(self, lont, latt):
        """calc_c2l_grid_info on the device; lont/latt: haloed cell centres, tiles concatenated (flat)"""
        pl, dl, kl = _f64_ptr(lont); pa, da, ka = _f64_ptr(latt)
        if not hasattr(self, "_apply_tiles"):
            self._apply_tiles = list(self.tiles)
        self._ck(self._L.xgb_plan_grad_setup(self._p, pl, pa, dl))

    def grad_set_metrics(self, tile, m):
        """the reference's own metrics for one tile: dict keyed like METRICS (host numpy arrays)"""
        arrs = [np.ascontiguousarray(m[k], np.float64) for k in self.METRICS]
        self._ck(self._L.xgb_plan_grad_set_metrics(self._p, int(tile), *[a.ctypes.data for a in arrs], 0))

    def grad_get_metrics(self, tile):
        nx, ny = self._apply_tiles[tile]
        sizes = {"dx": nx * (ny + 1), "dy": (nx + 1) * ny, "area": nx * ny, "edge_w": ny + 1, "edge_e": ny + 1, "edge_s": nx + 1,
                 "edge_n": nx + 1, "en_n": 3 * nx * (ny + 1), "en_e": 3 * (nx + 1) * ny, "vlon": 3 * nx * ny, "vlat": 3 * nx * ny}
        out = {k: np.empty(sizes[k]) for k in self.METRICS}
        self._ck(self._L.xgb_plan_grad_get_metrics(self._p, int(tile), *[out[k].ctypes.data for k in self.METRICS]))
        return out

    def _alloc_like(self, ref, n, dtype_np, dtype_t):
        if _is_torch(ref) and ref.is_cuda:
            import torch
            return torch.empty(n, dtype=getattr(torch, dtype_t), device=ref.device)
        return np.empty(n, dtype_np)

    @staticmethod
    def _ptr(a):
        if a is None:
            return None
        return a.data_ptr() if _is_torch(a) else a.ctypes.data

    def grad_c2l(self, data, nfields=1, has_missing=False, missing=0.0, with_mask=True):
        """-> (grad_x, grad_y, grad_mask) for nfields haloed field-levels (flat, back to back)"""
        ncell, nhalo, ndst = self._field_sizes()
        pd, dev, kd = _f64_ptr(data)
        gx = self._alloc_like(kd, nfields * ncell, np.float64, "float64")
        gy = self._alloc_like(kd, nfields * ncell, np.float64, "float64")
        gm = self._alloc_like(kd, nfields * ncell, np.int32, "int32") if with_mask else None
        self._ck(self._L.xgb_plan_grad_c2l(self._p, int(nfields), pd, self._ptr(gx), self._ptr(gy), self._ptr(gm),
                                           int(bool(has_missing)), float(missing), dev))
        return gx, gy, gm

    def apply(self, opcode, data, nfields=1, grad_x=None, grad_y=None, grad_mask=None, has_missing=False, missing=0.0, out=None):
        """do_scalar_conserve_interp for nfields field-levels -> flat [nfields * nx_out * ny_out]"""
        ncell, nhalo, ndst = self._field_sizes()
        pd, dev, kd = _f64_ptr(data)
        px, dx_, kx = _f64_ptr(grad_x); py, dy_, ky = _f64_ptr(grad_y)
        if grad_mask is not None and not _is_torch(grad_mask):
            grad_mask = np.ascontiguousarray(grad_mask, np.int32)
        if out is None:
            out = self._alloc_like(kd, nfields * ndst, np.float64, "float64")
        self._ck(self._L.xgb_plan_apply(self._p, int(opcode), int(nfields), pd, px, py, self._ptr(grad_mask),
                                        int(bool(has_missing)), float(missing), self._ptr(out), dev))
        return out

    def apply_options(self, cell_methods=0, weight=None, src_cell_area=None, field_area=None, area_missing=-1e20,
                      target_grid=False, dst_cell_area=None):
        """weight field / cell_methods sum / cell_measures / --target_grid for the following apply()/regrid() calls
        (host numpy arrays, tiles concatenated); call without arguments to return to the plain mean"""
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (weight, src_cell_area, field_area, dst_cell_area)]
        ptr = [None if a is None else a.ctypes.data for a in arrs]
        self._ck(self._L.xgb_plan_apply_options(self._p, int(cell_methods), ptr[0], ptr[1], ptr[2], float(area_missing),
                                                int(bool(target_grid)), ptr[3], 0))

    def regrid(self, opcode, data, nfields=1, has_missing=False, missing=0.0, out=None):
        """gradient (order 2) + apply in one call -> flat [nfields * nx_out * ny_out]"""
        ncell, nhalo, ndst = self._field_sizes()
        pd, dev, kd = _f64_ptr(data)
        if out is None:
            out = self._alloc_like(kd, nfields * ndst, np.float64, "float64")
        self._ck(self._L.xgb_plan_regrid(self._p, int(opcode), int(nfields), pd, int(bool(has_missing)), float(missing),
                                         self._ptr(out), dev))
        return out

    def src_area(self):
        a = np.empty(self.ncell_src)
        self._ck(self._L.xgb_plan_src_area_host(self._p, a.ctypes.data))
        return a

    def great_circle_area(self, which="src"):
        """get_grid_great_circle_area on the device (create_xgrid.c:98): 'src' (tiles concatenated) or 'dst'"""
        a = np.empty(self.ncell_src if which == "src" else self.nx_dst * self.ny_dst)
        self._ck(self._L.xgb_plan_great_circle_area_host(self._p, 0 if which == "src" else 1, a.ctypes.data))
        return a

    def dst_area(self):
        a = np.empty(self.nx_dst * self.ny_dst)
        self._ck(self._L.xgb_plan_dst_area_host(self._p, a.ctypes.data))
        return a


def kernel_launches():
    return int(lib().xgb_kernel_launches())


def fp64_peak_tflops(device=0):
    v = C.c_double(0)
    if lib().xgb_fp64_peak_tflops(int(device), C.byref(v)):
        raise XgridError(_err())
    return float(v.value)


# ---------------------------------------------------------------------------------------------
# reference-signature calls (host buffers in, host buffers out), as fregrid makes them
# (conserve_interp.c:186-200): outputs are allocated at get_maxxgrid() like the reference's callers.
# ---------------------------------------------------------------------------------------------
def get_maxxgrid():
    return int(lib().get_maxxgrid())


def _create_xgrid(order, lon_in, lat_in, lon_out, lat_out, mask_in=None, capacity=None):
    L = lib()
    lon_in = np.ascontiguousarray(lon_in, np.float64); lat_in = np.ascontiguousarray(lat_in, np.float64)
    lon_out = np.ascontiguousarray(lon_out, np.float64); lat_out = np.ascontiguousarray(lat_out, np.float64)
    nlat_in, nlon_in = lon_in.shape[0] - 1, lon_in.shape[1] - 1
    nlat_out, nlon_out = lon_out.shape[0] - 1, lon_out.shape[1] - 1
    if mask_in is None:
        mask_in = np.ones(nlon_in * nlat_in)
    mask_in = np.ascontiguousarray(mask_in, np.float64)
    cap = int(capacity or get_maxxgrid())
    ii, ji, io, jo = (np.empty(cap, np.int32) for _ in range(4))
    xa = np.empty(cap)
    ci = lambda v: C.byref(C.c_int(v))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if order == 1:
        n = L.create_xgrid_2dx2d_order1(ci(nlon_in), ci(nlat_in), ci(nlon_out), ci(nlat_out), p(lon_in), p(lat_in),
                                        p(lon_out), p(lat_out), p(mask_in), p(ii), p(ji), p(io), p(jo), p(xa))
        return n, ii[:n], ji[:n], io[:n], jo[:n], xa[:n]
    xc = np.empty(cap); yc = np.empty(cap)
    n = L.create_xgrid_2dx2d_order2(ci(nlon_in), ci(nlat_in), ci(nlon_out), ci(nlat_out), p(lon_in), p(lat_in),
                                    p(lon_out), p(lat_out), p(mask_in), p(ii), p(ji), p(io), p(jo), p(xa), p(xc), p(yc))
    return n, ii[:n], ji[:n], io[:n], jo[:n], xa[:n], xc[:n], yc[:n]


def create_xgrid_2dx2d_order1(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:621 — returns (nxgrid, i_in, j_in, i_out, j_out, xgrid_area)."""
    return _create_xgrid(1, lon_in, lat_in, lon_out, lat_out, mask_in)


def create_xgrid_2dx2d_order2(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:893 — returns (nxgrid, i_in, j_in, i_out, j_out, xgrid_area, xgrid_clon, xgrid_clat)."""
    return _create_xgrid(2, lon_in, lat_in, lon_out, lat_out, mask_in)


def _create_xgrid_box(kind, order, lon_in, lat_in, lon_out, lat_out, mask_in=None, capacity=None):
    """create_xgrid_1dx2d_* (kind "1dx2d": lon_in/lat_in are 1-D cell bounds, lon_out/lat_out 2-D vertex arrays) and
    create_xgrid_2dx1d_* (kind "2dx1d": the other way round), create_xgrid.c:208-598"""
    L = lib()
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    lon_in, lat_in, lon_out, lat_out = f64(lon_in), f64(lat_in), f64(lon_out), f64(lat_out)
    if kind == "1dx2d":
        nlon_in, nlat_in = lon_in.size - 1, lat_in.size - 1
        nlat_out, nlon_out = lon_out.shape[0] - 1, lon_out.shape[1] - 1
    else:
        nlat_in, nlon_in = lon_in.shape[0] - 1, lon_in.shape[1] - 1
        nlon_out, nlat_out = lon_out.size - 1, lat_out.size - 1
    mask_in = np.ones(nlon_in * nlat_in) if mask_in is None else f64(mask_in)
    cap = int(capacity or get_maxxgrid())
    ii, ji, io, jo = (np.empty(cap, np.int32) for _ in range(4))
    xa = np.empty(cap); xc = np.empty(cap); yc = np.empty(cap)
    ci = lambda v: C.byref(C.c_int(v))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn = getattr(L, f"create_xgrid_{kind}_order{order}")
    args = [ci(nlon_in), ci(nlat_in), ci(nlon_out), ci(nlat_out), p(lon_in), p(lat_in), p(lon_out), p(lat_out), p(mask_in),
            p(ii), p(ji), p(io), p(jo), p(xa)] + ([p(xc), p(yc)] if order == 2 else [])
    n = fn(*args)
    out = (n, ii[:n], ji[:n], io[:n], jo[:n], xa[:n])
    return out + (xc[:n], yc[:n]) if order == 2 else out


def create_xgrid_1dx2d_order1(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:208 — 1-D input cell bounds x 2-D output grid -> (nxgrid, i_in, j_in, i_out, j_out, xgrid_area)"""
    return _create_xgrid_box("1dx2d", 1, lon_in, lat_in, lon_out, lat_out, mask_in)


def create_xgrid_1dx2d_order2(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:312 — ... + (xgrid_clon, xgrid_clat)"""
    return _create_xgrid_box("1dx2d", 2, lon_in, lat_in, lon_out, lat_out, mask_in)


def create_xgrid_2dx1d_order1(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:413 — 2-D input grid x 1-D output cell bounds"""
    return _create_xgrid_box("2dx1d", 1, lon_in, lat_in, lon_out, lat_out, mask_in)


def create_xgrid_2dx1d_order2(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:514"""
    return _create_xgrid_box("2dx1d", 2, lon_in, lat_in, lon_out, lat_out, mask_in)


def create_xgrid_great_circle(lon_in, lat_in, lon_out, lat_out, mask_in=None):
    """create_xgrid.c:1366 — returns (nxgrid, i_in, j_in, i_out, j_out, xgrid_area, xgrid_clon, xgrid_clat)."""
    L = lib()
    lon_in = np.ascontiguousarray(lon_in, np.float64); lat_in = np.ascontiguousarray(lat_in, np.float64)
    lon_out = np.ascontiguousarray(lon_out, np.float64); lat_out = np.ascontiguousarray(lat_out, np.float64)
    nlat_in, nlon_in = lon_in.shape[0] - 1, lon_in.shape[1] - 1
    nlat_out, nlon_out = lon_out.shape[0] - 1, lon_out.shape[1] - 1
    mask_in = np.ones(nlon_in * nlat_in) if mask_in is None else np.ascontiguousarray(mask_in, np.float64)
    cap = get_maxxgrid()
    ii, ji, io, jo = (np.empty(cap, np.int32) for _ in range(4))
    xa = np.empty(cap); xc = np.empty(cap); yc = np.empty(cap)
    ci = lambda v: C.byref(C.c_int(v))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = L.create_xgrid_great_circle(ci(nlon_in), ci(nlat_in), ci(nlon_out), ci(nlat_out), p(lon_in), p(lat_in), p(lon_out), p(lat_out),
                                    p(mask_in), p(ii), p(ji), p(io), p(jo), p(xa), p(xc), p(yc))
    return n, ii[:n], ji[:n], io[:n], jo[:n], xa[:n], xc[:n], yc[:n]


def get_grid_great_circle_area(lon, lat):
    """create_xgrid.c:98 — spherical-excess cell areas [ny, nx] in m^2."""
    lon = np.ascontiguousarray(lon, np.float64); lat = np.ascontiguousarray(lat, np.float64)
    ny, nx = lon.shape[0] - 1, lon.shape[1] - 1
    area = np.empty((ny, nx))
    lib().get_grid_great_circle_area(C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), lon.ctypes.data_as(C.c_void_p),
                                     lat.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p))
    return area


def get_grid_area(lon, lat):
    """create_xgrid.c:66 — cell areas [ny, nx] in m^2."""
    lon = np.ascontiguousarray(lon, np.float64); lat = np.ascontiguousarray(lat, np.float64)
    ny, nx = lon.shape[0] - 1, lon.shape[1] - 1
    area = np.empty((ny, nx))
    lib().get_grid_area(C.byref(C.c_int(nx)), C.byref(C.c_int(ny)), lon.ctypes.data_as(C.c_void_p),
                        lat.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p))
    return area


# ---------------------------------------------------------------------------------------------
# make_coupler_mosaic's exchange grids (csrc/coupler.cu, include/xgrid_b200.h Part 5)
# ---------------------------------------------------------------------------------------------
class _MosaicGrid(C.Structure):
    _fields_ = [("ntiles", C.c_int), ("nx", C.c_void_p), ("ny", C.c_void_p), ("lon", C.c_void_p), ("lat", C.c_void_p)]


class _CouplerList(C.Structure):
    _fields_ = [("n", C.c_longlong)] + [(k, C.POINTER(C.c_int)) for k in ("t1", "i1", "j1", "t2", "i2", "j2")] + \
               [(k, C.POINTER(C.c_double)) for k in ("area", "d1i", "d1j", "d2i", "d2j")]


class _CouplerResult(C.Structure):
    _fields_ = [("atmxlnd", _CouplerList), ("atmxocn", _CouplerList), ("lndxocn", _CouplerList),
                ("ncell_atm", C.c_longlong), ("ncell_lnd", C.c_longlong), ("ncell_ocn", C.c_longlong)] + \
               [(k, C.POINTER(C.c_double)) for k in ("area_atm", "area_lnd", "area_ocn", "lnd_xarea", "ocn_xarea")]


def _mosaic_struct(tiles):
    """tiles: list of (lon, lat) vertex arrays [ny+1, nx+1] in radians -> (_MosaicGrid, keepalive)"""
    nx = np.array([t[0].shape[1] - 1 for t in tiles], np.int32)
    ny = np.array([t[0].shape[0] - 1 for t in tiles], np.int32)
    lon = np.concatenate([np.ascontiguousarray(t[0], np.float64).ravel() for t in tiles])
    lat = np.concatenate([np.ascontiguousarray(t[1], np.float64).ravel() for t in tiles])
    g = _MosaicGrid(len(tiles), nx.ctypes.data, ny.ctypes.data, lon.ctypes.data, lat.ctypes.data)
    return g, (nx, ny, lon, lat)


def make_coupler_xgrid(atm, ocn, omask, lnd=None, interp_order=2, area_ratio_thresh=1.0e-6, tile_nest=-1,
                       ocn_same_as_atm=False, device=0):
    """The exchange grids of make_coupler_mosaic (make_coupler_mosaic.c:1250-1720, :2556-2808) on the GPU.
    atm / lnd / ocn: lists of (lon, lat) vertex arrays per tile, radians; lnd=None: the land model runs on the atmosphere mosaic.
    omask: list of ocean-fraction arrays [ny, nx] per ocean tile.  Returns a dict of the three lists (numpy arrays, 0-based
    parent cells) and the per-cell sums behind land_mask / ocean_mask."""
    L = lib()
    L.xgb_make_coupler_xgrid.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    L.xgb_coupler_result_free.argtypes = [C.c_void_p]
    ga, ka = _mosaic_struct(atm)
    go, ko = _mosaic_struct(ocn)
    gl, kl = (_mosaic_struct(lnd) if lnd is not None else (None, None))
    mask = np.concatenate([np.ascontiguousarray(m, np.float64).ravel() for m in omask])
    if mask.size != int(sum((t[0].shape[0] - 1) * (t[0].shape[1] - 1) for t in ocn)):
        raise ValueError("omask does not match the ocean mosaic")
    res = _CouplerResult()
    rc = L.xgb_make_coupler_xgrid(device, interp_order, area_ratio_thresh, tile_nest, 1 if lnd is None else 0,
                                  1 if ocn_same_as_atm else 0, C.addressof(ga), C.addressof(gl) if gl is not None else None,
                                  C.addressof(go), mask.ctypes.data, C.addressof(res))
    if rc != 0:
        raise XgridError(_err())
    try:
        def lst(l):
            n = int(l.n)
            d = {k: np.ctypeslib.as_array(getattr(l, k), (n,)).copy() if n else np.zeros(0, np.int32) for k in ("t1", "i1", "j1", "t2", "i2", "j2")}
            d["area"] = np.ctypeslib.as_array(l.area, (n,)).copy() if n else np.zeros(0)
            if interp_order == 2:
                for k in ("d1i", "d1j", "d2i", "d2j"):
                    d[k] = np.ctypeslib.as_array(getattr(l, k), (n,)).copy() if n else np.zeros(0)
            return d
        out = dict(atmxlnd=lst(res.atmxlnd), atmxocn=lst(res.atmxocn), lndxocn=lst(res.lndxocn))
        for k, n in (("area_atm", res.ncell_atm), ("area_lnd", res.ncell_lnd), ("area_ocn", res.ncell_ocn),
                     ("lnd_xarea", res.ncell_lnd), ("ocn_xarea", res.ncell_ocn)):
            out[k] = np.ctypeslib.as_array(getattr(res, k), (int(n),)).copy()
        return out
    finally:
        L.xgb_coupler_result_free(C.addressof(res))
