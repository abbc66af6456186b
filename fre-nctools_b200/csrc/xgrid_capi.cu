// C-ABI layer of libxgrid_b200 (see include/xgrid_b200.h).
// Plan management, host<->device staging, and the reference-signature entry points.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/xgrid_b200.h"
#include "xgrid_internal.h"
#include "xgrid_plan.h"

using namespace xgb;

static SrcMap single_window(long long begin, long long end);
static SrcMap sub_map(const SrcMap& m, long long tb, long long te);

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void xgb_set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* xgb_last_error(void) { return g_err; }
extern "C" void xgb_set_error_str(const char* s) { xgb_set_error("%s", s); }      // for the plain-C files (remap_file.c)

// The reference's error_handler (mosaic_util.c:57-65): message on stderr, exit(1).
[[noreturn]] static void fatal(const char* msg)
{
  fprintf(stderr, "FATAL Error: %s\n", msg);
  exit(1);
}

#define CU_OK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      xgb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

int DevBuf::reserve(size_t bytes)
{
  if (bytes <= cap) return 0;
  if (p) { cudaFree(p); p = nullptr; cap = 0; }
  // grow geometrically so repeated generates with slowly varying sizes do not re-allocate
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    xgb_set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    p = nullptr; cap = 0;
    return 1;
  }
  cap = want;
  return 0;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

extern "C" int xgb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
extern "C" xgb_plan* xgb_plan_create(int device)
{
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    xgb_set_error("no CUDA device available (%s); libxgrid_b200 has no CPU path", cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) { xgb_set_error("device %d out of range (have %d)", device, ndev); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { xgb_set_error("cudaSetDevice(%d) failed", device); return nullptr; }
  xgb_plan* p = new xgb_plan();
  p->device = device;
  if (cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&p->err_dev, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&p->total_dev, 2 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMallocHost(&p->total_host, 4 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMallocHost(&p->err_host, sizeof(int)) != cudaSuccess ||
      cudaMallocHost(&p->win_host, (kMaxWindows + 1) * sizeof(unsigned)) != cudaSuccess ||
      cudaMallocHost(&p->rect_host, sizeof(int)) != cudaSuccess) {
    xgb_set_error("plan resource allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete p;
    return nullptr;
  }
  for (int k = 0; k < 6; ++k) cudaEventCreate(&p->ev[k]);
  if (cudaStreamCreateWithFlags(&p->aux_st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->join_ev, cudaEventDisableTiming) != cudaSuccess) {
    xgb_set_error("cannot create the auxiliary stream: %s", cudaGetErrorString(cudaGetLastError()));
    xgb_plan_destroy(p);
    return nullptr;
  }
  cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
  return p;
}

extern "C" void xgb_plan_destroy(xgb_plan* p)
{
  if (!p) return;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->st);
  DevBuf* bufs[] = {&p->dst_lon, &p->dst_lat, &p->dst_store, &p->pyr_store, &p->src_lon, &p->src_lat, &p->mask,
                    &p->src_store, &p->tiles_dev, &p->cnt, &p->pair_off, &p->pair_cnt, &p->out_off, &p->pairs, &p->parea,
                    &p->pclon, &p->pclat, &p->scan_tmp, &p->t_in, &p->i_in, &p->j_in, &p->i_out, &p->j_out,
                    &p->area, &p->clon, &p->clat, &p->di, &p->dj, &p->bounds_dev,
                    &p->heavy_ctl, &p->heavy_flag, &p->heavy_list, &p->heavy_items, &p->heavy_pairs,
                    &p->gc_src_xyz, &p->gc_dst_xyz, &p->gc_pyr_store,
                    &p->rect_store, &p->rect_rows, &p->rect_invalid, &p->o2_acc, &p->o2_tmp,
                    &p->clip_vx, &p->clip_vy, &p->clip_meta};
  for (DevBuf* b : bufs) b->release();
  xgb_apply_release(p);
  for (int k = 0; k < 6; ++k) if (p->ev[k]) cudaEventDestroy(p->ev[k]);
  if (p->err_dev) cudaFree(p->err_dev);
  if (p->total_dev) cudaFree(p->total_dev);
  if (p->total_host) cudaFreeHost(p->total_host);
  if (p->err_host) cudaFreeHost(p->err_host);
  if (p->win_host) cudaFreeHost(p->win_host);
  if (p->rect_host) cudaFreeHost(p->rect_host);
  if (p->fork_ev) cudaEventDestroy(p->fork_ev);
  if (p->join_ev) cudaEventDestroy(p->join_ev);
  if (p->aux_st) cudaStreamDestroy(p->aux_st);
  if (p->copy_ev) cudaEventDestroy(p->copy_ev);
  if (p->copy_st) cudaStreamDestroy(p->copy_st);
  if (p->st) cudaStreamDestroy(p->st);
  delete p;
}

extern "C" void* xgb_plan_stream(xgb_plan* p) { return p ? (void*)p->st : nullptr; }

extern "C" int xgb_plan_sync(xgb_plan* p)
{
  if (!p) { xgb_set_error("null plan"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

static int carve_cellset(DevBuf& store, long long ncell, CellSet* cs)
{
  const size_t nd = (size_t)ncell;
  const size_t bytes = nd * sizeof(double) * (6 + 2 * kMaxV) + nd + 64;
  if (store.reserve(bytes)) return 1;
  double* b = (double*)store.p;
  cs->ncell = ncell;
  cs->box = (Box*)b; b += 4 * nd;
  cs->xavg = b; b += nd; cs->area = b; b += nd;
  cs->vx = b; b += nd * kMaxV;
  cs->vy = b; b += nd * kMaxV;
  cs->nv = (unsigned char*)b;
  return 0;
}

static int upload(DevBuf& dst, const double* src, size_t n, int on_device, cudaStream_t st)
{
  if (dst.reserve(n * sizeof(double))) return 1;
  CU_OK(cudaMemcpyAsync(dst.p, src, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  return 0;
}

static int report_kernel_error(xgb_plan* p, int e, bool fatal_like_reference);

int xgb_check_kernel_errors(xgb_plan* p, bool fatal_like_reference)
{
  launch_publish(p->err_host, p->err_dev, 1, p->st);
  CU_OK(cudaStreamSynchronize(p->st));
  CU_OK(cudaGetLastError());
  return report_kernel_error(p, *p->err_host, fatal_like_reference);
}

// e: the error word a kernel left (already on the host)
static int report_kernel_error(xgb_plan* p, int e, bool fatal_like_reference)
{
  if (e == 0) return 0;
  cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
  const char* msg = "internal kernel error";
  if (e & kErrTooManyVertices) msg = "create_xgrid.c: n2_in is greater than MAX_V";
  else if (e & kErrParallelEdges)
    msg = "the line between <x1_0,y1_0> and  <x1_1,y1_1> should not parallel to the line between <x2_0,y2_0> and  <x2_1,y2_1>";
  else if (e & kErrClipOverflow) msg = "clip_2dx2d: clipped polygon has more than MV vertices";
  else if (e & kErrStackOverflow) msg = "candidate search: traversal stack exhausted";
  else if (e & kErrGcNotConvex) msg = "create_xgrid.c(clip_2dx2d_great_circle): grid box is not convex";
  else if (e & kErrGcWalk) msg = "clip_2dx2d_great_circle: polygon walk did not return to the first intersection";
  else if (e & kErrGcNodePool) msg = "getNext: curListPos >= MAXNODELIST";
  else if (e & kErrHeavyOverflow) msg = "candidate search: heavy-cell work buffer exhausted; raise XGB_HEAVY_CAP";
  else if (e & kErrBadIndex) msg = "exchange-grid list names a cell outside its grid (does the remap file belong to these grids?)";
  if (fatal_like_reference) fatal(msg);
  xgb_set_error("%s (kernel error bits 0x%x)", msg, e);
  return 1;
}

static int finish_set_dst(xgb_plan* p, int nx, int ny, bool defer_checks = false);

extern "C" int xgb_plan_set_dst(xgb_plan* p, int nx, int ny, const double* lon, const double* lat, int on_device)
{
  if (!p || nx <= 0 || ny <= 0 || !lon || !lat) { xgb_set_error("xgb_plan_set_dst: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t nv = (size_t)(nx + 1) * (ny + 1);
  const long long nc = (long long)nx * ny;
  if (nc >= (1ll << 31)) { xgb_set_error("destination tile too large for 32-bit cell indices"); return 1; }
  if (upload(p->dst_lon, lon, nv, on_device, p->st) || upload(p->dst_lat, lat, nv, on_device, p->st)) return 1;
  return finish_set_dst(p, nx, ny);
}

// fregrid's --nlon/--nlat output grid (get_output_grid_by_size, fregrid_util.c:588-603) built on the device with the
// reference's arithmetic, so the caller does not upload (nlon+1)*(nlat+1) vertices it never had as data.
extern "C" int xgb_plan_set_dst_latlon(xgb_plan* p, int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend)
{
  if (!p || nlon <= 0 || nlat <= 0 || !(lonend > lonbegin) || !(latend > latbegin)) { xgb_set_error("xgb_plan_set_dst_latlon: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t nv = (size_t)(nlon + 1) * (nlat + 1);
  if ((long long)nlon * nlat >= (1ll << 31)) { xgb_set_error("destination tile too large for 32-bit cell indices"); return 1; }
  if (p->dst_lon.reserve(nv * sizeof(double)) || p->dst_lat.reserve(nv * sizeof(double))) return 1;
  launch_latlon_fill(nlon, nlat, lonbegin, lonend, latbegin, latend, (double*)p->dst_lon.p, (double*)p->dst_lat.p, p->st);
  // no host synchronisation here: a regular lat-lon grid is separable by construction; the check kernel's verdict and any kernel
  // error are collected at the next generate's synchronisation (which repeats the window with the pyramid walk if needed)
  return finish_set_dst(p, nlon, nlat, true);
}

static int finish_set_dst(xgb_plan* p, int nx, int ny, bool defer_checks)
{
  const long long nc = (long long)nx * ny;
  if (carve_cellset(p->dst_store, nc, &p->dst)) return 1;
  p->nx2 = nx; p->ny2 = ny;
  TileDesc td{nx, ny, 0, 0};
  launch_cell_precompute(td, (const double*)p->dst_lon.p, (const double*)p->dst_lat.p, p->dst, p->err_dev, p->st);

  // pyramid: level 0 aliases the cell arrays, upper levels live in pyr_store
  Pyramid& P = p->pyr;
  P.nlev = 1;
  P.lev[0] = PyrLevel{nx, ny, p->dst.box};
  size_t upper = 0;
  {
    int lx = nx, ly = ny;
    while ((long long)lx * ly > 32 && P.nlev < kMaxLevels) {
      lx = (lx + 1) / 2; ly = (ly + 1) / 2;
      upper += (size_t)lx * ly;
      P.lev[P.nlev].nx = lx; P.lev[P.nlev].ny = ly;
      ++P.nlev;
    }
  }
  if (p->pyr_store.reserve(upper * sizeof(Box) + 64)) return 1;
  Box* b = (Box*)p->pyr_store.p;
  for (int l = 1; l < P.nlev; ++l) {
    const size_t n = (size_t)P.lev[l].nx * P.lev[l].ny;
    launch_pyramid_level(P.lev[l - 1], b, P.lev[l].nx, P.lev[l].ny, p->st);
    P.lev[l].box = b;
    b += n;
  }
  // separable tile (regular lat-lon): 1-D row / column boxes for the candidate search; XGB_NO_RECT=1 keeps the pyramid walk
  p->rect.valid = 0;
  const char* no_rect = getenv("XGB_NO_RECT");
  if (!(no_rect && no_rect[0] == '1')) {
    if (p->rect_store.reserve((size_t)(2 * ny + 3 * nx) * sizeof(double) + 64) || p->rect_rows.reserve((size_t)ny + 64) ||
        p->rect_invalid.reserve(64))
      return 1;
    launch_rect_setup(p->dst, nx, ny, (double*)p->rect_store.p, (unsigned char*)p->rect_rows.p, (int*)p->rect_invalid.p, &p->rect, p->st);
    if (defer_checks) {
      p->rect_pending = true;                  // p->rect.valid stays 1 (launch_rect_setup) until the verdict is read
    } else {
      int invalid = 1;
      CU_OK(cudaMemcpyAsync(&invalid, p->rect_invalid.p, sizeof(int), cudaMemcpyDeviceToHost, p->st));
      CU_OK(cudaStreamSynchronize(p->st));
      if (invalid) p->rect.valid = 0;
      p->rect_pending = false;
    }
  }
  p->have_dst = true;
  p->gc_dst_ready = false;
  return defer_checks ? 0 : xgb_check_kernel_errors(p, false);
}

extern "C" int xgb_plan_set_src(xgb_plan* p, int ntiles, const int* nx, const int* ny,
                                const double* lon, const double* lat, const double* mask, int on_device)
{
  if (!p || ntiles <= 0 || !nx || !ny || !lon || !lat) { xgb_set_error("xgb_plan_set_src: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  p->tiles.clear();
  long long coff = 0, voff = 0;
  for (int n = 0; n < ntiles; ++n) {
    if (nx[n] <= 0 || ny[n] <= 0) { xgb_set_error("xgb_plan_set_src: empty tile %d", n); return 1; }
    p->tiles.push_back(TileDesc{nx[n], ny[n], coff, voff});
    coff += (long long)nx[n] * ny[n];
    voff += (long long)(nx[n] + 1) * (ny[n] + 1);
  }
  if (coff >= (1ll << 31)) { xgb_set_error("source mosaic too large for 32-bit cell indices"); return 1; }
  if (upload(p->src_lon, lon, (size_t)voff, on_device, p->st) || upload(p->src_lat, lat, (size_t)voff, on_device, p->st)) return 1;
  p->has_mask = (mask != nullptr);
  if (mask && upload(p->mask, mask, (size_t)coff, on_device, p->st)) return 1;
  if (p->tiles_dev.reserve(sizeof(TileDesc) * ntiles)) return 1;
  CU_OK(cudaMemcpyAsync(p->tiles_dev.p, p->tiles.data(), sizeof(TileDesc) * ntiles, cudaMemcpyHostToDevice, p->st));
  if (carve_cellset(p->src_store, coff, &p->src)) return 1;
  for (int n = 0; n < ntiles; ++n)
    launch_cell_precompute(p->tiles[n], (const double*)p->src_lon.p, (const double*)p->src_lat.p, p->src, p->err_dev, p->st);
  p->s0 = 0; p->ns = coff;
  p->map = single_window(0, coff);
  p->have_src = true;
  p->gc_src_ready = false;
  p->o2_state = 0;
  // the tile table was copied from pageable host memory: make sure it has landed before `tiles` can change
  return xgb_check_kernel_errors(p, false);
}

// The source mosaic for a rank that only ever generates the given windows: the plan keeps the mosaic's full cell index space
// (window bounds, t_in/i_in/j_in and emission order are those of the whole mosaic), but only the vertex rows the windows touch
// are uploaded and only their cells are precomputed — host-to-device traffic and setup time scale with the rank's share.  No
// host synchronisation: kernel errors surface at the next generate.  The windows become the active windows.
extern "C" int xgb_plan_set_src_sharded(xgb_plan* p, int ntiles, const int* nx, const int* ny, const double* lon, const double* lat,
                                        const double* mask, int nwin, const long long* begin, const long long* end)
{
  if (!p || ntiles <= 0 || !nx || !ny || !lon || !lat || nwin <= 0 || nwin > kMaxWindows || !begin || !end) {
    xgb_set_error("xgb_plan_set_src_sharded: bad arguments");
    return 1;
  }
  CU_OK(cudaSetDevice(p->device));
  bool same = (int)p->tiles.size() == ntiles;
  for (int n = 0; same && n < ntiles; ++n) same = p->tiles[n].nx == nx[n] && p->tiles[n].ny == ny[n];
  long long coff = 0, voff = 0;
  if (!same) {
    p->tiles.clear();
    for (int n = 0; n < ntiles; ++n) {
      if (nx[n] <= 0 || ny[n] <= 0) { xgb_set_error("xgb_plan_set_src_sharded: empty tile %d", n); return 1; }
      p->tiles.push_back(TileDesc{nx[n], ny[n], coff, voff});
      coff += (long long)nx[n] * ny[n];
      voff += (long long)(nx[n] + 1) * (ny[n] + 1);
    }
    if (coff >= (1ll << 31)) { xgb_set_error("source mosaic too large for 32-bit cell indices"); return 1; }
    if (p->tiles_dev.reserve(sizeof(TileDesc) * ntiles)) return 1;
    CU_OK(cudaMemcpyAsync(p->tiles_dev.p, p->tiles.data(), sizeof(TileDesc) * ntiles, cudaMemcpyHostToDevice, p->st));
  } else {
    coff = p->tiles.back().cell_off + (long long)p->tiles.back().nx * p->tiles.back().ny;
    voff = p->tiles.back().vert_off + (long long)(p->tiles.back().nx + 1) * (p->tiles.back().ny + 1);
  }
  if (p->src_lon.reserve((size_t)voff * sizeof(double)) || p->src_lat.reserve((size_t)voff * sizeof(double))) return 1;
  p->has_mask = (mask != nullptr);
  if (mask && p->mask.reserve((size_t)coff * sizeof(double))) return 1;
  if (carve_cellset(p->src_store, coff, &p->src)) return 1;
  SrcMap m{};
  m.nwin = nwin; m.cum[0] = 0;
  for (int w = 0; w < nwin; ++w) {
    if (begin[w] < 0 || end[w] < begin[w] || end[w] > coff) { xgb_set_error("xgb_plan_set_src_sharded: bad window %d", w); return 1; }
    m.begin[w] = begin[w];
    m.cum[w + 1] = m.cum[w] + (end[w] - begin[w]);
    // the window's cells tile by tile: vertex rows j0 .. j1 + 1 of the tile, then the cells themselves
    for (int t = 0; t < ntiles; ++t) {
      const TileDesc& td = p->tiles[t];
      const long long tb = td.cell_off, te = td.cell_off + (long long)td.nx * td.ny;
      const long long lo = begin[w] > tb ? begin[w] : tb, hi = end[w] < te ? end[w] : te;
      if (hi <= lo) continue;
      const long long c0 = lo - tb, c1 = hi - tb;
      const long long j0 = c0 / td.nx, j1 = (c1 - 1) / td.nx;
      const size_t v0 = (size_t)td.vert_off + (size_t)j0 * (td.nx + 1), nv = (size_t)(j1 - j0 + 2) * (td.nx + 1);
      CU_OK(cudaMemcpyAsync((double*)p->src_lon.p + v0, lon + v0, nv * sizeof(double), cudaMemcpyHostToDevice, p->st));
      CU_OK(cudaMemcpyAsync((double*)p->src_lat.p + v0, lat + v0, nv * sizeof(double), cudaMemcpyHostToDevice, p->st));
      if (mask) CU_OK(cudaMemcpyAsync((double*)p->mask.p + lo, mask + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyHostToDevice, p->st));
      launch_cell_precompute(td, (const double*)p->src_lon.p, (const double*)p->src_lat.p, p->src, p->err_dev, p->st, c0, c1);
    }
  }
  p->map = m;
  p->s0 = begin[0]; p->ns = m.total();
  p->have_src = true;
  p->gc_src_ready = false;
  p->o2_state = 0;
  return 0;
}

static SrcMap single_window(long long begin, long long end)
{
  SrcMap m{};
  m.nwin = 1; m.begin[0] = begin; m.cum[0] = 0; m.cum[1] = end - begin;
  return m;
}

// the part [tb, te) of a map's window-relative index space as a map of its own
static SrcMap sub_map(const SrcMap& m, long long tb, long long te)
{
  SrcMap r{};
  r.nwin = 0; r.cum[0] = 0;
  for (int w = 0; w < m.nwin; ++w) {
    const long long lo = tb > m.cum[w] ? tb : m.cum[w], hi = te < m.cum[w + 1] ? te : m.cum[w + 1];
    if (hi <= lo) continue;
    r.begin[r.nwin] = m.begin[w] + (lo - m.cum[w]);
    r.cum[r.nwin + 1] = r.cum[r.nwin] + (hi - lo);
    ++r.nwin;
  }
  if (r.nwin == 0) { r.nwin = 1; r.begin[0] = 0; r.cum[1] = 0; }
  return r;
}

extern "C" int xgb_plan_set_src_windows(xgb_plan* p, int nwin, const long long* begin, const long long* end)
{
  if (!p || !p->have_src) { xgb_set_error("xgb_plan_set_src_windows: no source grid"); return 1; }
  if (nwin <= 0 || nwin > kMaxWindows || !begin || !end) { xgb_set_error("xgb_plan_set_src_windows: between 1 and %d windows", kMaxWindows); return 1; }
  SrcMap m{};
  m.nwin = nwin; m.cum[0] = 0;
  for (int w = 0; w < nwin; ++w) {
    if (begin[w] < 0 || end[w] < begin[w] || end[w] > p->src.ncell) { xgb_set_error("xgb_plan_set_src_windows: bad window %d", w); return 1; }
    m.begin[w] = begin[w];
    m.cum[w + 1] = m.cum[w] + (end[w] - begin[w]);
  }
  p->map = m;
  p->s0 = begin[0]; p->ns = m.total();
  return 0;
}

// exchange cells each window of the last generate produced (nwin values)
extern "C" int xgb_plan_window_counts(xgb_plan* p, long long* counts)
{
  if (!p || p->nxgrid < 0 || !counts) { xgb_set_error("xgb_plan_window_counts: no result"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  for (int w = 0; w < p->win_nx_n; ++w) counts[w] = p->win_nx[w];
  return 0;
}

extern "C" int xgb_plan_set_src_window(xgb_plan* p, long long begin, long long end)
{
  if (!p || !p->have_src) { xgb_set_error("xgb_plan_set_src_window: no source grid"); return 1; }
  if (begin < 0 || end < begin || end > p->src.ncell) { xgb_set_error("xgb_plan_set_src_window: bad window"); return 1; }
  p->s0 = begin; p->ns = end - begin;
  p->map = single_window(begin, end);
  return 0;
}

// count pass + scan over [s0, s0+ns); leaves pair_off (ns+1 entries) and the total on the host
static int heavy_work(xgb_plan* p, long long ns, HeavyWork* hw)
{
  // capacity of the heavy-cell work lists; grown automatically when a pass overflows (generate_window), XGB_HEAVY_CAP
  // presets it
  size_t cap = (size_t)1 << 20;
  const size_t quarter = (size_t)(ns + p->dst.ncell) / 4;
  if (quarter > cap) cap = quarter;
  if (p->heavy_cap > cap) cap = p->heavy_cap;
  if (const char* env = getenv("XGB_HEAVY_CAP")) { const size_t v = (size_t)atoll(env); if (v > cap) cap = v; }
  if (p->heavy_ctl.reserve(sizeof(HeavyCtl)) || p->heavy_flag.reserve((size_t)ns + 16) ||
      p->heavy_list.reserve((size_t)(ns + 1) * sizeof(int)) || p->heavy_items.reserve(2 * cap * sizeof(int2)) ||
      p->heavy_pairs.reserve(cap * sizeof(int2)))
    return 1;
  hw->ctl = (HeavyCtl*)p->heavy_ctl.p;
  hw->flag = (unsigned char*)p->heavy_flag.p;
  hw->list = (int*)p->heavy_list.p;
  hw->items[0] = (int2*)p->heavy_items.p;
  hw->items[1] = hw->items[0] + cap;
  hw->pairs = (int2*)p->heavy_pairs.p;
  hw->cap = (unsigned)cap;
  return 0;
}

static int count_candidates(xgb_plan* p, const SrcMap& sm, unsigned long long* total)
{
  const long long ns = sm.total();
  if (p->cnt.reserve((size_t)(ns + 1) * sizeof(uint32_t)) || p->pair_off.reserve((size_t)(ns + 1) * sizeof(uint32_t)) ||
      p->scan_tmp.reserve(scan_tmp_bytes(ns)))
    return 1;
  HeavyWork hw;
  if (heavy_work(p, ns, &hw)) return 1;
  for (int attempt = 0;; ++attempt) {
    launch_candidates_count(p->src, sm, p->has_mask ? (const double*)p->mask.p : nullptr, p->pyr, p->dst,
                            (uint32_t*)p->cnt.p, hw, p->err_dev, p->st);
    launch_exclusive_scan((const uint32_t*)p->cnt.p, (uint32_t*)p->pair_off.p, ns, p->total_dev, p->scan_tmp.p, p->st);
    launch_publish(p->total_host, p->total_dev, 2, p->st);
    launch_publish(p->err_host, p->err_dev, 1, p->st);
    CU_OK(cudaStreamSynchronize(p->st));
    if (*p->err_host != kErrHeavyOverflow || attempt >= 8) break;
    p->heavy_cap = (size_t)hw.cap * 2;                    // work lists too small (coarse source on fine destination)
    cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
    if (heavy_work(p, ns, &hw)) return 1;
  }
  if (report_kernel_error(p, *p->err_host, false)) return 1;      // still overflowing after 8 retries, or any other kernel error
  *total = p->total_host[0];
  if (*total >= (1ull << 32)) { xgb_set_error("more than 2^32 candidate pairs in one window; shard the source cells"); return 1; }
  return 0;
}

extern "C" int xgb_plan_partition_shares(xgb_plan* p, int nparts, const double* share, long long* bounds)
{
  if (!p || !p->have_src || !p->have_dst || nparts <= 0 || !bounds) { xgb_set_error("xgb_plan_partition: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  unsigned long long total = 0;
  const long long nc = p->src.ncell;
  if (count_candidates(p, single_window(0, nc), &total)) return 1;
  if (p->bounds_dev.reserve((size_t)(nparts + 1) * sizeof(long long) + (size_t)(nparts + 1) * sizeof(unsigned long long))) return 1;
  unsigned long long* targets_dev = nullptr;
  std::vector<unsigned long long> targets;
  if (share) {                                   // window k ends where the running pair count reaches its cumulative share
    double sum = 0.0, run = 0.0;
    for (int k = 0; k < nparts; ++k) {
      if (!(share[k] > 0.0) || !(share[k] < 1e300)) { xgb_set_error("xgb_plan_partition_shares: share[%d] is not a positive number", k); return 1; }
      sum += share[k];
    }
    targets.assign((size_t)nparts + 1, 0ull);
    for (int k = 1; k < nparts; ++k) {
      run += share[k - 1];
      targets[k] = (unsigned long long)((double)total * (run / sum));
      if (targets[k] > total) targets[k] = total;
      if (targets[k] < targets[k - 1]) targets[k] = targets[k - 1];
    }
    targets[nparts] = total;
    targets_dev = (unsigned long long*)((long long*)p->bounds_dev.p + nparts + 1);
    CU_OK(cudaMemcpyAsync(targets_dev, targets.data(), targets.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, p->st));
  }
  launch_partition((const uint32_t*)p->pair_off.p, nc, total, nparts, (long long*)p->bounds_dev.p, p->st, targets_dev);
  CU_OK(cudaMemcpyAsync(bounds, p->bounds_dev.p, (size_t)(nparts + 1) * sizeof(long long), cudaMemcpyDeviceToHost, p->st));
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int xgb_plan_partition(xgb_plan* p, int nparts, long long* bounds)
{
  return xgb_plan_partition_shares(p, nparts, nullptr, bounds);
}

// One window [s0, s0+ns) of source cells: candidate search, clip, ordered compaction, order-2 correction.
// Results are written at entries [base, base + n) of the plan's result arrays.  stream_cap == 0: the arrays are
// (re)sized to hold base + n entries (only valid for base == 0); otherwise they were sized to stream_cap entries
// beforehand and must not move (asynchronous copies of earlier windows may be in flight).
// The device-resident form (xgb_plan_generate): ONE host synchronisation per window.  The pair total stays on the device —
// clip and scatter are launched for the capacity of the pair buffer and read the true count there — and the result arrays
// are sized for that capacity (nxgrid <= pairs).  Totals, window offsets and the error word come back through pinned memory
// at the end; a buffer that turned out too small (first call, or a grid change) repeats the window with the sizes just
// learnt.  Split in two so that xgb_plan_generate_async can enqueue a window without waiting for it.
// scratch of the two-kernel clip for a launch over `pairs` candidate pairs
static int reserve_clip_scratch(xgb_plan* p, size_t pairs)
{
  const size_t nv = clip_scratch_vertices(pairs);
  return p->clip_vx.reserve(nv * sizeof(double) + 16) || p->clip_vy.reserve(nv * sizeof(double) + 16) ||
         p->clip_meta.reserve(pairs * sizeof(unsigned short) + 16);
}

static int resident_enqueue(xgb_plan* p, int order, const SrcMap& sm, size_t cap, HeavyWork& hw)
{
  const long long ns = sm.total();
  const double* mask = p->has_mask ? (const double*)p->mask.p : nullptr;
  if (p->cnt.reserve((size_t)(ns + 1) * sizeof(uint32_t)) || p->pair_off.reserve((size_t)(ns + 1) * sizeof(uint32_t)) ||
      p->pair_cnt.reserve((size_t)(ns + 1) * sizeof(uint32_t)) || p->out_off.reserve((size_t)(ns + 1) * sizeof(uint32_t)) ||
      p->scan_tmp.reserve(scan_tmp_bytes(ns)))
    return -1;
  if (heavy_work(p, ns, &hw)) return -1;
  if (p->pairs.reserve(cap * sizeof(int2) + 16) || p->parea.reserve(cap * sizeof(double) + 16)) return -1;
  if (order == 2 && (p->pclon.reserve(cap * sizeof(double) + 16) || p->pclat.reserve(cap * sizeof(double) + 16))) return -1;
  if (reserve_clip_scratch(p, cap)) return -1;
  const size_t ni = cap * sizeof(int) + 16, nd = cap * sizeof(double) + 16;
  if (p->t_in.reserve(ni) || p->i_in.reserve(ni) || p->j_in.reserve(ni) || p->i_out.reserve(ni) || p->j_out.reserve(ni) ||
      p->area.reserve(nd))
    return -1;
  if (order == 2 && (p->clon.reserve(nd) || p->clat.reserve(nd) || p->di.reserve(nd) || p->dj.reserve(nd))) return -1;
  const unsigned long long* npairs_dev = &hw.ctl->total;
  cudaEventRecord(p->ev[0], p->st);
  cudaMemsetAsync(p->cnt.p, 0, (size_t)(ns + 1) * sizeof(uint32_t), p->st);
  launch_candidates_single(p->src, sm, mask, p->pyr, p->rect, p->dst, (uint32_t*)p->pair_off.p, (uint32_t*)p->pair_cnt.p,
                           (int2*)p->pairs.p, cap, (uint32_t*)p->cnt.p, hw, p->err_dev, p->st);
  cudaEventRecord(p->ev[1], p->st);
  cudaMemsetAsync(p->cnt.p, 0, (size_t)(ns + 1) * sizeof(uint32_t), p->st);
  cudaEventRecord(p->ev[2], p->st);
  launch_clip(order, p->src, p->dst, mask, (const int2*)p->pairs.p, cap, npairs_dev, sm,
              (double*)p->parea.p, (double*)p->pclon.p, (double*)p->pclat.p, (uint32_t*)p->cnt.p, p->err_dev, p->st,
              (double*)p->clip_vx.p, (double*)p->clip_vy.p, (unsigned short*)p->clip_meta.p);
  cudaEventRecord(p->ev[3], p->st);
  launch_exclusive_scan((const uint32_t*)p->cnt.p, (uint32_t*)p->out_off.p, ns, p->total_dev + 1, p->scan_tmp.p, p->st);
  cudaEventRecord(p->ev[4], p->st);
  launch_scatter(order, (const int2*)p->pairs.p, cap, (const double*)p->parea.p, (const double*)p->pclon.p,
                 (const double*)p->pclat.p, (const uint32_t*)p->pair_off.p, (const uint32_t*)p->pair_cnt.p, (const uint32_t*)p->out_off.p,
                 (const TileDesc*)p->tiles_dev.p, (int)p->tiles.size(), sm, p->nx2,
                 (int*)p->t_in.p, (int*)p->i_in.p, (int*)p->j_in.p, (int*)p->i_out.p, (int*)p->j_out.p,
                 (double*)p->area.p, (double*)p->clon.p, (double*)p->clat.p, &hw, p->st, p->aux_st, p->fork_ev, p->join_ev, npairs_dev);
  if (order == 2 && p->o2_state == 1)
    launch_order2_accumulate(sm, (const uint32_t*)p->out_off.p, (const double*)p->area.p, (const double*)p->clon.p,
                             (const double*)p->clat.p, (double*)p->o2_acc.p, p->src.ncell, p->st);
  else if (order == 2)
    launch_order2_finalize(p->src, sm, (const uint32_t*)p->out_off.p, (const double*)p->area.p, (const double*)p->clon.p,
                           (const double*)p->clat.p, (double*)p->di.p, (double*)p->dj.p, hw.list, &hw.ctl->nheavy, p->st,
                           p->aux_st, p->fork_ev, p->join_ev);
  cudaEventRecord(p->ev[5], p->st);
  if (sm.nwin > 1) launch_publish_windows(p->win_host, p->out_off.p, sm, p->st);
  launch_publish(p->total_host, &hw.ctl->total, 4, p->st);          // pair total (2 words), nheavy, pairs of the pyramid heavy path
  launch_publish(p->total_host + 2, p->total_dev + 1, 2, p->st);    // exchange cells
  launch_publish(p->err_host, p->err_dev, 1, p->st);
  if (p->rect_pending) launch_publish(p->rect_host, p->rect_invalid.p, 1, p->st);
  return 0;
}

// after the stream has drained: 0 = the window is complete (*nx_out set, bookkeeping done), 1 = repeat with the grown
// buffers (cap / heavy lists updated), -1 = error
static int resident_evaluate(xgb_plan* p, const SrcMap& sm, size_t& cap, const HeavyWork& hw, int attempt, unsigned long long* nx_out)
{
  const int e = *p->err_host;
  if (e) cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
  if (p->rect_pending) {                       // verdict of the separability check of xgb_plan_set_dst_latlon
    p->rect_pending = false;
    if (*p->rect_host) { p->rect.valid = 0; return 1; }          // not separable after all: repeat with the pyramid walk
  }
  if (e == kErrHeavyOverflow && attempt < 8) {
    // the level-synchronous work lists were too small (coarse source on a fine curvilinear destination): grow and repeat
    const unsigned heavy_pairs = ((const unsigned*)p->total_host)[3];
    size_t want = (size_t)hw.cap * 2;
    if ((size_t)heavy_pairs + heavy_pairs / 8 > want) want = (size_t)heavy_pairs + heavy_pairs / 8;
    p->heavy_cap = want;
    return 1;
  }
  const unsigned long long npairs = p->total_host[0];
  if (npairs >= (1ull << 32)) { xgb_set_error("more than 2^32 candidate pairs in one window; shard the source cells"); return -1; }
  if (npairs > cap) {
    if (attempt > 9) { xgb_set_error("candidate search: pair buffers keep overflowing"); return -1; }
    cap = (size_t)npairs + (size_t)npairs / 16 + 1024;
    return 1;
  }
  if (report_kernel_error(p, e, false)) return -1;
  const unsigned long long nx = p->total_host[2];
  {                                  // later calls launch for this capacity: keep it close to what the window needs
    const size_t tight = (size_t)npairs + (size_t)npairs / 64 + 1024;
    p->pairs_cap = tight < cap ? tight : cap;
  }
  p->npairs = npairs;
  p->win_nx_n = sm.nwin;
  for (int w = 0; w < sm.nwin; ++w) {
    const unsigned long long lo = (w == 0) ? 0ull : p->win_host[w], hi = (w + 1 == sm.nwin) ? nx : p->win_host[w + 1];
    p->win_nx[w] = (long long)(hi - lo);
  }
  for (int k = 0; k < 5; ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]);
    p->phase_ms[k] = ms;
    p->phase_ms_sum[k] += ms;
  }
  *nx_out = nx;
  return 0;
}

static long long generate_window_resident(xgb_plan* p, int order, const SrcMap& sm)
{
  size_t cap = p->pairs_cap ? p->pairs_cap : (size_t)sm.total() * 8 + (1u << 20);
  for (int attempt = 0;; ++attempt) {
    HeavyWork hw;
    if (resident_enqueue(p, order, sm, cap, hw)) return -1;
    if (cudaStreamSynchronize(p->st) != cudaSuccess) {
      xgb_set_error("xgrid generation failed: %s", cudaGetErrorString(cudaGetLastError()));
      return -1;
    }
    unsigned long long nx = 0;
    const int r = resident_evaluate(p, sm, cap, hw, attempt, &nx);
    if (r < 0) return -1;
    if (r == 0) return (long long)nx;
  }
}

static long long generate_window(xgb_plan* p, int order, const SrcMap& sm, size_t base, size_t stream_cap)
{
  if (stream_cap == 0 && base == 0) return generate_window_resident(p, order, sm);
  const long long ns = sm.total();
  const double* mask = p->has_mask ? (const double*)p->mask.p : nullptr;
  unsigned long long npairs = 0;
  cudaEventRecord(p->ev[0], p->st);
  if (p->cnt.reserve((size_t)(ns + 1) * sizeof(uint32_t)) || p->pair_off.reserve((size_t)(ns + 1) * sizeof(uint32_t)) ||
      p->pair_cnt.reserve((size_t)(ns + 1) * sizeof(uint32_t)) || p->out_off.reserve((size_t)(ns + 1) * sizeof(uint32_t)) ||
      p->scan_tmp.reserve(scan_tmp_bytes(ns)))
    return -1;
  HeavyWork hw;
  if (heavy_work(p, ns, &hw)) return -1;
  // single-pass candidate search into pair buffers sized from the last generate (first call: 8 pairs per source cell);
  // the kernels drop what does not fit and report the true total, then the buffers grow and the pass is repeated
  size_t cap = p->pairs_cap ? p->pairs_cap : (size_t)ns * 8 + (1u << 20);
  for (int attempt = 0;; ++attempt) {
    if (p->pairs.reserve(cap * sizeof(int2) + 16) || p->parea.reserve(cap * sizeof(double) + 16)) return -1;
    if (order == 2 && (p->pclon.reserve(cap * sizeof(double) + 16) || p->pclat.reserve(cap * sizeof(double) + 16))) return -1;
    cudaMemsetAsync(p->cnt.p, 0, (size_t)(ns + 1) * sizeof(uint32_t), p->st);
    launch_candidates_single(p->src, sm, mask, p->pyr, p->rect, p->dst, (uint32_t*)p->pair_off.p, (uint32_t*)p->pair_cnt.p,
                             (int2*)p->pairs.p, cap, (uint32_t*)p->cnt.p, hw, p->err_dev, p->st);
    launch_publish(p->total_host, &hw.ctl->total, 4, p->st);          // total (2 words), nheavy, npairs of the heavy path
    launch_publish(p->err_host, p->err_dev, 1, p->st);
    if (p->rect_pending) launch_publish(p->rect_host, p->rect_invalid.p, 1, p->st);
    if (cudaStreamSynchronize(p->st) != cudaSuccess) {
      xgb_set_error("candidate search failed: %s", cudaGetErrorString(cudaGetLastError()));
      return -1;
    }
    if (p->rect_pending) {
      p->rect_pending = false;
      if (*p->rect_host) { p->rect.valid = 0; continue; }             // not separable after all: repeat with the pyramid walk
    }
    if (*p->err_host == kErrHeavyOverflow && attempt < 8) {
      // the level-synchronous work lists were too small (coarse source on a fine destination): grow and repeat
      const unsigned heavy_pairs = ((const unsigned*)p->total_host)[3];
      size_t want = (size_t)hw.cap * 2;
      if ((size_t)heavy_pairs + heavy_pairs / 8 > want) want = (size_t)heavy_pairs + heavy_pairs / 8;
      p->heavy_cap = want;
      cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
      if (heavy_work(p, ns, &hw)) return -1;
      continue;
    }
    npairs = p->total_host[0];
    if (npairs >= (1ull << 32)) { xgb_set_error("more than 2^32 candidate pairs in one window; shard the source cells"); return -1; }
    if (npairs <= cap) break;
    if (attempt > 9) { xgb_set_error("candidate search: pair buffers keep overflowing"); return -1; }
    cap = (size_t)npairs + (size_t)npairs / 16 + 1024;
  }
  p->pairs_cap = cap;
  p->npairs = npairs;
  cudaEventRecord(p->ev[1], p->st);
  cudaMemsetAsync(p->cnt.p, 0, (size_t)(ns + 1) * sizeof(uint32_t), p->st);
  cudaEventRecord(p->ev[2], p->st);
  if (reserve_clip_scratch(p, (size_t)npairs)) return -1;
  launch_clip(order, p->src, p->dst, mask, (const int2*)p->pairs.p, npairs, nullptr, sm,
              (double*)p->parea.p, (double*)p->pclon.p, (double*)p->pclat.p, (uint32_t*)p->cnt.p, p->err_dev, p->st,
              (double*)p->clip_vx.p, (double*)p->clip_vy.p, (unsigned short*)p->clip_meta.p);
  cudaEventRecord(p->ev[3], p->st);
  launch_exclusive_scan((const uint32_t*)p->cnt.p, (uint32_t*)p->out_off.p, ns, p->total_dev + 1, p->scan_tmp.p, p->st);
  launch_publish(p->total_host + 1, p->total_dev + 1, 2, p->st);
  if (cudaStreamSynchronize(p->st) != cudaSuccess) {
    xgb_set_error("xgrid generation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  const unsigned long long nx = p->total_host[1];

  if (stream_cap == 0) {
    const size_t ni = (size_t)(base + nx) * sizeof(int) + 16, nd = (size_t)(base + nx) * sizeof(double) + 16;
    if (p->t_in.reserve(ni) || p->i_in.reserve(ni) || p->j_in.reserve(ni) || p->i_out.reserve(ni) || p->j_out.reserve(ni) ||
        p->area.reserve(nd))
      return -1;
    if (order == 2 && (p->clon.reserve(nd) || p->clat.reserve(nd) || p->di.reserve(nd) || p->dj.reserve(nd))) return -1;
  } else if (base + nx > stream_cap) {
    xgb_set_error("The xgrid size is too large for resources: %llu exchange cells do not fit the %zu-entry output buffers",
                  (unsigned long long)(base + nx), stream_cap);
    return -1;
  }

  cudaEventRecord(p->ev[4], p->st);
  launch_scatter(order, (const int2*)p->pairs.p, npairs, (const double*)p->parea.p, (const double*)p->pclon.p,
                 (const double*)p->pclat.p, (const uint32_t*)p->pair_off.p, (const uint32_t*)p->pair_cnt.p, (const uint32_t*)p->out_off.p,
                 (const TileDesc*)p->tiles_dev.p, (int)p->tiles.size(), sm, p->nx2,
                 (int*)p->t_in.p + base, (int*)p->i_in.p + base, (int*)p->j_in.p + base, (int*)p->i_out.p + base, (int*)p->j_out.p + base,
                 (double*)p->area.p + base, (double*)p->clon.p + (order == 2 ? base : 0), (double*)p->clat.p + (order == 2 ? base : 0), &hw, p->st,
                 p->aux_st, p->fork_ev, p->join_ev);
  if (order == 2 && p->o2_state == 1)
    launch_order2_accumulate(sm, (const uint32_t*)p->out_off.p, (const double*)p->area.p + base, (const double*)p->clon.p + base,
                             (const double*)p->clat.p + base, (double*)p->o2_acc.p, p->src.ncell, p->st);
  else if (order == 2)
    launch_order2_finalize(p->src, sm, (const uint32_t*)p->out_off.p, (const double*)p->area.p + base,
                           (const double*)p->clon.p + base, (const double*)p->clat.p + base, (double*)p->di.p + base,
                           (double*)p->dj.p + base, hw.list, &hw.ctl->nheavy, p->st, p->aux_st, p->fork_ev, p->join_ev);
  cudaEventRecord(p->ev[5], p->st);
  if (sm.nwin > 1 && stream_cap == 0) {                       // exchange cells per window, for the callers' global offsets
    launch_publish_windows(p->win_host, p->out_off.p, sm, p->st);
  }
  if (xgb_check_kernel_errors(p, false)) return -1;
  if (stream_cap == 0) {
    p->win_nx_n = sm.nwin;
    for (int w = 0; w < sm.nwin; ++w) {
      const unsigned long long lo = (w == 0) ? 0ull : p->win_host[w], hi = (w + 1 == sm.nwin) ? nx : p->win_host[w + 1];
      p->win_nx[w] = (long long)(hi - lo);
    }
  }
  for (int k = 0; k < 5; ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]);
    p->phase_ms[k] = ms;
    p->phase_ms_sum[k] += ms;
  }
  return (long long)nx;
}

extern "C" long long xgb_plan_generate(xgb_plan* p, unsigned int opcode)
{
  if (!p || !p->have_src || !p->have_dst) { xgb_set_error("xgb_plan_generate: set source and destination grids first"); return -1; }
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2))) {
    xgb_set_error("conserve_interp: interp_method should be CONSERVE_ORDER1 or CONSERVE_ORDER2");   // conserve_interp.c:230
    return -1;
  }
  if (cudaSetDevice(p->device) != cudaSuccess) { xgb_set_error("cudaSetDevice failed"); return -1; }
  if (opcode & XGB_GREAT_CIRCLE) return xgb_generate_great_circle(p, order);
  const long long n = generate_window(p, order, p->map, 0, 0);
  if (n < 0) return -1;
  p->nxgrid = n;
  p->order = order;
  p->generates += 1;
  return p->nxgrid;
}

// xgb_plan_generate without the wait: the window is enqueued on the plan's stream with the buffer sizes the last completed
// generate settled on (so one synchronous call has to come first), the result stays in HBM, per-window counts can be taken
// on the device (xgb_plan_window_counts_device) and further work — the next window, a collective — queued behind it.
// xgb_plan_generate_finish waits, checks what the kernels reported and returns the count; if a buffer turned out too small it
// repeats the window synchronously, so the result is always complete afterwards.
extern "C" int xgb_plan_generate_async(xgb_plan* p, unsigned int opcode)
{
  if (!p || !p->have_src || !p->have_dst) { xgb_set_error("xgb_plan_generate_async: set source and destination grids first"); return 1; }
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)) || (opcode & XGB_GREAT_CIRCLE)) {
    xgb_set_error("xgb_plan_generate_async: needs CONSERVE_ORDER1/2 (no great circle)");
    return 1;
  }
  if (p->pairs_cap == 0) { xgb_set_error("xgb_plan_generate_async: run xgb_plan_generate once first (it sizes the buffers)"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  p->pending_order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  p->pending_map = p->map;
  p->pending_cap = p->pairs_cap;
  if (resident_enqueue(p, p->pending_order, p->pending_map, p->pending_cap, p->pending_hw)) return 1;
  p->pending = true;
  return 0;
}

extern "C" long long xgb_plan_generate_finish(xgb_plan* p)
{
  if (!p || !p->pending) { xgb_set_error("xgb_plan_generate_finish: nothing pending"); return -1; }
  if (cudaSetDevice(p->device) != cudaSuccess) { xgb_set_error("cudaSetDevice failed"); return -1; }
  p->pending = false;
  if (cudaStreamSynchronize(p->st) != cudaSuccess) {
    xgb_set_error("xgrid generation failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  unsigned long long nx = 0;
  size_t cap = p->pending_cap;
  const int r = resident_evaluate(p, p->pending_map, cap, p->pending_hw, 0, &nx);
  if (r < 0) return -1;
  long long n = (long long)nx;
  if (r == 1) {                                  // a buffer was too small: repeat the window the synchronous way
    p->pairs_cap = cap;
    n = generate_window_resident(p, p->pending_order, p->pending_map);
    if (n < 0) return -1;
  }
  p->nxgrid = n;
  p->order = p->pending_order;
  p->generates += 1;
  return n;
}

// exchange cells per window of the window last enqueued, written to a device array on the plan's stream (nwin int64 values):
// the one exchange of the multi-GPU path (all-gather of the counts) then needs no trip through the host
extern "C" int xgb_plan_window_counts_device(xgb_plan* p, long long* counts_dev)
{
  if (!p || !counts_dev || !p->have_src) { xgb_set_error("xgb_plan_window_counts_device: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const SrcMap& sm = p->pending ? p->pending_map : p->map;
  launch_window_counts((const uint32_t*)p->out_off.p, sm, p->total_dev + 1, counts_dev, p->st);
  return 0;
}

// Generate the current window in `nchunks` consecutive pieces of source cells and copy every piece to the caller's
// host arrays while the next one is being computed (second stream): the D2H transfer of the result, which is longer
// than its computation on PCIe gen5, overlaps with it.  Host arrays hold `capacity` entries (any of di, dj,
// xgrid_clon, xgrid_clat may be NULL); pinned memory gives real overlap.  The device copy of the result is complete
// afterwards as well (xgb_plan_result_device, xgb_plan_apply_setup).
extern "C" long long xgb_plan_generate_to_host(xgb_plan* p, unsigned int opcode, int nchunks, long long capacity,
                                               int* t_in, int* i_in, int* j_in, int* i_out, int* j_out,
                                               double* area, double* di, double* dj, double* xgrid_clon, double* xgrid_clat)
{
  if (!p || !p->have_src || !p->have_dst) { xgb_set_error("xgb_plan_generate_to_host: set source and destination grids first"); return -1; }
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)) || (opcode & XGB_GREAT_CIRCLE) || capacity <= 0 || nchunks <= 0) {
    xgb_set_error("xgb_plan_generate_to_host: needs CONSERVE_ORDER1/2 (no great circle), capacity > 0, nchunks > 0");
    return -1;
  }
  if (cudaSetDevice(p->device) != cudaSuccess) { xgb_set_error("cudaSetDevice failed"); return -1; }
  if (!p->copy_st) {
    if (cudaStreamCreateWithFlags(&p->copy_st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&p->copy_ev, cudaEventDisableTiming) != cudaSuccess) {
      xgb_set_error("xgb_plan_generate_to_host: cannot create the copy stream");
      return -1;
    }
  }
  const size_t cap = (size_t)capacity;
  const size_t ni = cap * sizeof(int) + 16, nd = cap * sizeof(double) + 16;
  cudaStreamSynchronize(p->copy_st);
  if (p->t_in.reserve(ni) || p->i_in.reserve(ni) || p->j_in.reserve(ni) || p->i_out.reserve(ni) || p->j_out.reserve(ni) || p->area.reserve(nd))
    return -1;
  if (order == 2 && (p->clon.reserve(nd) || p->clat.reserve(nd) || p->di.reserve(nd) || p->dj.reserve(nd))) return -1;

  const long long wn = p->map.total();
  size_t base = 0;
  for (int c = 0; c < nchunks; ++c) {
    const long long b = wn * c / nchunks, e = wn * (c + 1) / nchunks;
    if (e <= b) continue;
    const long long n = generate_window(p, order, sub_map(p->map, b, e), base, cap);
    if (n < 0) { cudaStreamSynchronize(p->copy_st); return -1; }
    // generate_window returned after synchronising p->st: the piece is complete; ship it
    struct { void* dst; const void* src; size_t esz; } cp[] = {
        {t_in, p->t_in.p, 4}, {i_in, p->i_in.p, 4}, {j_in, p->j_in.p, 4}, {i_out, p->i_out.p, 4}, {j_out, p->j_out.p, 4},
        {area, p->area.p, 8}, {order == 2 ? di : nullptr, p->di.p, 8}, {order == 2 ? dj : nullptr, p->dj.p, 8},
        {order == 2 ? xgrid_clon : nullptr, p->clon.p, 8}, {order == 2 ? xgrid_clat : nullptr, p->clat.p, 8}};
    for (auto& x : cp)
      if (x.dst && n > 0)
        cudaMemcpyAsync((char*)x.dst + base * x.esz, (const char*)x.src + base * x.esz, (size_t)n * x.esz, cudaMemcpyDeviceToHost, p->copy_st);
    base += (size_t)n;
  }
  if (cudaStreamSynchronize(p->copy_st) != cudaSuccess) {
    xgb_set_error("xgb_plan_generate_to_host: result copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  p->nxgrid = (long long)base;
  p->order = order;
  p->generates += 1;
  return p->nxgrid;
}

// ---------------------------------------------------------------------------------------------
// Order 2 with several output tiles (conserve_interp.c:148-227, :319-358).  Between _begin and _end every order-2 generate
// adds its exchange cells to per-source-cell sums that persist across calls (output tiles in call order, list order inside
// a call: the reference's order) and leaves di/dj unset; the caller keeps each tile's lists and raw xgrid_clon/xgrid_clat
// (xgb_plan_result_centroids_host).  _end turns the sums into centroids; _distance then gives one tile's tile1_distance.
// ---------------------------------------------------------------------------------------------
extern "C" int xgb_plan_order2_begin(xgb_plan* p)
{
  if (!p || !p->have_src) { xgb_set_error("xgb_plan_order2_begin: set the source grid first"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t bytes = 3 * (size_t)p->src.ncell * sizeof(double);
  if (p->o2_acc.reserve(bytes)) return 1;
  CU_OK(cudaMemsetAsync(p->o2_acc.p, 0, bytes, p->st));
  p->o2_state = 1;
  return 0;
}

extern "C" int xgb_plan_order2_end(xgb_plan* p)
{
  if (!p || p->o2_state != 1) { xgb_set_error("xgb_plan_order2_end: no accumulation in progress"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  launch_order2_centroids(p->src, (double*)p->o2_acc.p, p->st);
  p->o2_state = 2;
  return xgb_check_kernel_errors(p, false);
}

extern "C" int xgb_plan_order2_distance(xgb_plan* p, long long n, const int* t_in, const int* i_in, const int* j_in,
                                        const double* area, const double* xgrid_clon, const double* xgrid_clat, double* di, double* dj)
{
  if (!p || p->o2_state != 2) { xgb_set_error("xgb_plan_order2_distance: call xgb_plan_order2_end first"); return 1; }
  if (n < 0 || (n > 0 && (!t_in || !i_in || !j_in || !area || !xgrid_clon || !xgrid_clat || !di || !dj))) {
    xgb_set_error("xgb_plan_order2_distance: bad arguments");
    return 1;
  }
  if (n == 0) return 0;
  CU_OK(cudaSetDevice(p->device));
  const size_t k = (size_t)n;
  if (p->o2_tmp.reserve(k * (3 * sizeof(int) + 5 * sizeof(double)) + 64)) return 1;
  double* d_area = (double*)p->o2_tmp.p;
  double *d_clon = d_area + k, *d_clat = d_clon + k, *d_di = d_clat + k, *d_dj = d_di + k;
  int* d_t = (int*)(d_dj + k);
  int *d_i = d_t + k, *d_j = d_i + k;
  CU_OK(cudaMemcpyAsync(d_area, area, k * sizeof(double), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaMemcpyAsync(d_clon, xgrid_clon, k * sizeof(double), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaMemcpyAsync(d_clat, xgrid_clat, k * sizeof(double), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaMemcpyAsync(d_t, t_in, k * sizeof(int), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaMemcpyAsync(d_i, i_in, k * sizeof(int), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaMemcpyAsync(d_j, j_in, k * sizeof(int), cudaMemcpyHostToDevice, p->st));
  launch_order2_distance(n, d_t, d_i, d_j, (const TileDesc*)p->tiles_dev.p, (int)p->tiles.size(), d_area, d_clon, d_clat,
                         (const double*)p->o2_acc.p, p->src.ncell, d_di, d_dj, p->err_dev, p->st);
  CU_OK(cudaMemcpyAsync(di, d_di, k * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  CU_OK(cudaMemcpyAsync(dj, d_dj, k * sizeof(double), cudaMemcpyDeviceToHost, p->st));
  return xgb_check_kernel_errors(p, false);
}

extern "C" void xgb_plan_order2_reset(xgb_plan* p) { if (p) p->o2_state = 0; }

extern "C" long long xgb_plan_last_npairs(xgb_plan* p) { return p ? (long long)p->npairs : -1; }

// phases: 0 candidate count+scan, 1 candidate fill, 2 clip, 3 scan of accepted counts, 4 scatter+finalize
extern "C" int xgb_plan_phase_ms(xgb_plan* p, float* last5, double* sum5, long long* generates)
{
  if (!p) { xgb_set_error("null plan"); return 1; }
  for (int k = 0; k < 5; ++k) { if (last5) last5[k] = p->phase_ms[k]; if (sum5) sum5[k] = p->phase_ms_sum[k]; }
  if (generates) *generates = p->generates;
  return 0;
}

extern "C" void xgb_plan_reset_phase_ms(xgb_plan* p)
{
  if (!p) return;
  for (int k = 0; k < 5; ++k) p->phase_ms_sum[k] = 0.0;
  p->generates = 0;
}

extern "C" long long xgb_kernel_launches(void) { return xgb::g_launches; }

extern "C" int xgb_plan_result_device(xgb_plan* p, xgb_xgrid_view* v)
{
  if (!p || !v || p->nxgrid < 0) { xgb_set_error("xgb_plan_result_device: no result"); return 1; }
  v->nxgrid = p->nxgrid;
  v->t_in = (int*)p->t_in.p; v->i_in = (int*)p->i_in.p; v->j_in = (int*)p->j_in.p;
  v->i_out = (int*)p->i_out.p; v->j_out = (int*)p->j_out.p;
  v->area = (double*)p->area.p;
  const bool o2 = (p->order == 2);
  v->di = o2 ? (double*)p->di.p : nullptr; v->dj = o2 ? (double*)p->dj.p : nullptr;
  v->xgrid_clon = o2 ? (double*)p->clon.p : nullptr; v->xgrid_clat = o2 ? (double*)p->clat.p : nullptr;
  return 0;
}

static int d2h(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
  if (!dst || bytes == 0) return 0;
  CU_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  return 0;
}

extern "C" int xgb_plan_result_host(xgb_plan* p, int* t_in, int* i_in, int* j_in, int* i_out, int* j_out,
                                    double* area, double* di, double* dj)
{
  if (!p || p->nxgrid < 0) { xgb_set_error("xgb_plan_result_host: no result"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t n = (size_t)p->nxgrid;
  if (d2h(t_in, p->t_in.p, n * sizeof(int), p->st) || d2h(i_in, p->i_in.p, n * sizeof(int), p->st) ||
      d2h(j_in, p->j_in.p, n * sizeof(int), p->st) || d2h(i_out, p->i_out.p, n * sizeof(int), p->st) ||
      d2h(j_out, p->j_out.p, n * sizeof(int), p->st) || d2h(area, p->area.p, n * sizeof(double), p->st))
    return 1;
  if (p->order == 2 && (d2h(di, p->di.p, n * sizeof(double), p->st) || d2h(dj, p->dj.p, n * sizeof(double), p->st))) return 1;
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int xgb_plan_result_centroids_host(xgb_plan* p, double* xclon, double* xclat)
{
  if (!p || p->nxgrid < 0 || p->order != 2) { xgb_set_error("xgb_plan_result_centroids_host: no order-2 result"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t n = (size_t)p->nxgrid;
  if (d2h(xclon, p->clon.p, n * sizeof(double), p->st) || d2h(xclat, p->clat.p, n * sizeof(double), p->st)) return 1;
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int xgb_plan_src_area_host(xgb_plan* p, double* area)
{
  if (!p || !p->have_src) { xgb_set_error("no source grid"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  if (d2h(area, p->src.area, (size_t)p->src.ncell * sizeof(double), p->st)) return 1;
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int xgb_plan_dst_area_host(xgb_plan* p, double* area)
{
  if (!p || !p->have_dst) { xgb_set_error("no destination grid"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  if (d2h(area, p->dst.area, (size_t)p->dst.ncell * sizeof(double), p->st)) return 1;
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Part 1: reference-signature entry points.  One process-wide plan on device XGB_DEVICE (default 0),
// mirroring the reference's single-threaded, non-reentrant callers.
// ---------------------------------------------------------------------------------------------
static xgb_plan* default_plan()
{
  static xgb_plan* plan = nullptr;
  if (!plan) {
    const char* env = getenv("XGB_DEVICE");
    plan = xgb_plan_create(env ? atoi(env) : 0);
    if (!plan) fatal(xgb_last_error());
  }
  return plan;
}

#ifndef XGB_MAXXGRID
#define XGB_MAXXGRID 5e6            /* create_xgrid.h:22-28, serial build */
#endif

extern "C" int get_maxxgrid(void) { return (int)XGB_MAXXGRID; }

extern "C" void get_grid_area(const int* nlon, const int* nlat, const double* lon, const double* lat, double* area)
{
  xgb_plan* p = default_plan();
  if (xgb_plan_set_dst(p, *nlon, *nlat, lon, lat, 0) || xgb_plan_dst_area_host(p, area)) fatal(xgb_last_error());
}

static int create_xgrid_2dx2d(int order, const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                              const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                              const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out,
                              double* xgrid_area, double* xgrid_clon, double* xgrid_clat)
{
  xgb_plan* p = default_plan();
  if (xgb_plan_set_dst(p, *nlon_out, *nlat_out, lon_out, lat_out, 0)) fatal(xgb_last_error());
  if (xgb_plan_set_src(p, 1, nlon_in, nlat_in, lon_in, lat_in, mask_in, 0)) fatal(xgb_last_error());
  // kernel-raised conditions that the reference treats as fatal are fatal here too
  const long long n = xgb_plan_generate(p, order == 2 ? XGB_CONSERVE_ORDER2 : XGB_CONSERVE_ORDER1);
  if (n < 0) fatal(xgb_last_error());
  if (n >= (long long)XGB_MAXXGRID)                                     // create_xgrid.c:809-812
    fatal("The xgrid size is too large for resources.\n"
          " nxgrid is greater than MAXXGRID/nthreads; increase MAXXGRID,\n"
          " decrease nthreads, or increase number of MPI ranks.");
  if (xgb_plan_result_host(p, nullptr, i_in, j_in, i_out, j_out, xgrid_area, nullptr, nullptr)) fatal(xgb_last_error());
  if (order == 2 && xgb_plan_result_centroids_host(p, xgrid_clon, xgrid_clat)) fatal(xgb_last_error());
  return (int)n;
}

extern "C" int create_xgrid_2dx2d_order1(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out, double* xgrid_area)
{
  return create_xgrid_2dx2d(1, nlon_in, nlat_in, nlon_out, nlat_out, lon_in, lat_in, lon_out, lat_out, mask_in,
                            i_in, j_in, i_out, j_out, xgrid_area, nullptr, nullptr);
}

extern "C" int create_xgrid_2dx2d_order2(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out,
                                         double* xgrid_area, double* xgrid_clon, double* xgrid_clat)
{
  return create_xgrid_2dx2d(2, nlon_in, nlat_in, nlon_out, nlat_out, lon_in, lat_in, lon_out, lat_out, mask_in,
                            i_in, j_in, i_out, j_out, xgrid_area, xgrid_clon, xgrid_clat);
}

extern "C" int create_xgrid_2dx2d_order1_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2,
                                          double* xa)
{
  return create_xgrid_2dx2d_order1(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa);
}

extern "C" int create_xgrid_2dx2d_order2_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2,
                                          double* xa, double* xc, double* yc)
{
  return create_xgrid_2dx2d_order2(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa, xc, yc);
}

// create_xgrid.h:41 / :78-80 (create_xgrid.c:98-137, :1366-1466)
extern "C" void get_grid_great_circle_area(const int* nlon, const int* nlat, const double* lon, const double* lat, double* area)
{
  xgb_plan* p = default_plan();
  p->have_src = false;                // only the destination side is needed for the areas
  if (xgb_plan_set_dst(p, *nlon, *nlat, lon, lat, 0) || xgb_plan_great_circle_area_host(p, 1, area)) fatal(xgb_last_error());
}

extern "C" int create_xgrid_great_circle(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out,
                                         double* xgrid_area, double* xgrid_clon, double* xgrid_clat)
{
  xgb_plan* p = default_plan();
  if (xgb_plan_set_dst(p, *nlon_out, *nlat_out, lon_out, lat_out, 0)) fatal(xgb_last_error());
  if (xgb_plan_set_src(p, 1, nlon_in, nlat_in, lon_in, lat_in, mask_in, 0)) fatal(xgb_last_error());
  const long long n = xgb_plan_generate(p, XGB_CONSERVE_ORDER1 | XGB_GREAT_CIRCLE);
  if (n < 0) fatal(xgb_last_error());
  if (n > (long long)XGB_MAXXGRID) fatal("nxgrid is greater than MAXXGRID, increase MAXXGRID");      // create_xgrid.c:1450
  if (xgb_plan_result_host(p, nullptr, i_in, j_in, i_out, j_out, xgrid_area, nullptr, nullptr)) fatal(xgb_last_error());
  for (long long k = 0; k < n; ++k) { xgrid_clon[k] = 0; xgrid_clat[k] = 0; }                            // :1444-1445
  return (int)n;
}

extern "C" int create_xgrid_great_circle_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2,
                                          double* xa, double* xc, double* yc)
{
  return create_xgrid_great_circle(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa, xc, yc);
}

// ---------------------------------------------------------------------------------------------
// Self-check hooks for ref_trig.cuh (tests only): the host build and the device build of the same
// source, so tests can pin both against the libm the reference links.
// ---------------------------------------------------------------------------------------------
extern "C" void xgb_ref_trig_host(long long n, const double* x, double* s, double* c, double* ss, double* sc)
{
  for (long long i = 0; i < n; ++i) {
    s[i] = ref_sin(x[i]);
    c[i] = ref_cos(x[i]);
    ref_sincos(x[i], &ss[i], &sc[i]);
  }
}

__global__ void ref_trig_kernel(long long n, const double* __restrict__ x, double* s, double* c, double* ss, double* sc)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  s[i] = ref_sin(x[i]);
  c[i] = ref_cos(x[i]);
  ref_sincos(x[i], &ss[i], &sc[i]);
}

extern "C" int xgb_ref_trig_device(long long n, const double* x, double* s, double* c, double* ss, double* sc)
{
  double* d = nullptr;
  const size_t nb = (size_t)n * sizeof(double);
  CU_OK(cudaMalloc(&d, 5 * nb));
  CU_OK(cudaMemcpy(d, x, nb, cudaMemcpyHostToDevice));
  ref_trig_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n);
  CU_OK(cudaGetLastError());
  CU_OK(cudaMemcpy(s, d + n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(c, d + 2 * n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(ss, d + 3 * n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(sc, d + 4 * n, nb, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

// the single-site variant the clip kernel uses (ref_trig_site / ref_sin_small): s1 = sin-only mode, (ss, sc) = sincos mode,
// sm = ref_sin_small; host build and device build (the device build reads the table from shared memory like the kernel)
extern "C" void xgb_ref_trig_site_host(long long n, const double* x, double* s1, double* ss, double* sc, double* sm)
{
  const double* T = ref_trig_table();
  for (long long i = 0; i < n; ++i) {
    double dummy;
    ref_trig_site(x[i], true, &s1[i], &dummy, T);
    ref_trig_site(x[i], false, &ss[i], &sc[i], T);
    sm[i] = ref_sin_small(x[i]);
  }
}

__global__ void ref_trig_site_kernel(long long n, const double* __restrict__ x, double* s1, double* ss, double* sc, double* sm)
{
  __shared__ __align__(16) double T[440];
  for (int k = threadIdx.x; k < 440; k += blockDim.x) T[k] = ref_trig_table()[k];
  __syncthreads();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dummy;
  ref_trig_site(x[i], true, &s1[i], &dummy, T);
  ref_trig_site(x[i], false, &ss[i], &sc[i], T);
  sm[i] = ref_sin_small(x[i]);
}

extern "C" int xgb_ref_trig_site_device(long long n, const double* x, double* s1, double* ss, double* sc, double* sm)
{
  double* d = nullptr;
  const size_t nb = (size_t)n * sizeof(double);
  CU_OK(cudaMalloc(&d, 5 * nb));
  CU_OK(cudaMemcpy(d, x, nb, cudaMemcpyHostToDevice));
  ref_trig_site_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n);
  CU_OK(cudaGetLastError());
  CU_OK(cudaMemcpy(s1, d + n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(ss, d + 2 * n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(sc, d + 3 * n, nb, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(sm, d + 4 * n, nb, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

// host build of the clip kernel's moments routine (tests only): out = {area, ctrlon, ctrlat}
extern "C" void xgb_poly_moments_site_host(int order, int n, const double* x, const double* y, double clon, double* out)
{
  PolyView pv{x, y, 1};
  out[1] = out[2] = 0.0;
  if (order == 2) poly_moments_site<2>(pv, n, clon, ref_trig_table(), &out[0], &out[1], &out[2]);
  else            poly_moments_site<1>(pv, n, clon, ref_trig_table(), &out[0], &out[1], &out[2]);
}
