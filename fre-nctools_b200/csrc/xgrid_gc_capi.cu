// Great-circle exchange-grid path behind xgb_plan_generate(opcode | XGB_GREAT_CIRCLE):
// create_xgrid_great_circle (reference create_xgrid.c:1366-1466) for all source tiles at once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xgrid_b200.h"
#include "xgrid_plan.h"

using namespace xgb;

int xgb_check_kernel_errors(xgb_plan* p, bool fatal_like_reference);   // xgrid_capi.cu

#define CU_OK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      xgb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

static int carve_gc_cells(DevBuf& store, long long ncell, GcCells* c)
{
  const size_t n = (size_t)ncell;
  if (store.reserve(n * (12 * 8 + sizeof(Box3) + 8) + 64)) return 1;
  char* b = (char*)store.p;
  c->ncell = ncell;
  c->box = (Box3*)b;  b += n * sizeof(Box3);
  c->v = (double*)b;  b += n * 12 * 8;
  c->area = (double*)b;
  return 0;
}

static int gc_prepare(xgb_plan* p)
{
  if (!p->gc_dst_ready) {
    const long long nc = (long long)p->nx2 * p->ny2;
    if (carve_gc_cells(p->gc_dst_xyz, nc, &p->gc_dst)) return 1;
    TileDesc td{p->nx2, p->ny2, 0, 0};
    launch_gc_cell_precompute(td, (const double*)p->dst_lon.p, (const double*)p->dst_lat.p, p->gc_dst, p->err_dev, p->st);
    Pyramid3& P = p->gc_pyr;
    P.nlev = 1;
    P.lev[0] = Pyr3Level{p->nx2, p->ny2, p->gc_dst.box};
    size_t upper = 0;
    int lx = p->nx2, ly = p->ny2;
    while ((long long)lx * ly > 32 && P.nlev < kMaxLevels) {
      lx = (lx + 1) / 2; ly = (ly + 1) / 2;
      upper += (size_t)lx * ly;
      P.lev[P.nlev].nx = lx; P.lev[P.nlev].ny = ly;
      ++P.nlev;
    }
    if (p->gc_pyr_store.reserve(upper * sizeof(Box3) + 64)) return 1;
    Box3* b = (Box3*)p->gc_pyr_store.p;
    for (int l = 1; l < P.nlev; ++l) {
      launch_gc_pyramid_level(P.lev[l - 1], b, P.lev[l].nx, P.lev[l].ny, p->st);
      P.lev[l].box = b;
      b += (size_t)P.lev[l].nx * P.lev[l].ny;
    }
    p->gc_dst_ready = true;
  }
  if (!p->gc_src_ready) {
    if (carve_gc_cells(p->gc_src_xyz, p->src.ncell, &p->gc_src)) return 1;
    for (const TileDesc& t : p->tiles)
      launch_gc_cell_precompute(t, (const double*)p->src_lon.p, (const double*)p->src_lat.p, p->gc_src, p->err_dev, p->st);
    p->gc_src_ready = true;
  }
  return 0;
}

long long xgb_generate_great_circle(xgb_plan* p, int order)
{
  if (order != 1) {                                                       // fregrid.c:763-765
    xgb_set_error("fregrid: when clip_method is 'conserve_great_circle', interp_methos need to be 'conserve_order1', contact developer");
    return -1;
  }
  if (gc_prepare(p)) return -1;
  if (p->map.nwin != 1) { xgb_set_error("great-circle generation takes a single source window"); return -1; }
  const long long s0 = p->s0, ns = p->ns;
  SrcMap sm{};
  sm.nwin = 1; sm.begin[0] = s0; sm.cum[0] = 0; sm.cum[1] = ns;
  const double* mask = p->has_mask ? (const double*)p->mask.p : nullptr;
  if (p->cnt.reserve((size_t)(ns + 1) * 4) || p->pair_off.reserve((size_t)(ns + 1) * 4) || p->pair_cnt.reserve((size_t)(ns + 1) * 4) ||
      p->out_off.reserve((size_t)(ns + 1) * 4) ||
      p->scan_tmp.reserve(scan_tmp_bytes(ns)) || p->clon.reserve(gc_slot_bytes(ns) + 16))       // clon: free on this path (order 1)
    return -1;
  int* slots = (int*)p->clon.p;
  cudaEventRecord(p->ev[0], p->st);
  launch_gc_candidates(false, p->gc_src, p->gc_dst, s0, ns, mask, p->gc_pyr, nullptr, (uint32_t*)p->pair_cnt.p, nullptr, p->err_dev, p->st, slots);
  launch_exclusive_scan((const uint32_t*)p->pair_cnt.p, (uint32_t*)p->pair_off.p, ns, p->total_dev, p->scan_tmp.p, p->st);
  launch_publish(p->total_host, p->total_dev, 2, p->st);
  if (cudaStreamSynchronize(p->st) != cudaSuccess) {
    xgb_set_error("great-circle candidate search failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  const unsigned long long npairs = p->total_host[0];
  if (npairs >= (1ull << 32)) { xgb_set_error("more than 2^32 candidate pairs in one window; shard the source cells"); return -1; }
  p->npairs = npairs;
  cudaEventRecord(p->ev[1], p->st);
  if (p->pairs.reserve((size_t)npairs * sizeof(int2) + 16) || p->parea.reserve((size_t)npairs * 8 + 16)) return -1;
  launch_gc_candidates(true, p->gc_src, p->gc_dst, s0, ns, mask, p->gc_pyr, (const uint32_t*)p->pair_off.p, (uint32_t*)p->pair_cnt.p,
                       (int2*)p->pairs.p, p->err_dev, p->st, slots);
  // second stage: drop the pairs a side separates (pclon / pclat are free on this path: flags, their scan; clip_vx: kept pairs)
  const unsigned long long nbox = npairs;
  unsigned long long nkept = nbox;
  if (p->pclon.reserve((size_t)(nbox + 1) * 4 + 16) || p->pclat.reserve((size_t)(nbox + 1) * 4 + 16) ||
      p->clip_vx.reserve((size_t)nbox * sizeof(int2) + 16) || p->scan_tmp.reserve(scan_tmp_bytes((long long)nbox > ns ? (long long)nbox : ns)))
    return -1;
  launch_gc_filter(p->gc_src, p->gc_dst, (const int2*)p->pairs.p, nbox, s0, (uint32_t*)p->pclon.p, p->st);
  launch_exclusive_scan((const uint32_t*)p->pclon.p, (uint32_t*)p->pclat.p, (long long)nbox, p->total_dev, p->scan_tmp.p, p->st);
  launch_gc_compact((const int2*)p->pairs.p, nbox, (const uint32_t*)p->pclon.p, (const uint32_t*)p->pclat.p, (int2*)p->clip_vx.p, ns,
                    (uint32_t*)p->pair_off.p, (uint32_t*)p->pair_cnt.p, p->st);
  launch_publish(p->total_host, p->total_dev, 2, p->st);
  if (cudaStreamSynchronize(p->st) != cudaSuccess) {
    xgb_set_error("great-circle candidate filter failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  nkept = p->total_host[0];
  p->npairs = nkept;
  const int2* kept = (const int2*)p->clip_vx.p;
  cudaMemsetAsync(p->cnt.p, 0, (size_t)(ns + 1) * 4, p->st);
  cudaEventRecord(p->ev[2], p->st);
  launch_gc_clip(p->gc_src, p->gc_dst, mask, kept, nkept, s0, (double*)p->parea.p, (uint32_t*)p->cnt.p,
                 p->err_dev, p->st);
  cudaEventRecord(p->ev[3], p->st);
  launch_exclusive_scan((const uint32_t*)p->cnt.p, (uint32_t*)p->out_off.p, ns, p->total_dev + 1, p->scan_tmp.p, p->st);
  launch_publish(p->total_host + 1, p->total_dev + 1, 2, p->st);
  if (cudaStreamSynchronize(p->st) != cudaSuccess) {
    xgb_set_error("great-circle clip failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  const unsigned long long nx = p->total_host[1];
  p->nxgrid = (long long)nx;
  p->order = 1;
  p->win_nx_n = 1; p->win_nx[0] = (long long)nx;
  const size_t ni = (size_t)nx * 4 + 16, nd = (size_t)nx * 8 + 16;
  if (p->t_in.reserve(ni) || p->i_in.reserve(ni) || p->j_in.reserve(ni) || p->i_out.reserve(ni) || p->j_out.reserve(ni) || p->area.reserve(nd))
    return -1;
  cudaEventRecord(p->ev[4], p->st);
  launch_scatter(1, kept, nkept, (const double*)p->parea.p, nullptr, nullptr, (const uint32_t*)p->pair_off.p,
                 (const uint32_t*)p->pair_cnt.p, (const uint32_t*)p->out_off.p, (const TileDesc*)p->tiles_dev.p, (int)p->tiles.size(), sm, p->nx2,
                 (int*)p->t_in.p, (int*)p->i_in.p, (int*)p->j_in.p, (int*)p->i_out.p, (int*)p->j_out.p, (double*)p->area.p,
                 nullptr, nullptr, nullptr, p->st, nullptr, nullptr, nullptr);
  cudaEventRecord(p->ev[5], p->st);
  if (xgb_check_kernel_errors(p, false)) return -1;
  for (int k = 0; k < 5; ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p->ev[k], p->ev[k + 1]);
    p->phase_ms[k] = ms;
    p->phase_ms_sum[k] += ms;
  }
  p->generates += 1;
  return p->nxgrid;
}

// which: 0 = source cells (concatenated), 1 = destination cells
extern "C" int xgb_plan_great_circle_area_host(xgb_plan* p, int which, double* area)
{
  if (!p || !area || (which == 0 ? !p->have_src : !p->have_dst)) { xgb_set_error("xgb_plan_great_circle_area_host: grid not set"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  // prepare only the side asked for
  if (which == 0 && !p->gc_src_ready) {
    if (carve_gc_cells(p->gc_src_xyz, p->src.ncell, &p->gc_src)) return 1;
    for (const TileDesc& t : p->tiles)
      launch_gc_cell_precompute(t, (const double*)p->src_lon.p, (const double*)p->src_lat.p, p->gc_src, p->err_dev, p->st);
    p->gc_src_ready = true;
  }
  if (which == 1 && !p->gc_dst_ready) {
    if (!p->have_src) {                              // gc_prepare needs both sides only for generation
      const long long nc = (long long)p->nx2 * p->ny2;
      GcCells tmp;
      if (carve_gc_cells(p->gc_dst_xyz, nc, &tmp)) return 1;
      TileDesc td{p->nx2, p->ny2, 0, 0};
      launch_gc_cell_precompute(td, (const double*)p->dst_lon.p, (const double*)p->dst_lat.p, tmp, p->err_dev, p->st);
      CU_OK(cudaMemcpyAsync(area, tmp.area, (size_t)nc * 8, cudaMemcpyDeviceToHost, p->st));
      CU_OK(cudaStreamSynchronize(p->st));
      return 0;
    }
    if (gc_prepare(p)) return 1;
  }
  const GcCells& c = which == 0 ? p->gc_src : p->gc_dst;
  CU_OK(cudaMemcpyAsync(area, c.area, (size_t)c.ncell * 8, cudaMemcpyDeviceToHost, p->st));
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}
