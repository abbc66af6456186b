// C-ABI of the conservative apply path (do_scalar_conserve_interp, conserve_interp.c:507-910; grad_c2l,
// gradient_c2l.c:58-118; calc_c2l_grid_info, :368-454).  See include/xgrid_b200.h, Part 2b.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/xgrid_b200.h"
#include "apply_internal.h"
#include "xgrid_plan.h"

using namespace xgb;

#define CU_OK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      xgb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

struct xgb_apply_state {
  // geometry of the field arrays
  std::vector<ApplyTile> tiles;
  DevBuf tiles_dev;
  long long ncell = 0, nhalo = 0, ndst = 0;
  int nx2 = 0, ny2 = 0;
  // exchange-grid list the CSR was built from (plan result, or lists uploaded by xgb_plan_set_xgrid)
  bool own_lists = false;
  DevBuf l_t, l_i, l_j, l_io, l_jo, l_area, l_di, l_dj;
  const int *t_in = nullptr, *i_in = nullptr, *j_in = nullptr, *i_out = nullptr, *j_out = nullptr;
  const double *area = nullptr, *di = nullptr, *dj = nullptr;
  long long nxgrid = -1;
  bool have_csr = false, has_dist = false;
  DevBuf off, cursor, perm, c_cell, c_hidx, c_area, c_di, c_dj, scan_tmp;
  ApplyCsr csr{};
  // gradient metrics
  bool have_metrics = false;
  std::vector<GradTile> gtiles;
  DevBuf gtiles_dev, metrics, centers;
  // staging for host-pointer calls, and scratch
  DevBuf s_data, s_gx, s_gy, s_gmask, s_out, s_xdata, s_fb, s_keys;
  // per-source-cell factors (xgb_plan_apply_options)
  int cell_methods = 0;
  bool target_grid = false, opt_weight = false, opt_farea = false, opt_carea = false, opt_dcarea = false;
  double area_missing = -1.e20;
  DevBuf o_weight, o_carea, o_farea, o_dcarea, o_eff;
};

static xgb_apply_state* state(xgb_plan* p)
{
  if (!p->apply) p->apply = new xgb_apply_state();
  return p->apply;
}

void xgb_apply_release(xgb_plan* p)
{
  if (!p || !p->apply) return;
  xgb_apply_state* a = p->apply;
  DevBuf* bufs[] = {&a->tiles_dev, &a->l_t, &a->l_i, &a->l_j, &a->l_io, &a->l_jo, &a->l_area, &a->l_di, &a->l_dj,
                    &a->off, &a->cursor, &a->perm, &a->c_cell, &a->c_hidx, &a->c_area, &a->c_di, &a->c_dj, &a->scan_tmp,
                    &a->gtiles_dev, &a->metrics, &a->centers, &a->s_data, &a->s_gx, &a->s_gy, &a->s_gmask, &a->s_out,
                    &a->s_xdata, &a->s_fb, &a->s_keys, &a->o_weight, &a->o_carea, &a->o_farea, &a->o_dcarea, &a->o_eff};
  for (DevBuf* b : bufs) b->release();
  delete a;
  p->apply = nullptr;
}

static int kernel_errors(xgb_plan* p)
{
  launch_publish(p->err_host, p->err_dev, 1, p->st);
  CU_OK(cudaStreamSynchronize(p->st));
  CU_OK(cudaGetLastError());
  const int e = *p->err_host;
  if (e == 0) return 0;
  cudaMemsetAsync(p->err_dev, 0, sizeof(int), p->st);
  const char* msg = "internal kernel error";
  if (e & kErrApplyIndex) msg = "conserve_interp: exchange-grid entry outside the source mosaic or the output tile (does the remap file belong to these grids?)";
  else if (e & kErrMonotoneMax) msg = " xdata is greater than f_bar_max ";     // conserve_interp.c:693
  else if (e & kErrMonotoneMin) msg = " xdata is less than f_bar_min ";        // conserve_interp.c:707
  else if (e & kErrAreaMissing) msg = "conserve_interp: data is not missing but area is missing";   // :578, :772
  xgb_set_error("%s (kernel error bits 0x%x)", msg, e);
  return 1;
}

static int set_tiles(xgb_plan* p, xgb_apply_state* a, int ntiles, const int* nx, const int* ny)
{
  a->tiles.clear();
  long long coff = 0, hoff = 0;
  for (int n = 0; n < ntiles; ++n) {
    a->tiles.push_back(ApplyTile{nx[n], ny[n], coff, hoff});
    coff += (long long)nx[n] * ny[n];
    hoff += (long long)(nx[n] + 2) * (ny[n] + 2);
  }
  if (hoff >= (1ll << 31)) { xgb_set_error("source mosaic too large for 32-bit cell indices"); return 1; }
  a->ncell = coff; a->nhalo = hoff;
  if (a->tiles_dev.reserve(sizeof(ApplyTile) * ntiles)) return 1;
  CU_OK(cudaMemcpyAsync(a->tiles_dev.p, a->tiles.data(), sizeof(ApplyTile) * ntiles, cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaStreamSynchronize(p->st));
  a->have_metrics = false;
  return 0;
}

template <class T>
static int stage_in(DevBuf& buf, const T* src, size_t n, int on_device, cudaStream_t st, const T** out)
{
  if (on_device) { *out = src; return 0; }
  if (buf.reserve(n * sizeof(T) + 16)) return 1;
  CU_OK(cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  *out = (const T*)buf.p;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// exchange-grid lists -> destination-major CSR
// ---------------------------------------------------------------------------------------------
static int build_csr(xgb_plan* p, xgb_apply_state* a)
{
  const long long n = a->nxgrid, ndst = a->ndst;
  if (n >= (1ll << 32)) { xgb_set_error("more than 2^32 exchange cells in one apply plan; shard the destination"); return 1; }
  const size_t nn = (size_t)(n > 0 ? n : 1);
  if (a->off.reserve((size_t)(ndst + 1) * 4) || a->cursor.reserve((size_t)(ndst + 1) * 4) || a->perm.reserve(nn * 4) ||
      a->c_cell.reserve(nn * 4) || a->c_hidx.reserve(nn * 4) || a->c_area.reserve(nn * 8) ||
      a->scan_tmp.reserve(scan_tmp_bytes(ndst)))
    return 1;
  if (a->has_dist && (a->c_di.reserve(nn * 8) || a->c_dj.reserve(nn * 8))) return 1;
  CU_OK(cudaMemsetAsync(a->cursor.p, 0, (size_t)(ndst + 1) * 4, p->st));
  launch_dst_count(n, a->i_out, a->j_out, a->nx2, a->ny2, (uint32_t*)a->cursor.p, p->err_dev, p->st);
  launch_exclusive_scan((const uint32_t*)a->cursor.p, (uint32_t*)a->off.p, ndst, p->total_dev, a->scan_tmp.p, p->st);
  CU_OK(cudaMemsetAsync(a->cursor.p, 0, (size_t)(ndst + 1) * 4, p->st));
  launch_dst_fill(n, a->i_out, a->j_out, a->nx2, a->ny2, (const uint32_t*)a->off.p, (uint32_t*)a->cursor.p, (uint32_t*)a->perm.p, p->st);
  a->csr.off = (const uint32_t*)a->off.p;
  a->csr.perm = (const uint32_t*)a->perm.p;
  a->csr.cell = (int*)a->c_cell.p; a->csr.hidx = (int*)a->c_hidx.p; a->csr.area = (double*)a->c_area.p;
  a->csr.di = a->has_dist ? (double*)a->c_di.p : nullptr;
  a->csr.dj = a->has_dist ? (double*)a->c_dj.p : nullptr;
  launch_dst_sort_gather(ndst, a->csr.off, (uint32_t*)a->perm.p, a->t_in, a->i_in, a->j_in, a->area, a->has_dist ? a->di : nullptr,
                         a->has_dist ? a->dj : nullptr, (const ApplyTile*)a->tiles_dev.p, (int)a->tiles.size(), a->csr, p->err_dev, p->st);
  if (kernel_errors(p)) return 1;
  a->have_csr = true;
  return 0;
}

extern "C" int xgb_plan_apply_setup(xgb_plan* p)
{
  if (!p || p->nxgrid < 0 || !p->have_src || !p->have_dst) { xgb_set_error("xgb_plan_apply_setup: generate an exchange grid first"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = state(p);
  std::vector<int> nx, ny;
  for (const TileDesc& t : p->tiles) { nx.push_back(t.nx); ny.push_back(t.ny); }
  if (set_tiles(p, a, (int)nx.size(), nx.data(), ny.data())) return 1;
  a->nx2 = p->nx2; a->ny2 = p->ny2; a->ndst = (long long)p->nx2 * p->ny2;
  a->own_lists = false;
  a->t_in = (const int*)p->t_in.p; a->i_in = (const int*)p->i_in.p; a->j_in = (const int*)p->j_in.p;
  a->i_out = (const int*)p->i_out.p; a->j_out = (const int*)p->j_out.p;
  a->area = (const double*)p->area.p;
  a->has_dist = (p->order == 2);
  a->di = a->has_dist ? (const double*)p->di.p : nullptr;
  a->dj = a->has_dist ? (const double*)p->dj.p : nullptr;
  a->nxgrid = p->nxgrid;
  return build_csr(p, a);
}

extern "C" int xgb_plan_set_xgrid(xgb_plan* p, int ntiles, const int* nx, const int* ny, int nx_out, int ny_out, long long nxgrid,
                                  const int* t_in, const int* i_in, const int* j_in, const int* i_out, const int* j_out,
                                  const double* area, const double* di, const double* dj, int on_device)
{
  if (!p || ntiles <= 0 || !nx || !ny || nx_out <= 0 || ny_out <= 0 || nxgrid < 0 || (nxgrid > 0 && (!t_in || !i_in || !j_in || !i_out || !j_out || !area))) {
    xgb_set_error("xgb_plan_set_xgrid: bad arguments");
    return 1;
  }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = state(p);
  if (set_tiles(p, a, ntiles, nx, ny)) return 1;
  a->nx2 = nx_out; a->ny2 = ny_out; a->ndst = (long long)nx_out * ny_out;
  a->own_lists = true;
  a->has_dist = (di != nullptr && dj != nullptr);
  a->nxgrid = nxgrid;
  const size_t n = (size_t)nxgrid;
  // the lists are always copied: the CSR outlives the caller's buffers
  struct { DevBuf* b; const void* src; size_t bytes; const void** dst; } cp[] = {
      {&a->l_t, t_in, n * 4, (const void**)&a->t_in},     {&a->l_i, i_in, n * 4, (const void**)&a->i_in},
      {&a->l_j, j_in, n * 4, (const void**)&a->j_in},     {&a->l_io, i_out, n * 4, (const void**)&a->i_out},
      {&a->l_jo, j_out, n * 4, (const void**)&a->j_out},  {&a->l_area, area, n * 8, (const void**)&a->area},
      {&a->l_di, di, n * 8, (const void**)&a->di},        {&a->l_dj, dj, n * 8, (const void**)&a->dj}};
  for (auto& c : cp) {
    if (!c.src) { *c.dst = nullptr; continue; }
    if (c.b->reserve(c.bytes + 16)) return 1;
    if (c.bytes) CU_OK(cudaMemcpyAsync(c.b->p, c.src, c.bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->st));
    *c.dst = c.b->p;
  }
  return build_csr(p, a);
}

// ---------------------------------------------------------------------------------------------
// gradient metrics
// ---------------------------------------------------------------------------------------------
// per-tile layout inside a->metrics, reference array sizes (gradient_c2l.c:30-47)
static size_t metric_doubles(int nx, int ny)
{
  const size_t nxp = nx + 1, nyp = ny + 1;
  return (size_t)nx * nyp + nxp * ny + (size_t)nx * ny + 2 * nyp + 2 * nxp + 3 * (size_t)nx * nyp + 3 * nxp * ny + 6 * (size_t)nx * ny;
}

static int carve_metrics(xgb_plan* p, xgb_apply_state* a)
{
  size_t tot = 0;
  for (const ApplyTile& t : a->tiles) tot += metric_doubles(t.nx, t.ny);
  if (a->metrics.reserve(tot * 8 + 64) || a->gtiles_dev.reserve(sizeof(GradTile) * a->tiles.size())) return 1;
  double* b = (double*)a->metrics.p;
  a->gtiles.clear();
  for (const ApplyTile& t : a->tiles) {
    const size_t nx = t.nx, ny = t.ny, nxp = nx + 1, nyp = ny + 1;
    GradTile g{};
    g.nx = t.nx; g.ny = t.ny; g.cell_off = t.cell_off; g.halo_off = t.halo_off;
    g.dx = b; b += nx * nyp;      g.dy = b; b += nxp * ny;      g.area = b; b += nx * ny;
    g.edge_w = b; b += nyp;       g.edge_e = b; b += nyp;       g.edge_s = b; b += nxp;       g.edge_n = b; b += nxp;
    g.en_n = b; b += 3 * nx * nyp; g.en_e = b; b += 3 * nxp * ny; g.vlon = b; b += 3 * nx * ny; g.vlat = b; b += 3 * nx * ny;
    a->gtiles.push_back(g);
  }
  CU_OK(cudaMemcpyAsync(a->gtiles_dev.p, a->gtiles.data(), sizeof(GradTile) * a->gtiles.size(), cudaMemcpyHostToDevice, p->st));
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

extern "C" int xgb_plan_grad_setup(xgb_plan* p, const double* lont, const double* latt, int on_device)
{
  if (!p || !p->have_src || !lont || !latt) { xgb_set_error("xgb_plan_grad_setup: set the source grid first"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = state(p);
  if (a->tiles.size() != p->tiles.size()) {
    std::vector<int> nx, ny;
    for (const TileDesc& t : p->tiles) { nx.push_back(t.nx); ny.push_back(t.ny); }
    if (set_tiles(p, a, (int)nx.size(), nx.data(), ny.data())) return 1;
  }
  if (carve_metrics(p, a)) return 1;
  const double *xt = nullptr, *yt = nullptr;
  if (a->centers.reserve((size_t)a->nhalo * 16 + 32)) return 1;
  if (on_device) { xt = lont; yt = latt; }
  else {
    double* c = (double*)a->centers.p;
    CU_OK(cudaMemcpyAsync(c, lont, (size_t)a->nhalo * 8, cudaMemcpyHostToDevice, p->st));
    CU_OK(cudaMemcpyAsync(c + a->nhalo, latt, (size_t)a->nhalo * 8, cudaMemcpyHostToDevice, p->st));
    xt = c; yt = c + a->nhalo;
  }
  for (size_t n = 0; n < a->tiles.size(); ++n) {
    const GradTile& g = a->gtiles[n];
    const double* xc = (const double*)p->src_lon.p + p->tiles[n].vert_off;
    const double* yc = (const double*)p->src_lat.p + p->tiles[n].vert_off;
    launch_c2l_grid_info(g.nx, g.ny, xt + g.halo_off, yt + g.halo_off, xc, yc, (double*)g.dx, (double*)g.dy, (double*)g.area,
                         (double*)g.edge_w, (double*)g.edge_e, (double*)g.edge_s, (double*)g.edge_n, (double*)g.en_n,
                         (double*)g.en_e, (double*)g.vlon, (double*)g.vlat, p->st);
  }
  CU_OK(cudaStreamSynchronize(p->st));
  CU_OK(cudaGetLastError());
  a->have_metrics = true;
  return 0;
}

extern "C" int xgb_plan_grad_set_metrics(xgb_plan* p, int tile, const double* dx, const double* dy, const double* area,
                                         const double* edge_w, const double* edge_e, const double* edge_s, const double* edge_n,
                                         const double* en_n, const double* en_e, const double* vlon, const double* vlat, int on_device)
{
  if (!p || !p->apply || p->apply->tiles.empty()) { xgb_set_error("xgb_plan_grad_set_metrics: no tile layout (call xgb_plan_apply_setup or xgb_plan_set_xgrid first)"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = p->apply;
  if (tile < 0 || tile >= (int)a->tiles.size()) { xgb_set_error("xgb_plan_grad_set_metrics: tile out of range"); return 1; }
  if (a->gtiles.size() != a->tiles.size() && carve_metrics(p, a)) return 1;
  const GradTile& g = a->gtiles[tile];
  const size_t nx = g.nx, ny = g.ny, nxp = nx + 1, nyp = ny + 1;
  const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  struct { const double* dst; const double* src; size_t n; } cp[] = {
      {g.dx, dx, nx * nyp}, {g.dy, dy, nxp * ny}, {g.area, area, nx * ny}, {g.edge_w, edge_w, nyp}, {g.edge_e, edge_e, nyp},
      {g.edge_s, edge_s, nxp}, {g.edge_n, edge_n, nxp}, {g.en_n, en_n, 3 * nx * nyp}, {g.en_e, en_e, 3 * nxp * ny},
      {g.vlon, vlon, 3 * nx * ny}, {g.vlat, vlat, 3 * nx * ny}};
  for (auto& c : cp) {
    if (!c.src) { xgb_set_error("xgb_plan_grad_set_metrics: null metric array"); return 1; }
    CU_OK(cudaMemcpyAsync((void*)c.dst, c.src, c.n * 8, k, p->st));
  }
  CU_OK(cudaStreamSynchronize(p->st));
  a->have_metrics = true;
  return 0;
}

extern "C" int xgb_plan_grad_get_metrics(xgb_plan* p, int tile, double* dx, double* dy, double* area, double* edge_w, double* edge_e,
                                         double* edge_s, double* edge_n, double* en_n, double* en_e, double* vlon, double* vlat)
{
  if (!p || !p->apply || !p->apply->have_metrics) { xgb_set_error("xgb_plan_grad_get_metrics: no metrics"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = p->apply;
  if (tile < 0 || tile >= (int)a->gtiles.size()) { xgb_set_error("xgb_plan_grad_get_metrics: tile out of range"); return 1; }
  const GradTile& g = a->gtiles[tile];
  const size_t nx = g.nx, ny = g.ny, nxp = nx + 1, nyp = ny + 1;
  struct { double* dst; const double* src; size_t n; } cp[] = {
      {dx, g.dx, nx * nyp}, {dy, g.dy, nxp * ny}, {area, g.area, nx * ny}, {edge_w, g.edge_w, nyp}, {edge_e, g.edge_e, nyp},
      {edge_s, g.edge_s, nxp}, {edge_n, g.edge_n, nxp}, {en_n, g.en_n, 3 * nx * nyp}, {en_e, g.en_e, 3 * nxp * ny},
      {vlon, g.vlon, 3 * nx * ny}, {vlat, g.vlat, 3 * nx * ny}};
  for (auto& c : cp)
    if (c.dst) CU_OK(cudaMemcpyAsync(c.dst, c.src, c.n * 8, cudaMemcpyDeviceToHost, p->st));
  CU_OK(cudaStreamSynchronize(p->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// grad_c2l, apply, fused regrid
// ---------------------------------------------------------------------------------------------
extern "C" int xgb_plan_grad_c2l(xgb_plan* p, int nfields, const double* data, double* grad_x, double* grad_y, int* grad_mask,
                                 int has_missing, double missing, int on_device)
{
  if (!p || !p->apply || !p->apply->have_metrics) { xgb_set_error("xgb_plan_grad_c2l: gradient metrics not set (xgb_plan_grad_setup)"); return 1; }
  if (nfields <= 0 || !data || !grad_x || !grad_y) { xgb_set_error("xgb_plan_grad_c2l: bad arguments"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  xgb_apply_state* a = p->apply;
  const size_t nd = (size_t)nfields * a->nhalo, ng = (size_t)nfields * a->ncell;
  const double* d_data = nullptr;
  if (stage_in(a->s_data, data, nd, on_device, p->st, &d_data)) return 1;
  double *d_gx = grad_x, *d_gy = grad_y;
  int* d_gm = grad_mask;
  if (!on_device) {
    if (a->s_gx.reserve(ng * 8 + 16) || a->s_gy.reserve(ng * 8 + 16) || (grad_mask && a->s_gmask.reserve(ng * 4 + 16))) return 1;
    d_gx = (double*)a->s_gx.p; d_gy = (double*)a->s_gy.p; d_gm = grad_mask ? (int*)a->s_gmask.p : nullptr;
  }
  launch_grad_c2l((const GradTile*)a->gtiles_dev.p, (int)a->gtiles.size(), a->ncell, nfields, d_data, a->nhalo, d_gx, d_gy, d_gm,
                  has_missing != 0, missing, p->st);
  if (!on_device) {
    CU_OK(cudaMemcpyAsync(grad_x, d_gx, ng * 8, cudaMemcpyDeviceToHost, p->st));
    CU_OK(cudaMemcpyAsync(grad_y, d_gy, ng * 8, cudaMemcpyDeviceToHost, p->st));
    if (grad_mask) CU_OK(cudaMemcpyAsync(grad_mask, d_gm, ng * 4, cudaMemcpyDeviceToHost, p->st));
    CU_OK(cudaStreamSynchronize(p->st));
  }
  CU_OK(cudaGetLastError());
  return 0;
}

extern "C" int xgb_plan_apply_options(xgb_plan* p, int cell_methods, const double* weight, const double* src_cell_area,
                                      const double* field_area, double area_missing, int target_grid, const double* dst_cell_area,
                                      int on_device)
{
  if (!p || !p->apply || !p->apply->have_csr) { xgb_set_error("xgb_plan_apply_options: call xgb_plan_apply_setup or xgb_plan_set_xgrid first"); return 1; }
  xgb_apply_state* a = p->apply;
  if (cell_methods != 0 && cell_methods != 1) { xgb_set_error("xgb_plan_apply_options: cell_methods must be 0 (mean) or 1 (sum)"); return 1; }
  if ((cell_methods == 1 || field_area) && !src_cell_area && !p->have_src) { xgb_set_error("xgb_plan_apply_options: source cell areas needed"); return 1; }
  if (target_grid && !dst_cell_area && !p->have_dst) { xgb_set_error("xgb_plan_apply_options: destination cell areas needed"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  auto put = [&](DevBuf& b, const double* src, size_t n, bool* flag) -> int {
    *flag = (src != nullptr);
    if (!src) return 0;
    if (b.reserve(n * 8 + 16)) return 1;
    return cudaMemcpyAsync(b.p, src, n * 8, k, p->st) != cudaSuccess;
  };
  if (put(a->o_weight, weight, (size_t)a->ncell, &a->opt_weight) || put(a->o_farea, field_area, (size_t)a->ncell, &a->opt_farea) ||
      put(a->o_carea, src_cell_area, (size_t)a->ncell, &a->opt_carea) || put(a->o_dcarea, dst_cell_area, (size_t)a->ndst, &a->opt_dcarea)) {
    xgb_set_error("xgb_plan_apply_options: copy failed");
    return 1;
  }
  // cell areas default to the plan's own get_grid_area values (grid_in[].cell_area, fregrid_util.c:363-408)
  if (!a->opt_carea && (cell_methods == 1 || field_area)) {
    if (a->o_carea.reserve((size_t)a->ncell * 8 + 16)) return 1;
    CU_OK(cudaMemcpyAsync(a->o_carea.p, p->src.area, (size_t)a->ncell * 8, cudaMemcpyDeviceToDevice, p->st));
    a->opt_carea = true;
  }
  if (!a->opt_dcarea && target_grid) {
    if (a->o_dcarea.reserve((size_t)a->ndst * 8 + 16)) return 1;
    CU_OK(cudaMemcpyAsync(a->o_dcarea.p, p->dst.area, (size_t)a->ndst * 8, cudaMemcpyDeviceToDevice, p->st));
    a->opt_dcarea = true;
  }
  CU_OK(cudaStreamSynchronize(p->st));
  a->cell_methods = cell_methods;
  a->target_grid = target_grid != 0;
  a->area_missing = area_missing;
  return 0;
}

// the CSR the kernels should use for this call: with weight / sum / cell_measures the per-entry area is replaced by the
// effective area (conserve_interp.c:572-585)
static int effective_csr(xgb_plan* p, xgb_apply_state* a, ApplyCsr* out)
{
  *out = a->csr;
  if (!a->opt_weight && a->cell_methods == 0 && !a->opt_farea) return 0;
  const size_t nn = (size_t)(a->nxgrid > 0 ? a->nxgrid : 1);
  if (a->o_eff.reserve(nn * 8 + 16)) return 1;
  launch_effective_area(a->csr, a->nxgrid, a->opt_weight ? (const double*)a->o_weight.p : nullptr, (const double*)a->o_carea.p,
                        (a->cell_methods == 0 && a->opt_farea) ? (const double*)a->o_farea.p : nullptr, a->cell_methods,
                        (double*)a->o_eff.p, p->st);
  out->area = (double*)a->o_eff.p;
  return 0;
}

static int check_variant_batch(xgb_apply_state* a, int nf, bool has_missing)
{
  // the reference refuses these for nz > 1 (conserve_interp.c:544-546); a batch of field-levels shares one field area
  if (a->opt_farea && a->cell_methods == 0 && nf != 1) { xgb_set_error("conserve_interp: cell_measures should be false when nz > 1"); return 1; }
  (void)has_missing;
  return 0;
}

static void finish_variants(xgb_plan* p, xgb_apply_state* a, int order, int nf, const double* d_data, bool has_missing, double miss, double* d_out)
{
  if (a->opt_farea && a->cell_methods == 0 && has_missing)
    launch_measure_check(a->csr, a->nxgrid, order, d_data, (const double*)a->o_farea.p, miss, a->area_missing, p->err_dev, p->st);
  if (a->target_grid && a->cell_methods == 0)
    launch_target_scale(a->csr, a->ndst, nf, a->opt_farea ? (const double*)a->o_farea.p : nullptr, (const double*)a->o_carea.p,
                        (const double*)a->o_dcarea.p, miss, d_out, p->st);
}

// device-side body shared by xgb_plan_apply and xgb_plan_regrid; all pointers are device pointers
static int apply_device(xgb_plan* p, xgb_apply_state* a, unsigned opcode, int nf, const double* d_data, const double* d_gx,
                        const double* d_gy, const int* d_gm, bool has_missing, double missing, double* d_out)
{
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  const double miss = has_missing ? missing : -1.e20;                       // conserve_interp.c:541-542, MAXVAL :36
  if (check_variant_batch(a, nf, has_missing)) return 1;
  ApplyCsr csr;
  if (effective_csr(p, a, &csr)) return 1;
  const int sum_mode = a->cell_methods;
  if (order == 2 && (opcode & XGB_MONOTONIC)) {
    // the limiter couples all exchange cells of a source cell: one field-level at a time (conserve_interp.c:617-742)
    const size_t nn = (size_t)(a->nxgrid > 0 ? a->nxgrid : 1);
    if (a->s_xdata.reserve(nn * 8) || a->s_fb.reserve((size_t)a->ncell * 16 + 16) || a->s_keys.reserve((size_t)a->ncell * 16 + 16)) return 1;
    double* fb = (double*)a->s_fb.p;
    unsigned long long* keys = (unsigned long long*)a->s_keys.p;
    for (int f = 0; f < nf; ++f) {
      launch_monotone(a->nxgrid, a->t_in, a->i_in, a->j_in, a->di, a->dj, (const ApplyTile*)a->tiles_dev.p, (int)a->tiles.size(), a->ncell,
                      d_data + (size_t)f * a->nhalo, d_gx + (size_t)f * a->ncell, d_gy + (size_t)f * a->ncell, d_gm + (size_t)f * a->ncell,
                      miss, fb, fb + a->ncell, keys, keys + a->ncell, (double*)a->s_xdata.p, p->err_dev, p->st);
      launch_apply(2, true, true, csr, a->ndst, 1, nullptr, 0, nullptr, nullptr, nullptr, a->ncell, (const double*)a->s_xdata.p,
                   a->nxgrid, miss, d_out + (size_t)f * a->ndst, p->st, sum_mode);
    }
    finish_variants(p, a, order, nf, d_data, has_missing, miss, d_out);
    return kernel_errors(p);
  }
  launch_apply(order, has_missing, false, csr, a->ndst, nf, d_data, order == 2 ? a->nhalo : a->ncell, d_gx, d_gy, d_gm, a->ncell,
               nullptr, a->nxgrid, miss, d_out, p->st, sum_mode);
  finish_variants(p, a, order, nf, d_data, has_missing, miss, d_out);
  if (a->opt_farea && has_missing) return kernel_errors(p);
  return 0;
}

extern "C" int xgb_plan_apply(xgb_plan* p, unsigned int opcode, int nfields, const double* data, const double* grad_x,
                              const double* grad_y, const int* grad_mask, int has_missing, double missing, double* out, int on_device)
{
  if (!p || !p->apply || !p->apply->have_csr) { xgb_set_error("xgb_plan_apply: call xgb_plan_apply_setup or xgb_plan_set_xgrid first"); return 1; }
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)) || nfields <= 0 || !data || !out) { xgb_set_error("xgb_plan_apply: bad arguments"); return 1; }
  xgb_apply_state* a = p->apply;
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  const bool mono = order == 2 && (opcode & XGB_MONOTONIC);
  if (order == 2 && (!grad_x || !grad_y || !a->has_dist)) { xgb_set_error("xgb_plan_apply: order 2 needs grad_x, grad_y and an order-2 exchange grid"); return 1; }
  if (order == 2 && (has_missing || mono) && !grad_mask) { xgb_set_error("xgb_plan_apply: grad_mask required with missing values / monotonic"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t nd = (size_t)nfields * (order == 2 ? a->nhalo : a->ncell), ng = (size_t)nfields * a->ncell, no = (size_t)nfields * a->ndst;
  const double *d_data = nullptr, *d_gx = nullptr, *d_gy = nullptr;
  const int* d_gm = nullptr;
  if (stage_in(a->s_data, data, nd, on_device, p->st, &d_data)) return 1;
  if (order == 2) {
    if (stage_in(a->s_gx, grad_x, ng, on_device, p->st, &d_gx) || stage_in(a->s_gy, grad_y, ng, on_device, p->st, &d_gy)) return 1;
    if (grad_mask && stage_in(a->s_gmask, grad_mask, ng, on_device, p->st, &d_gm)) return 1;
  }
  double* d_out = out;
  if (!on_device) { if (a->s_out.reserve(no * 8 + 16)) return 1; d_out = (double*)a->s_out.p; }
  if (apply_device(p, a, opcode, nfields, d_data, d_gx, d_gy, d_gm, has_missing != 0, missing, d_out)) return 1;
  if (!on_device) {
    CU_OK(cudaMemcpyAsync(out, d_out, no * 8, cudaMemcpyDeviceToHost, p->st));
    CU_OK(cudaStreamSynchronize(p->st));
  }
  CU_OK(cudaGetLastError());
  return 0;
}

// get_input_data's gradient step (fregrid_util.c:2186-2219) + do_scalar_conserve_interp in one call:
// gradients and masks never leave HBM.
extern "C" int xgb_plan_regrid(xgb_plan* p, unsigned int opcode, int nfields, const double* data, int has_missing, double missing,
                               double* out, int on_device)
{
  if (!p || !p->apply || !p->apply->have_csr) { xgb_set_error("xgb_plan_regrid: call xgb_plan_apply_setup or xgb_plan_set_xgrid first"); return 1; }
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)) || nfields <= 0 || !data || !out) { xgb_set_error("xgb_plan_regrid: bad arguments"); return 1; }
  xgb_apply_state* a = p->apply;
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  if (order == 2 && (!a->have_metrics || !a->has_dist)) { xgb_set_error("xgb_plan_regrid: order 2 needs xgb_plan_grad_setup and an order-2 exchange grid"); return 1; }
  CU_OK(cudaSetDevice(p->device));
  const size_t nd = (size_t)nfields * (order == 2 ? a->nhalo : a->ncell), ng = (size_t)nfields * a->ncell, no = (size_t)nfields * a->ndst;
  const double* d_data = nullptr;
  if (stage_in(a->s_data, data, nd, on_device, p->st, &d_data)) return 1;
  double* d_out = out;
  if (!on_device) { if (a->s_out.reserve(no * 8 + 16)) return 1; d_out = (double*)a->s_out.p; }
  if (order == 2 && !(opcode & XGB_MONOTONIC)) {
    if (check_variant_batch(a, nfields, has_missing != 0)) return 1;
    ApplyCsr csr;
    if (effective_csr(p, a, &csr)) return 1;
    const double miss = has_missing ? missing : -1.e20;
    static const bool use_packed = getenv("XGB_APPLY_PACKED") && getenv("XGB_APPLY_PACKED")[0] == '1';
    if (!use_packed) {
      // field-transposed path: per source cell the values / gradients of all field-levels side by side; a warp covers 32
      // field-levels of one destination cell (apply_kernels.cu: grad_c2l_rec_kernel, apply_rec_kernel)
      if (a->s_gx.reserve(apply_rec_doubles(a->ncell, nfields, has_missing != 0) * 8 + 64)) return 1;
      launch_regrid_rec((const GradTile*)a->gtiles_dev.p, (int)a->gtiles.size(), a->ncell, nfields, d_data, a->nhalo, (double*)a->s_gx.p,
                        has_missing != 0, missing, csr, a->ndst, miss, a->cell_methods, d_out, p->st, a->nx2);
    } else {
      // packed path (round 1): one 32-byte record (value, grad_x, grad_y, grad_mask) per source cell and field-level
      if (a->s_gx.reserve(ng * 32 + 32)) return 1;
      launch_grad_c2l_packed((const GradTile*)a->gtiles_dev.p, (int)a->gtiles.size(), a->ncell, nfields, d_data, a->nhalo,
                             (double*)a->s_gx.p, has_missing != 0, missing, p->st);
      launch_apply_packed(has_missing != 0, csr, a->ndst, nfields, (const double*)a->s_gx.p, a->ncell, miss, d_out, p->st, a->cell_methods, a->nx2);
    }
    finish_variants(p, a, 2, nfields, d_data, has_missing != 0, miss, d_out);
    if (a->opt_farea && has_missing && kernel_errors(p)) return 1;
  } else {
    double *d_gx = nullptr, *d_gy = nullptr;
    int* d_gm = nullptr;
    if (order == 2) {
      if (a->s_gx.reserve(ng * 8 + 16) || a->s_gy.reserve(ng * 8 + 16) || a->s_gmask.reserve(ng * 4 + 16)) return 1;
      d_gx = (double*)a->s_gx.p; d_gy = (double*)a->s_gy.p; d_gm = (int*)a->s_gmask.p;
      launch_grad_c2l((const GradTile*)a->gtiles_dev.p, (int)a->gtiles.size(), a->ncell, nfields, d_data, a->nhalo, d_gx, d_gy, d_gm,
                      has_missing != 0, missing, p->st);
    }
    if (apply_device(p, a, opcode, nfields, d_data, d_gx, d_gy, d_gm, has_missing != 0, missing, d_out)) return 1;
  }
  if (!on_device) {
    CU_OK(cudaMemcpyAsync(out, d_out, no * 8, cudaMemcpyDeviceToHost, p->st));
    CU_OK(cudaStreamSynchronize(p->st));
  }
  CU_OK(cudaGetLastError());
  return 0;
}

// test hook: SharedDiv (csrc/shared_div.cuh) against the compiler's division on n host pairs; *nbad = quotients that differ
extern "C" int xgb_shared_div_check(long long n, const double* a, const double* b, long long* nbad)
{
  unsigned long long bad = 0;
  if (n <= 0 || !a || !b || !nbad) { xgb_set_error("xgb_shared_div_check: bad arguments"); return 1; }
  if (shared_div_check(n, a, b, &bad)) { xgb_set_error("xgb_shared_div_check: device run failed"); return 1; }
  *nbad = (long long)bad;
  return 0;
}

extern "C" long long xgb_plan_apply_nxgrid(xgb_plan* p) { return (p && p->apply) ? p->apply->nxgrid : -1; }
