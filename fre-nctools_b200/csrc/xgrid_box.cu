// Exchange grids between a regular lon-lat BOX grid given by its 1-D cell bounds and a curvilinear 2-D grid:
//   create_xgrid_1dx2d_order1 / _order2   (reference create_xgrid.c:208-398)   the box grid is the input grid
//   create_xgrid_2dx1d_order1 / _order2   (reference create_xgrid.c:413-598)   the box grid is the output grid
// used by runoff_regrid / river_regrid and, through interp.c, by every tool that remaps onto a regular grid.  Reference
// signatures (create_xgrid.h:47-64, Fortran twins with a trailing underscore), host pointers in, caller-allocated host
// arrays out, emission order of the reference: box cells row-major, then the 2-D cells row-major.
//
// Per pair the reference does: latitude reject on the raw corners, fix_lon of the 2-D cell towards the box's mean longitude,
// `clip` (Sutherland-Hodgman against the four box sides, create_xgrid.c:1159-1258), poly_area (* mask), the 1e-6 area-ratio
// test, and for order 2 poly_ctrlon / poly_ctrlat.  fix_lon depends on the box only through the final +-2 pi shift
// (mosaic_util.c:727-729), so the pole handling and the unwrap are done once per 2-D cell (fix_lon_unshifted) and every box
// only applies the shift — the same operations on the same numbers, and a box whose longitude range the shifted cell misses
// is dropped before the clip (the clip itself would return 0 after its first or second stage).
//
//   box_area / cell_area   get_grid_area of both grids (get_grid_area_no_adjust for a one-column box grid, :233-236)
//   box_pairs<FILL>        one thread per 2-D cell: rows of boxes its latitude range meets x all box columns; counts per box,
//                          then (second run) writes the accepted pairs into the box's segment
//   box_emit               one thread per box: orders its segment by 2-D cell index and writes the reference's lists
// FP64, -fmad=false, reference association order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/xgrid_b200.h"
#include "xgrid_internal.h"
#include "xgrid_plan.h"

namespace xgb {

constexpr int kBoxMaxV = 16;      // a fix_lon'd cell has at most 8 vertices; each of the four box sides adds at most one

// clip (create_xgrid.c:1159-1258): polygon against the box [ll_lon, ur_lon] x [ll_lat, ur_lat]; returns the vertex count, -1 when
// a stage would exceed kBoxMaxV vertices
__device__ __forceinline__ int clip_box(const double* lon_in, const double* lat_in, int n_in, double ll_lon, double ll_lat, double ur_lon,
                                        double ur_lat, double* lon_out, double* lat_out)
{
  double x_tmp[kBoxMaxV], y_tmp[kBoxMaxV];
  int i_out = 0, n_out;
  // LEFT
  double x_last = lon_in[n_in - 1], y_last = lat_in[n_in - 1];
  bool inside_last = (x_last >= ll_lon);
  for (int i = 0; i < n_in; ++i) {
    const bool inside = (lon_in[i] >= ll_lon);
    if (inside != inside_last) {
      if (i_out >= kBoxMaxV) return -1;
      x_tmp[i_out] = ll_lon;
      y_tmp[i_out++] = y_last + (ll_lon - x_last) * (lat_in[i] - y_last) / (lon_in[i] - x_last);
    }
    if (inside) {
      if (i_out >= kBoxMaxV) return -1;
      x_tmp[i_out] = lon_in[i]; y_tmp[i_out++] = lat_in[i];
    }
    x_last = lon_in[i]; y_last = lat_in[i]; inside_last = inside;
  }
  if (!(n_out = i_out)) return 0;
  // RIGHT
  x_last = x_tmp[n_out - 1]; y_last = y_tmp[n_out - 1];
  inside_last = (x_last <= ur_lon);
  i_out = 0;
  for (int i = 0; i < n_out; ++i) {
    const bool inside = (x_tmp[i] <= ur_lon);
    if (inside != inside_last) {
      if (i_out >= kBoxMaxV) return -1;
      lon_out[i_out] = ur_lon;
      lat_out[i_out++] = y_last + (ur_lon - x_last) * (y_tmp[i] - y_last) / (x_tmp[i] - x_last);
    }
    if (inside) {
      if (i_out >= kBoxMaxV) return -1;
      lon_out[i_out] = x_tmp[i]; lat_out[i_out++] = y_tmp[i];
    }
    x_last = x_tmp[i]; y_last = y_tmp[i]; inside_last = inside;
  }
  if (!(n_out = i_out)) return 0;
  // BOTTOM
  x_last = lon_out[n_out - 1]; y_last = lat_out[n_out - 1];
  inside_last = (y_last >= ll_lat);
  i_out = 0;
  for (int i = 0; i < n_out; ++i) {
    const bool inside = (lat_out[i] >= ll_lat);
    if (inside != inside_last) {
      if (i_out >= kBoxMaxV) return -1;
      y_tmp[i_out] = ll_lat;
      x_tmp[i_out++] = x_last + (ll_lat - y_last) * (lon_out[i] - x_last) / (lat_out[i] - y_last);
    }
    if (inside) {
      if (i_out >= kBoxMaxV) return -1;
      x_tmp[i_out] = lon_out[i]; y_tmp[i_out++] = lat_out[i];
    }
    x_last = lon_out[i]; y_last = lat_out[i]; inside_last = inside;
  }
  if (!(n_out = i_out)) return 0;
  // TOP
  x_last = x_tmp[n_out - 1]; y_last = y_tmp[n_out - 1];
  inside_last = (y_last <= ur_lat);
  i_out = 0;
  for (int i = 0; i < n_out; ++i) {
    const bool inside = (y_tmp[i] <= ur_lat);
    if (inside != inside_last) {
      if (i_out >= kBoxMaxV) return -1;
      lat_out[i_out] = ur_lat;
      lon_out[i_out++] = x_last + (ur_lat - y_last) * (x_tmp[i] - x_last) / (y_tmp[i] - y_last);
    }
    if (inside) {
      if (i_out >= kBoxMaxV) return -1;
      lon_out[i_out] = x_tmp[i]; lat_out[i_out++] = y_tmp[i];
    }
    x_last = x_tmp[i]; y_last = y_tmp[i]; inside_last = inside;
  }
  return i_out;
}

// poly_area_no_adjust (mosaic_util.c:608-634)
__device__ __forceinline__ double poly_area_no_adjust(const double* x, const double* y, int n)
{
  double area = 0.0;
  for (int i = 0; i < n; ++i) {
    const int ip = (i + 1) % n;
    const double dx = x[ip] - x[i];
    const double lat1 = y[ip], lat2 = y[i];
    if (dx == 0.0) continue;
    if (fabs(lat1 - lat2) < kSmall) area -= dx * ref_sin(0.5 * (lat1 + lat2));
    else area += dx * (ref_cos(lat1) - ref_cos(lat2)) / (lat1 - lat2);
  }
  return area * kRadius * kRadius;
}

// get_grid_area (create_xgrid.c:66-88) of the box grid expanded to 2-D (:224-236), or get_grid_area_no_adjust (:168-188)
__global__ void __launch_bounds__(128)
box_area_kernel(int nxb, int nyb, const double* __restrict__ lonb, const double* __restrict__ latb, int no_adjust, double* __restrict__ area)
{
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= (long long)nxb * nyb) return;
  const int i = (int)(b % nxb), j = (int)(b / nxb);
  double x[kMaxV + 2] = {lonb[i], lonb[i + 1], lonb[i + 1], lonb[i]};
  double y[kMaxV + 2] = {latb[j], latb[j], latb[j + 1], latb[j + 1]};
  if (no_adjust) { area[b] = poly_area_no_adjust(x, y, 4); return; }
  const int n = fix_lon(x, y, 4, kPi);
  PolyView pv{x, y, 1};
  area[b] = poly_area(pv, n < 0 ? 4 : n);
}

// per 2-D cell: get_grid_area, the corner latitudes' range, and the cell after the box-independent part of fix_lon
struct BoxCells {
  long long ncell;
  double* area;        // [ncell]
  double* ymin;        // raw corner latitudes: min, max (the reference rejects a pair when all four are <= ll_lat or >= ur_lat)
  double* ymax;
  double* fx;          // [kMaxV][ncell] unshifted fix_lon'd vertices
  double* fy;
  double* xsum_over_n; // x_sum / nn of fix_lon (mosaic_util.c:727)
  unsigned char* nv;
};

__global__ void __launch_bounds__(128)
box_cell_kernel(int nxc, int nyc, const double* __restrict__ lon, const double* __restrict__ lat, BoxCells c, int* err)
{
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= (long long)nxc * nyc) return;
  const int i = (int)(g % nxc), j = (int)(g / nxc);
  const int nxp = nxc + 1;
  const long long n0 = (long long)j * nxp + i, n3 = (long long)(j + 1) * nxp + i;
  double x[kMaxV + 2], y[kMaxV + 2];
  x[0] = lon[n0]; y[0] = lat[n0]; x[1] = lon[n0 + 1]; y[1] = lat[n0 + 1];
  x[2] = lon[n3 + 1]; y[2] = lat[n3 + 1]; x[3] = lon[n3]; y[3] = lat[n3];
  double ymin = y[0], ymax = y[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) { if (y[k] < ymin) ymin = y[k]; if (y[k] > ymax) ymax = y[k]; }
  c.ymin[g] = ymin; c.ymax[g] = ymax;
  {
    double ax[kMaxV + 2], ay[kMaxV + 2];
    for (int k = 0; k < 4; ++k) { ax[k] = x[k]; ay[k] = y[k]; }
    int n = fix_lon(ax, ay, 4, kPi);
    if (n < 0 || n > kMaxV) { atomicOr(err, kErrTooManyVertices); n = 4; }
    PolyView pv{ax, ay, 1};
    c.area[g] = poly_area(pv, n);
  }
  double sum = 0.0;
  int n = fix_lon_unshifted(x, y, 4, &sum);
  if (n < 0 || n > kMaxV) { atomicOr(err, kErrTooManyVertices); n = 4; }
  c.nv[g] = (unsigned char)n;
  c.xsum_over_n[g] = (n > 0) ? sum / n : 0.0;
  for (int k = 0; k < n; ++k) { c.fx[(long long)k * c.ncell + g] = x[k]; c.fy[(long long)k * c.ncell + g] = y[k]; }
}

struct BoxEntry { int cell; double area, clon, clat; };

// One thread per 2-D cell.  FILL == false: cnt[box] += accepted pairs.  FILL == true: the same pairs again, written to the box's
// segment (off[box] + cursor[box]++), unordered inside the segment.
template <int ORDER, bool FILL>
__global__ void __launch_bounds__(128)
box_pairs_kernel(int nxb, int nyb, const double* __restrict__ lonb, const double* __restrict__ latb, const double* __restrict__ area_box,
                 BoxCells c, const double* __restrict__ mask, int mask_on_box, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off,
                 int* __restrict__ e_cell, double* __restrict__ e_area, double* __restrict__ e_clon, double* __restrict__ e_clat, int* err)
{
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= c.ncell) return;
  double cell_mask = 1.0;
  if (!mask_on_box) { cell_mask = mask[g]; if (!(cell_mask > kMaskThresh)) return; }    // 2dx1d: mask on the 2-D input cells
  const double ymin = c.ymin[g], ymax = c.ymax[g];
  const int n_in = c.nv[g];
  if (n_in == 0) return;
  double bx[kMaxV], by[kMaxV];
  for (int k = 0; k < n_in; ++k) { bx[k] = c.fx[(long long)k * c.ncell + g]; by[k] = c.fy[(long long)k * c.ncell + g]; }
  const double mean0 = c.xsum_over_n[g];
  const double a_cell = c.area[g];
  for (int j = 0; j < nyb; ++j) {
    const double ll_lat = latb[j], ur_lat = latb[j + 1];
    if (ymax <= ll_lat) continue;                                   // all four corners <= ll_lat (create_xgrid.c:258-259)
    if (ymin >= ur_lat) continue;                                   // all four corners >= ur_lat (:260-261)
    for (int i = 0; i < nxb; ++i) {
      const long long b = (long long)j * nxb + i;
      double m = cell_mask;
      if (mask_on_box) { m = mask[b]; if (!(m > kMaskThresh)) continue; }     // 1dx2d: mask on the box input cells (:244)
      const double ll_lon = lonb[i], ur_lon = lonb[i + 1];
      const double dx = mean0 - (ll_lon + ur_lon) / 2;              // mosaic_util.c:727: x_sum/nn - tlon
      double x[kMaxV];
      double xmin = 1e300, xmax = -1e300;
      for (int k = 0; k < n_in; ++k) {
        double v = bx[k];
        if (dx < -kPi) v += kTwoPi; else if (dx > kPi) v -= kTwoPi;
        x[k] = v;
        if (v < xmin) xmin = v;
        if (v > xmax) xmax = v;
      }
      if (xmax < ll_lon || xmin > ur_lon) continue;                 // clip returns 0 after its LEFT / RIGHT stage
      double ox[kBoxMaxV], oy[kBoxMaxV];
      const int n_out = clip_box(x, by, n_in, ll_lon, ll_lat, ur_lon, ur_lat, ox, oy);
      if (n_out < 0) { atomicOr(err, kErrClipOverflow); continue; }
      if (n_out == 0) continue;
      PolyView pv{ox, oy, 1};
      const double xarea = poly_area(pv, n_out) * m;
      const double a_box = area_box[b];
      const double min_area = (a_box < a_cell) ? a_box : a_cell;
      if (!(xarea / min_area > kAreaRatioThresh)) continue;
      if (!FILL) { atomicAdd(&cnt[b], 1u); continue; }
      const uint32_t slot = off[b] + atomicAdd(&cnt[b], 1u);
      e_cell[slot] = (int)g;
      e_area[slot] = xarea;
      if (ORDER == 2) {
        double s = 0.0;
        for (int k = 0; k < n_in; ++k) s += x[k];
        e_clon[slot] = poly_ctrlon(pv, n_out, s / n_in);            // lon_in_avg = avgval_double(n_in, x_in) (:343)
        e_clat[slot] = poly_ctrlat(pv, n_out);
      }
    }
  }
}

// one thread per box: its segment ordered by 2-D cell index (the reference's inner loops run over the 2-D cells row-major),
// written out as the reference's lists
__global__ void __launch_bounds__(128)
box_emit_kernel(long long nbox, int nxb, int nxc, const uint32_t* __restrict__ off, int* __restrict__ e_cell, double* __restrict__ e_area,
                double* __restrict__ e_clon, double* __restrict__ e_clat, int order, int box_is_in,
                int* __restrict__ i_in, int* __restrict__ j_in, int* __restrict__ i_out, int* __restrict__ j_out,
                double* __restrict__ xarea, double* __restrict__ xclon, double* __restrict__ xclat)
{
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= nbox) return;
  const uint32_t lo = off[b], hi = off[b + 1];
  for (uint32_t q = lo + 1; q < hi; ++q) {                          // insertion sort by cell index, payload moved along
    const int cc = e_cell[q];
    const double a = e_area[q], u = (order == 2) ? e_clon[q] : 0.0, v = (order == 2) ? e_clat[q] : 0.0;
    uint32_t r = q;
    while (r > lo && e_cell[r - 1] > cc) {
      e_cell[r] = e_cell[r - 1]; e_area[r] = e_area[r - 1];
      if (order == 2) { e_clon[r] = e_clon[r - 1]; e_clat[r] = e_clat[r - 1]; }
      --r;
    }
    e_cell[r] = cc; e_area[r] = a;
    if (order == 2) { e_clon[r] = u; e_clat[r] = v; }
  }
  const int ib = (int)(b % nxb), jb = (int)(b / nxb);
  for (uint32_t q = lo; q < hi; ++q) {
    const int cc = e_cell[q];
    const int ic = cc % nxc, jc = cc / nxc;
    if (box_is_in) { i_in[q] = ib; j_in[q] = jb; i_out[q] = ic; j_out[q] = jc; }
    else           { i_in[q] = ic; j_in[q] = jc; i_out[q] = ib; j_out[q] = jb; }
    xarea[q] = e_area[q];
    if (order == 2) { xclon[q] = e_clon[q]; xclat[q] = e_clat[q]; }
  }
}

}  // namespace xgb

using namespace xgb;

[[noreturn]] static void box_fatal(const char* msg)
{
  fprintf(stderr, "FATAL Error: %s\n", msg);          // error_handler, mosaic_util.c:57-65
  exit(1);
}

#define BOX_CU(call)                                                                                     \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      char m_[256];                                                                                      \
      snprintf(m_, sizeof(m_), "%s failed: %s (xgrid_box.cu:%d); libxgrid_b200 has no CPU path", #call, cudaGetErrorString(e_), __LINE__); \
      box_fatal(m_);                                                                                     \
    }                                                                                                    \
  } while (0)

// box_is_in: 1 = create_xgrid_1dx2d (the 1-D grid is the input grid), 0 = create_xgrid_2dx1d
static int create_xgrid_box(int box_is_in, int order, int nxb, int nyb, const double* lonb, const double* latb, int nxc, int nyc,
                            const double* lon2d, const double* lat2d, const double* mask_in, int* i_in, int* j_in, int* i_out,
                            int* j_out, double* xgrid_area, double* xgrid_clon, double* xgrid_clat)
{
  if (nxb <= 0 || nyb <= 0 || nxc <= 0 || nyc <= 0) return 0;
  const char* env = getenv("XGB_DEVICE");
  BOX_CU(cudaSetDevice(env ? atoi(env) : 0));
  const long long nbox = (long long)nxb * nyb, ncell = (long long)nxc * nyc;
  const size_t nvc = (size_t)(nxc + 1) * (nyc + 1);
  cudaStream_t st;
  BOX_CU(cudaStreamCreate(&st));
  auto dmalloc = [&](size_t bytes) { void* p = nullptr; BOX_CU(cudaMalloc(&p, bytes ? bytes : 8)); return p; };
  double* d_lonb = (double*)dmalloc((nxb + 1) * sizeof(double));
  double* d_latb = (double*)dmalloc((nyb + 1) * sizeof(double));
  double* d_lon = (double*)dmalloc(nvc * sizeof(double));
  double* d_lat = (double*)dmalloc(nvc * sizeof(double));
  const long long nmask = box_is_in ? nbox : ncell;
  double* d_mask = (double*)dmalloc((size_t)nmask * sizeof(double));
  double* d_abox = (double*)dmalloc((size_t)nbox * sizeof(double));
  double* d_cells = (double*)dmalloc((size_t)ncell * sizeof(double) * (4 + 2 * kMaxV) + (size_t)ncell + 64);
  uint32_t* d_cnt = (uint32_t*)dmalloc((size_t)(nbox + 1) * sizeof(uint32_t));
  uint32_t* d_off = (uint32_t*)dmalloc((size_t)(nbox + 1) * sizeof(uint32_t));
  unsigned long long* d_total = (unsigned long long*)dmalloc(sizeof(unsigned long long));
  void* d_scan = dmalloc(scan_tmp_bytes(nbox));
  int* d_err = (int*)dmalloc(sizeof(int));
  BOX_CU(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  BOX_CU(cudaMemcpyAsync(d_lonb, lonb, (nxb + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  BOX_CU(cudaMemcpyAsync(d_latb, latb, (nyb + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
  BOX_CU(cudaMemcpyAsync(d_lon, lon2d, nvc * sizeof(double), cudaMemcpyHostToDevice, st));
  BOX_CU(cudaMemcpyAsync(d_lat, lat2d, nvc * sizeof(double), cudaMemcpyHostToDevice, st));
  BOX_CU(cudaMemcpyAsync(d_mask, mask_in, (size_t)nmask * sizeof(double), cudaMemcpyHostToDevice, st));
  BoxCells c{};
  c.ncell = ncell;
  {
    double* b = d_cells;
    c.area = b; b += ncell; c.ymin = b; b += ncell; c.ymax = b; b += ncell; c.xsum_over_n = b; b += ncell;
    c.fx = b; b += ncell * kMaxV; c.fy = b; b += ncell * kMaxV;
    c.nv = (unsigned char*)b;
  }
  // the temporary fix of create_xgrid_1dx2d_order1 (:233-236): a one-column input grid takes get_grid_area_no_adjust
  const int no_adjust = (box_is_in && order == 1 && nxb == 1) ? 1 : 0;
  xgb::g_launches += 2;
  box_area_kernel<<<(unsigned)((nbox + 127) / 128), 128, 0, st>>>(nxb, nyb, d_lonb, d_latb, no_adjust, d_abox);
  box_cell_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(nxc, nyc, d_lon, d_lat, c, d_err);
  BOX_CU(cudaMemsetAsync(d_cnt, 0, (size_t)(nbox + 1) * sizeof(uint32_t), st));
  const unsigned cblocks = (unsigned)((ncell + 127) / 128);
  ++xgb::g_launches;
  if (order == 2) box_pairs_kernel<2, false><<<cblocks, 128, 0, st>>>(nxb, nyb, d_lonb, d_latb, d_abox, c, d_mask, box_is_in, d_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, d_err);
  else            box_pairs_kernel<1, false><<<cblocks, 128, 0, st>>>(nxb, nyb, d_lonb, d_latb, d_abox, c, d_mask, box_is_in, d_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, d_err);
  launch_exclusive_scan(d_cnt, d_off, nbox, d_total, d_scan, st);
  unsigned long long total = 0;
  int err = 0;
  BOX_CU(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, st));
  BOX_CU(cudaMemcpyAsync(&err, d_err, sizeof(err), cudaMemcpyDeviceToHost, st));
  BOX_CU(cudaStreamSynchronize(st));
  if (err & kErrTooManyVertices) box_fatal("create_xgrid.c: n2_in is greater than MAX_V");
  if (err & kErrClipOverflow) box_fatal("clip: clipped polygon has more than MV vertices");
  if (total > (unsigned long long)get_maxxgrid()) box_fatal("nxgrid is greater than MAXXGRID, increase MAXXGRID");   // :283, :362
  const size_t n = (size_t)total;
  if (n > 0) {
    int* d_ecell = (int*)dmalloc(n * sizeof(int));
    double* d_ea = (double*)dmalloc(n * sizeof(double) * 3);
    int* d_idx = (int*)dmalloc(n * sizeof(int) * 4);
    double* d_out = (double*)dmalloc(n * sizeof(double) * 3);
    BOX_CU(cudaMemsetAsync(d_cnt, 0, (size_t)(nbox + 1) * sizeof(uint32_t), st));
    xgb::g_launches += 2;
    if (order == 2) box_pairs_kernel<2, true><<<cblocks, 128, 0, st>>>(nxb, nyb, d_lonb, d_latb, d_abox, c, d_mask, box_is_in, d_cnt, d_off, d_ecell, d_ea, d_ea + n, d_ea + 2 * n, d_err);
    else            box_pairs_kernel<1, true><<<cblocks, 128, 0, st>>>(nxb, nyb, d_lonb, d_latb, d_abox, c, d_mask, box_is_in, d_cnt, d_off, d_ecell, d_ea, d_ea + n, d_ea + 2 * n, d_err);
    box_emit_kernel<<<(unsigned)((nbox + 127) / 128), 128, 0, st>>>(nbox, nxb, nxc, d_off, d_ecell, d_ea, d_ea + n, d_ea + 2 * n, order, box_is_in,
                                                                    d_idx, d_idx + n, d_idx + 2 * n, d_idx + 3 * n, d_out, d_out + n, d_out + 2 * n);
    BOX_CU(cudaMemcpyAsync(i_in, d_idx, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    BOX_CU(cudaMemcpyAsync(j_in, d_idx + n, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    BOX_CU(cudaMemcpyAsync(i_out, d_idx + 2 * n, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    BOX_CU(cudaMemcpyAsync(j_out, d_idx + 3 * n, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    BOX_CU(cudaMemcpyAsync(xgrid_area, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (order == 2) {
      BOX_CU(cudaMemcpyAsync(xgrid_clon, d_out + n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
      BOX_CU(cudaMemcpyAsync(xgrid_clat, d_out + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    BOX_CU(cudaStreamSynchronize(st));
    BOX_CU(cudaGetLastError());
    cudaFree(d_ecell); cudaFree(d_ea); cudaFree(d_idx); cudaFree(d_out);
  }
  cudaFree(d_lonb); cudaFree(d_latb); cudaFree(d_lon); cudaFree(d_lat); cudaFree(d_mask); cudaFree(d_abox); cudaFree(d_cells);
  cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_total); cudaFree(d_scan); cudaFree(d_err);
  cudaStreamDestroy(st);
  return (int)n;
}

extern "C" int create_xgrid_1dx2d_order1(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out, double* xgrid_area)
{
  return create_xgrid_box(1, 1, *nlon_in, *nlat_in, lon_in, lat_in, *nlon_out, *nlat_out, lon_out, lat_out, mask_in, i_in, j_in, i_out, j_out,
                          xgrid_area, nullptr, nullptr);
}

extern "C" int create_xgrid_1dx2d_order2(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out, double* xgrid_area,
                                         double* xgrid_clon, double* xgrid_clat)
{
  return create_xgrid_box(1, 2, *nlon_in, *nlat_in, lon_in, lat_in, *nlon_out, *nlat_out, lon_out, lat_out, mask_in, i_in, j_in, i_out, j_out,
                          xgrid_area, xgrid_clon, xgrid_clat);
}

extern "C" int create_xgrid_2dx1d_order1(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out, double* xgrid_area)
{
  return create_xgrid_box(0, 1, *nlon_out, *nlat_out, lon_out, lat_out, *nlon_in, *nlat_in, lon_in, lat_in, mask_in, i_in, j_in, i_out, j_out,
                          xgrid_area, nullptr, nullptr);
}

extern "C" int create_xgrid_2dx1d_order2(const int* nlon_in, const int* nlat_in, const int* nlon_out, const int* nlat_out,
                                         const double* lon_in, const double* lat_in, const double* lon_out, const double* lat_out,
                                         const double* mask_in, int* i_in, int* j_in, int* i_out, int* j_out, double* xgrid_area,
                                         double* xgrid_clon, double* xgrid_clat)
{
  return create_xgrid_box(0, 2, *nlon_out, *nlat_out, lon_out, lat_out, *nlon_in, *nlat_in, lon_in, lat_in, mask_in, i_in, j_in, i_out, j_out,
                          xgrid_area, xgrid_clon, xgrid_clat);
}

// Fortran-callable twins (create_xgrid.c:196, :296, :404, :504)
extern "C" int create_xgrid_1dx2d_order1_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2, double* xa)
{
  return create_xgrid_1dx2d_order1(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa);
}
extern "C" int create_xgrid_1dx2d_order2_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2, double* xa,
                                          double* xc, double* yc)
{
  return create_xgrid_1dx2d_order2(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa, xc, yc);
}
extern "C" int create_xgrid_2dx1d_order1_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2, double* xa)
{
  return create_xgrid_2dx1d_order1(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa);
}
extern "C" int create_xgrid_2dx1d_order2_(const int* a, const int* b, const int* c, const int* d, const double* e, const double* f,
                                          const double* g, const double* h, const double* m, int* i1, int* j1, int* i2, int* j2, double* xa,
                                          double* xc, double* yc)
{
  return create_xgrid_2dx1d_order2(a, b, c, d, e, f, g, h, m, i1, j1, i2, j2, xa, xc, yc);
}
