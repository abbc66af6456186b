// Conservative apply path for sm_100a: do_scalar_conserve_interp (reference conserve_interp.c:507-910) and the
// order-2 gradient terms grad_c2l / a2b_ord2 (gradient_c2l.c:58-195), batched over field-levels.
//
//   dst_count / dst_fill / dst_sort_gather : one-time regrouping of the exchange-grid list by destination cell
//        (a CSR over destination cells).  Inside a destination cell the entries keep the order of the original
//        list, so the per-cell sums below add in exactly the order of the reference's scatter loop
//        (conserve_interp.c:563-614, :745-811) and the remapped fields are bit-identical to the reference.
//   apply<ORDER, MISSING>  : one thread per destination cell, kApplyBT field-levels per thread; the weights of
//        a cell are read once per kApplyBT fields; output written coalesced.  No atomics.
//   grad_c2l               : one thread per source cell, kGradBT field-levels per thread: corner averages
//        (a2b_ord2 with its edge/corner rules), edge fluxes, Green's theorem, projection on (lon, lat).
//   monotone_*             : the conserve_order2_monotonic limiter (conserve_interp.c:617-742).
//
// FP64, -fmad=false, reference association order throughout.
#include "apply_internal.h"
#include "shared_div.cuh"

namespace xgb {

extern long long g_launches;

// =============================================================================================
// destination-major regrouping
// =============================================================================================
// A list entry that names a destination cell outside the tile (a stale remap file written for another output grid) is
// dropped and reported (kErrApplyIndex), never counted or written.
__global__ void __launch_bounds__(256)
dst_count_kernel(long long n, const int* __restrict__ i_out, const int* __restrict__ j_out, int nx2, int ny2, uint32_t* __restrict__ cnt,
                 int* err)
{
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int i = i_out[k], j = j_out[k];
    if (i < 0 || i >= nx2 || j < 0 || j >= ny2) { atomicOr(err, kErrApplyIndex); continue; }
    atomicAdd(&cnt[(long long)j * nx2 + i], 1u);
  }
}

__global__ void __launch_bounds__(256)
dst_fill_kernel(long long n, const int* __restrict__ i_out, const int* __restrict__ j_out, int nx2, int ny2,
                const uint32_t* __restrict__ off, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm)
{
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const int i = i_out[k], j = j_out[k];
    if (i < 0 || i >= nx2 || j < 0 || j >= ny2) continue;         // reported by dst_count_kernel
    const long long d = (long long)j * nx2 + i;
    perm[off[d] + atomicAdd(&cursor[d], 1u)] = (uint32_t)k;
  }
}

// per destination cell: restore list order inside the segment (the fill above is unordered), then gather the
// per-entry data next to each other so the apply kernel streams it
__global__ void __launch_bounds__(128)
dst_sort_gather_kernel(long long ndst, const uint32_t* __restrict__ off, uint32_t* __restrict__ perm,
                       const int* __restrict__ t_in, const int* __restrict__ i_in, const int* __restrict__ j_in,
                       const double* __restrict__ area, const double* __restrict__ di, const double* __restrict__ dj,
                       const ApplyTile* __restrict__ tiles, int ntiles,
                       int* __restrict__ c_cell, int* __restrict__ c_hidx, double* __restrict__ c_area,
                       double* __restrict__ c_di, double* __restrict__ c_dj, int* err)
{
  const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (d >= ndst) return;
  const uint32_t b = off[d], e = off[d + 1];
  for (uint32_t q = b + 1; q < e; ++q) {                 // insertion sort: segments are a handful of entries
    const uint32_t v = perm[q];
    uint32_t r = q;
    while (r > b && perm[r - 1] > v) { perm[r] = perm[r - 1]; --r; }
    perm[r] = v;
  }
  for (uint32_t q = b; q < e; ++q) {
    const uint32_t n = perm[q];
    const int t = t_in[n], i = i_in[n], j = j_in[n];
    if (t < 0 || t >= ntiles || i < 0 || j < 0 || i >= tiles[t].nx || j >= tiles[t].ny) { atomicOr(err, kErrApplyIndex); continue; }
    c_cell[q] = (int)(tiles[t].cell_off + (long long)j * tiles[t].nx + i);
    c_hidx[q] = (int)(tiles[t].halo_off + (long long)(j + 1) * (tiles[t].nx + 2) + i + 1);
    c_area[q] = area[n];
    if (di) { c_di[q] = di[n]; c_dj[q] = dj[n]; }
  }
}

void launch_dst_count(long long n, const int* i_out, const int* j_out, int nx2, int ny2, uint32_t* cnt, int* err, cudaStream_t st)
{
  if (n <= 0) return;
  ++g_launches;
  dst_count_kernel<<<148 * 8, 256, 0, st>>>(n, i_out, j_out, nx2, ny2, cnt, err);
}

void launch_dst_fill(long long n, const int* i_out, const int* j_out, int nx2, int ny2, const uint32_t* off, uint32_t* cursor,
                     uint32_t* perm, cudaStream_t st)
{
  if (n <= 0) return;
  ++g_launches;
  dst_fill_kernel<<<148 * 8, 256, 0, st>>>(n, i_out, j_out, nx2, ny2, off, cursor, perm);
}

void launch_dst_sort_gather(long long ndst, const uint32_t* off, uint32_t* perm, const int* t_in, const int* i_in, const int* j_in,
                            const double* area, const double* di, const double* dj, const ApplyTile* tiles, int ntiles,
                            ApplyCsr csr, int* err, cudaStream_t st)
{
  if (ndst <= 0) return;
  ++g_launches;
  dst_sort_gather_kernel<<<(unsigned)((ndst + 127) / 128), 128, 0, st>>>(ndst, off, perm, t_in, i_in, j_in, area, di, dj, tiles, ntiles,
                                                                         csr.cell, csr.hidx, csr.area, csr.di, csr.dj, err);
}

// =============================================================================================
// apply
// =============================================================================================
// out[f][d] = sum_q (data + grad_x*di + grad_y*dj)*area / sum_q area        (conserve_interp.c:745-839)
// Field f of the batch starts at data + f*data_stride (all source tiles concatenated; order 2: each tile with its
// one-cell halo, (nx+2)*(ny+2)); gradients and grad_mask at + f*ncell_src; output at out + f*ndst.
// XDATA: the per-entry values come from xdata[f][n] (monotone limiter output) instead of data/grad.
template <int ORDER, bool MISSING, bool XDATA>
__global__ void __launch_bounds__(128)
apply_kernel(ApplyCsr csr, long long ndst, int nf, const double* __restrict__ data, long long data_stride,
             const double* __restrict__ gx, const double* __restrict__ gy, const int* __restrict__ gmask, long long ncell_src,
             const double* __restrict__ xdata, long long nxgrid, double missing, int sum_mode, double* __restrict__ out)
{
  // field tile varies fastest over the blocks: the blocks resident at any moment cover a narrow band of destination
  // cells for ALL field tiles, so a chunk of the weights is fetched from HBM once and re-read from L2 by the other
  // field tiles (with destination fastest the 3.3 GB output stream evicted it between tiles: 3.3 GB of re-reads)
  const int nft = (nf + kApplyBT - 1) / kApplyBT;
  const long long d = (long long)(blockIdx.x / nft) * blockDim.x + threadIdx.x;
  if (d >= ndst) return;
  const int f0 = (int)(blockIdx.x % nft) * kApplyBT;
  const uint32_t b = csr.off[d], e = csr.off[d + 1];
  double acc[kApplyBT], asum[kApplyBT];
  bool seen[kApplyBT];
#pragma unroll
  for (int k = 0; k < kApplyBT; ++k) { acc[k] = 0.0; asum[k] = 0.0; seen[k] = false; }
  double asum_all = 0.0;

  for (uint32_t q = b; q < e; ++q) {
    const double area = csr.area[q];
    const int cell = csr.cell[q];
    const int src = (ORDER == 2) ? csr.hidx[q] : cell;
    double di = 0.0, dj = 0.0;
    if (ORDER == 2 && !XDATA) { di = csr.di[q]; dj = csr.dj[q]; }
    const uint32_t n = XDATA ? csr.perm[q] : 0u;
    if (!MISSING) asum_all += area;
#pragma unroll
    for (int k = 0; k < kApplyBT; ++k) {
      const int f = f0 + k;
      if (f >= nf) break;
      if (XDATA) {
        const double v = xdata[(long long)f * nxgrid + n];
        if (v == missing) continue;                                         // :729
        acc[k] += v * area;                                                 // :738
        asum[k] += area;
        continue;
      }
      const double v = data[(long long)f * data_stride + src];
      if (MISSING && v == missing) continue;                                // :575, :766
      if (ORDER == 2) {
        const long long g = (long long)f * ncell_src + cell;
        if (MISSING && gmask[g]) acc[k] += v * area;                        // :779
        else acc[k] += (v + gx[g] * di + gy[g] * dj) * area;                // :782, :806
      } else {
        acc[k] += v * area;                                                 // :586, :608
      }
      if (MISSING) { asum[k] += area; seen[k] = true; }
    }
  }
#pragma unroll
  for (int k = 0; k < kApplyBT; ++k) {
    const int f = f0 + k;
    if (f >= nf) break;
    const double a = (MISSING || XDATA) ? asum[k] : asum_all;
    const bool any = XDATA ? false : (MISSING ? seen[k] : (e > b));         // out_miss (not set by the limiter branch)
    double r;
    if (sum_mode) r = (a == 0) ? (any ? 0.0 : missing) : acc[k];            // cell_methods "sum", :821-830
    else if (a > 0) r = acc[k] / a;                                         // :833-834
    else if (any) r = 0.0;                                                  // :835-836
    else r = missing;                                                       // :837-838
    out[(long long)f * ndst + d] = r;
  }
}

void launch_apply(int order, bool has_missing, bool from_xdata, const ApplyCsr& csr, long long ndst, int nf,
                  const double* data, long long data_stride, const double* gx, const double* gy, const int* gmask,
                  long long ncell_src, const double* xdata, long long nxgrid, double missing, double* out, cudaStream_t st, int sum_mode)
{
  if (ndst <= 0 || nf <= 0) return;
  const long long nblk = ((ndst + 127) / 128) * ((nf + kApplyBT - 1) / kApplyBT);
  if (nblk >= (1ll << 31)) return;                                 // callers batch fields; unreachable for sane batches
  dim3 grid((unsigned)nblk);
  ++g_launches;
#define XGB_APPLY(O, M, X) apply_kernel<O, M, X><<<grid, 128, 0, st>>>(csr, ndst, nf, data, data_stride, gx, gy, gmask, ncell_src, xdata, nxgrid, missing, sum_mode, out)
  if (from_xdata) XGB_APPLY(2, true, true);
  else if (order == 2) { if (has_missing) XGB_APPLY(2, true, false); else XGB_APPLY(2, false, false); }
  else { if (has_missing) XGB_APPLY(1, true, false); else XGB_APPLY(1, false, false); }
#undef XGB_APPLY
}

// Fused-path variant: the gradient kernel leaves (value, grad_x, grad_y, grad_mask) of a source cell next to each other
// (32 bytes, one sector), so an exchange-cell entry costs one gather per field instead of three.
#ifndef XGB_PACK_BT
#define XGB_PACK_BT 8
#endif
#ifndef XGB_APPLY_PATCH_ROWS
#define XGB_APPLY_PATCH_ROWS 0     // 0: a block takes 128 consecutive destination cells; R: a 32-wide, R-row patch (128 / R... see kernel)
#endif
constexpr int kPackBT = XGB_PACK_BT;  // field-levels per thread (16 per thread and shared-memory staging of the entries
                                      // both measured slower on configs[1]: the L1 footprint of the gathers grows)

template <bool MISSING>
__global__ void __launch_bounds__(128)
apply_packed_kernel(ApplyCsr csr, long long ndst, int nf, const double4* __restrict__ packed, long long ncell_src,
                    double missing, int sum_mode, double* __restrict__ out, int nx_out)
{
  const int nft = (nf + kPackBT - 1) / kPackBT;
#if XGB_APPLY_PATCH_ROWS > 0
  // a block = a patch of 32 columns x 4 rows of the destination tile: its cells share far fewer source records than 128 cells
  // of one row do, so the gathers of a block stay in L1
  const int ny_out = (int)(ndst / nx_out);
  const int npx = (nx_out + 31) / 32;
  const long long patch = blockIdx.x / nft;
  const int i = (int)(patch % npx) * 32 + (threadIdx.x & 31), j = (int)(patch / npx) * 4 + (threadIdx.x >> 5);
  if (i >= nx_out || j >= ny_out) return;
  const long long d = (long long)j * nx_out + i;
#else
  const long long d = (long long)(blockIdx.x / nft) * blockDim.x + threadIdx.x;
  if (d >= ndst) return;
#endif
  const int f0 = (int)(blockIdx.x % nft) * kPackBT;
  const uint32_t b = csr.off[d], e = csr.off[d + 1];
  double acc[kPackBT], asum[kPackBT];
  bool seen[kPackBT];
#pragma unroll
  for (int k = 0; k < kPackBT; ++k) { acc[k] = 0.0; asum[k] = 0.0; seen[k] = false; }
  double asum_all = 0.0;
  for (uint32_t q = b; q < e; ++q) {
    const double area = csr.area[q], di = csr.di[q], dj = csr.dj[q];
    const int cell = csr.cell[q];
    if (!MISSING) asum_all += area;
#pragma unroll
    for (int k = 0; k < kPackBT; ++k) {
      const int f = f0 + k;
      if (f >= nf) break;
      const double2* pk = reinterpret_cast<const double2*>(&packed[(long long)f * ncell_src + cell]);
      const double2 lo = __ldg(pk), hi = __ldg(pk + 1);
      const double4 v = make_double4(lo.x, lo.y, hi.x, hi.y);
      if (MISSING && v.x == missing) continue;                              // :766
      if (MISSING && v.w != 0.0) acc[k] += v.x * area;                      // :779
      else acc[k] += (v.x + v.y * di + v.z * dj) * area;                    // :782, :806
      if (MISSING) { asum[k] += area; seen[k] = true; }
    }
  }
#pragma unroll
  for (int k = 0; k < kPackBT; ++k) {
    const int f = f0 + k;
    if (f >= nf) break;
    const double a = MISSING ? asum[k] : asum_all;
    const bool any = MISSING ? seen[k] : (e > b);
    double r;
    if (sum_mode) r = (a == 0) ? (any ? 0.0 : missing) : acc[k];
    else if (a > 0) r = acc[k] / a;
    else if (any) r = 0.0;
    else r = missing;
    out[(long long)f * ndst + d] = r;
  }
}

void launch_apply_packed(bool has_missing, const ApplyCsr& csr, long long ndst, int nf, const double* packed, long long ncell_src,
                         double missing, double* out, cudaStream_t st, int sum_mode, int nx_out)
{
  if (ndst <= 0 || nf <= 0) return;
#if XGB_APPLY_PATCH_ROWS > 0
  const long long ny_out = ndst / nx_out;
  const long long nblk = (long long)((nx_out + 31) / 32) * ((ny_out + 3) / 4) * ((nf + kPackBT - 1) / kPackBT);
#else
  const long long nblk = ((ndst + 127) / 128) * ((nf + kPackBT - 1) / kPackBT);
#endif
  if (nblk >= (1ll << 31)) return;
  ++g_launches;
  if (has_missing) apply_packed_kernel<true><<<(unsigned)nblk, 128, 0, st>>>(csr, ndst, nf, (const double4*)packed, ncell_src, missing, sum_mode, out, nx_out);
  else             apply_packed_kernel<false><<<(unsigned)nblk, 128, 0, st>>>(csr, ndst, nf, (const double4*)packed, ncell_src, missing, sum_mode, out, nx_out);
}

// =============================================================================================
// per-source-cell factors of do_scalar_conserve_interp: weight field, cell_methods "sum", cell_measures
// (conserve_interp.c:572-585, :731-737, :767-777), folded into a per-entry effective area with the reference's operation
// order; and the --target_grid rescale (:841-865)
// =============================================================================================
__global__ void __launch_bounds__(256)
effective_area_kernel(ApplyCsr csr, long long n, const double* __restrict__ weight, const double* __restrict__ carea,
                      const double* __restrict__ farea, int sum_mode, double* __restrict__ eff)
{
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int cell = csr.cell[q];
  double a = csr.area[q];
  if (weight) a *= weight[cell];
  if (sum_mode) a /= carea[cell];
  else if (farea) a *= (farea[cell] / carea[cell]);
  eff[q] = a;
}

// "data is not missing but area is missing" (:576-579, :770-773): one field-level
__global__ void __launch_bounds__(256)
measure_check_kernel(ApplyCsr csr, long long n, int order, const double* __restrict__ data, const double* __restrict__ farea,
                     double missing, double area_missing, int* err)
{
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int cell = csr.cell[q];
  const double v = data[order == 2 ? csr.hidx[q] : cell];
  if (v != missing && farea[cell] == area_missing) atomicOr(err, kErrAreaMissing);
}

__global__ void __launch_bounds__(128)
target_scale_kernel(ApplyCsr csr, long long ndst, int nf, const double* __restrict__ farea, const double* __restrict__ carea,
                    const double* __restrict__ dst_carea, double missing, double* __restrict__ out)
{
  const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (d >= ndst) return;
  double t = 0.0;
  for (uint32_t q = csr.off[d], e = csr.off[d + 1]; q < e; ++q) {
    const int cell = csr.cell[q];
    if (farea) t += (csr.area[q] * farea[cell] / carea[cell]);     // :855-856
    else t += csr.area[q];
  }
  const double s = t / dst_carea[d];
  for (int f = 0; f < nf; ++f) {
    const long long o = (long long)f * ndst + d;
    const double v = out[o];
    if (v != missing) out[o] = v * s;                                // :860-863
  }
}

void launch_effective_area(const ApplyCsr& csr, long long n, const double* weight, const double* carea, const double* farea,
                           int sum_mode, double* eff, cudaStream_t st)
{
  if (n <= 0) return;
  ++g_launches;
  effective_area_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(csr, n, weight, carea, farea, sum_mode, eff);
}

void launch_measure_check(const ApplyCsr& csr, long long n, int order, const double* data, const double* farea, double missing,
                          double area_missing, int* err, cudaStream_t st)
{
  if (n <= 0) return;
  ++g_launches;
  measure_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(csr, n, order, data, farea, missing, area_missing, err);
}

void launch_target_scale(const ApplyCsr& csr, long long ndst, int nf, const double* farea, const double* carea, const double* dst_carea,
                         double missing, double* out, cudaStream_t st)
{
  if (ndst <= 0) return;
  ++g_launches;
  target_scale_kernel<<<(unsigned)((ndst + 127) / 128), 128, 0, st>>>(csr, ndst, nf, farea, carea, dst_carea, missing, out);
}

// =============================================================================================
// grad_c2l
// =============================================================================================
// a2b_ord2 (gradient_c2l.c:124-195) at corner (ci, cj) of a tile, all four on_*_edge flags set as fregrid
// passes them (fregrid_util.c:2197-2200).  q: the field with its one-cell halo, row length nx+2.
__device__ __forceinline__ double corner_value(const double* __restrict__ q, int nx, int ny, int ci, int cj, const GradTile& g)
{
  const int w = nx + 2, nxp = nx + 1, nyp = ny + 1;
  const double r3 = 1. / 3.;
  const bool W = (ci == 0), E = (ci == nx), S = (cj == 0), N = (cj == ny);
  if (W && S) return r3 * (q[1 * w + 1] + q[1 * w] + q[1]);                              // :169
  if (E && S) return r3 * (q[1 * w + nx] + q[nx] + q[1 * w + nxp]);                       // :170
  if (E && N) return r3 * (q[ny * w + nx] + q[ny * w + nxp] + q[nyp * w + nx]);           // :171
  if (W && N) return r3 * (q[ny * w + 1] + q[ny * w] + q[nyp * w + 1]);                   // :172
  if (W) {                                                                                // :175-178
    const double a = 0.5 * (q[cj * w] + q[cj * w + 1]), b = 0.5 * (q[(cj + 1) * w] + q[(cj + 1) * w + 1]);
    return g.edge_w[cj] * a + (1 - g.edge_w[cj]) * b;
  }
  if (E) {                                                                                // :181-184
    const double a = 0.5 * (q[cj * w + nx] + q[cj * w + nxp]), b = 0.5 * (q[(cj + 1) * w + nx] + q[(cj + 1) * w + nxp]);
    return g.edge_e[cj] * a + (1 - g.edge_e[cj]) * b;
  }
  if (S) {                                                                                // :187-190
    const double a = 0.5 * (q[ci] + q[w + ci]), b = 0.5 * (q[ci + 1] + q[w + ci + 1]);
    return g.edge_s[ci] * a + (1 - g.edge_s[ci]) * b;
  }
  if (N) {                                                                                // :193-196
    const double a = 0.5 * (q[ny * w + ci] + q[nyp * w + ci]), b = 0.5 * (q[ny * w + ci + 1] + q[nyp * w + ci + 1]);
    return g.edge_n[ci] * a + (1 - g.edge_n[ci]) * b;
  }
  return 0.25 * (q[cj * w + ci] + q[cj * w + ci + 1] + q[(cj + 1) * w + ci] + q[(cj + 1) * w + ci + 1]);   // :163-166
}

// one thread per source cell (concatenated index), kGradBT field-levels per thread
// PACKED: write (value, grad_x, grad_y, grad_mask) as one double4 per cell into gx (see apply_packed_kernel)
template <bool MISSING, bool PACKED>
__global__ void __launch_bounds__(128, 3)
grad_c2l_kernel(const GradTile* __restrict__ tiles, int ntiles, long long ncell, int nf,
                const double* __restrict__ data, long long data_stride,
                double* __restrict__ gx, double* __restrict__ gy, int* __restrict__ gmask, double missing)
{
  const int nft = (nf + kGradBT - 1) / kGradBT;
  const long long c = (long long)(blockIdx.x / nft) * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int t = 0;
  while (t + 1 < ntiles && c >= tiles[t + 1].cell_off) ++t;
  const GradTile g = tiles[t];
  const int nx = g.nx, ny = g.ny, nxp = nx + 1;
  const long long lc = c - g.cell_off;
  const int i = (int)(lc % nx), j = (int)(lc / nx);
  // metrics of this cell (gradient_c2l.c:58-118), read once for all fields of the batch
  const double dx_s = g.dx[(long long)j * nx + i], dx_n = g.dx[(long long)(j + 1) * nx + i];
  const double dy_w = g.dy[(long long)j * nxp + i], dy_e = g.dy[(long long)j * nxp + i + 1];
  double en_s[3], en_nn[3], ee_w[3], ee_e[3], vlon[3], vlat[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    en_s[n] = g.en_n[3 * ((long long)j * nx + i) + n];
    en_nn[n] = g.en_n[3 * ((long long)(j + 1) * nx + i) + n];
    ee_w[n] = g.en_e[3 * ((long long)j * nxp + i) + n];
    ee_e[n] = g.en_e[3 * ((long long)j * nxp + i + 1) + n];
    vlon[n] = g.vlon[3 * lc + n];
    vlat[n] = g.vlat[3 * lc + n];
  }
  const double area = g.area[lc];
  const int f0 = (int)(blockIdx.x % nft) * kGradBT;
  for (int k = 0; k < kGradBT; ++k) {
    const int f = f0 + k;
    if (f >= nf) break;
    const double* q = data + (long long)f * data_stride + g.halo_off;
    const double p00 = corner_value(q, nx, ny, i, j, g), p10 = corner_value(q, nx, ny, i + 1, j, g);
    const double p01 = corner_value(q, nx, ny, i, j + 1, g), p11 = corner_value(q, nx, ny, i + 1, j + 1, g);
    double g3[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const double pdx_s = 0.5 * (p00 + p10) * dx_s * en_s[n];                // :86-92
      const double pdx_n = 0.5 * (p01 + p11) * dx_n * en_nn[n];
      const double pdy_w = 0.5 * (p00 + p01) * dy_w * ee_w[n];                // :94-99
      const double pdy_e = 0.5 * (p10 + p11) * dy_e * ee_e[n];
      g3[n] = pdx_n - pdx_s - pdy_w + pdy_e;                                  // :102-107
    }
    double vx = (vlon[0] * g3[0] + vlon[1] * g3[1] + vlon[2] * g3[2]) / area; // :110-117
    vx *= kRadius;
    double vy = (vlat[0] * g3[0] + vlat[1] * g3[1] + vlat[2] * g3[2]) / area;
    vy *= kRadius;
    const long long o = (long long)f * ncell + c;
    int m = 0;
    if (MISSING && (PACKED || gmask)) {                                       // fregrid_util.c:2203-2216
      const int w = nx + 2, ii = i + 1, jj = j + 1;
      m = (q[(jj - 1) * w + ii - 1] == missing || q[(jj - 1) * w + ii] == missing || q[(jj - 1) * w + ii + 1] == missing ||
           q[jj * w + ii - 1] == missing || q[jj * w + ii + 1] == missing || q[(jj + 1) * w + ii - 1] == missing ||
           q[(jj + 1) * w + ii] == missing || q[(jj + 1) * w + ii + 1] == missing) ? 1 : 0;
    }
    if (PACKED) {
      reinterpret_cast<double4*>(gx)[o] = make_double4(q[(long long)(j + 1) * (nx + 2) + i + 1], vx, vy, (double)m);
    } else {
      gx[o] = vx; gy[o] = vy;
      if (gmask) gmask[o] = m;
    }
  }
}

void launch_grad_c2l(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride,
                     double* gx, double* gy, int* gmask, bool has_missing, double missing, cudaStream_t st)
{
  if (ncell <= 0 || nf <= 0) return;
  const long long nblk = ((ncell + 127) / 128) * ((nf + kGradBT - 1) / kGradBT);
  if (nblk >= (1ll << 31)) return;
  dim3 grid((unsigned)nblk);
  ++g_launches;
  if (has_missing) grad_c2l_kernel<true, false><<<grid, 128, 0, st>>>(tiles, ntiles, ncell, nf, data, data_stride, gx, gy, gmask, missing);
  else             grad_c2l_kernel<false, false><<<grid, 128, 0, st>>>(tiles, ntiles, ncell, nf, data, data_stride, gx, gy, gmask, missing);
}

void launch_grad_c2l_packed(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride,
                            double* packed, bool has_missing, double missing, cudaStream_t st)
{
  if (ncell <= 0 || nf <= 0) return;
  const long long nblk = ((ncell + 127) / 128) * ((nf + kGradBT - 1) / kGradBT);
  if (nblk >= (1ll << 31)) return;
  dim3 grid((unsigned)nblk);
  ++g_launches;
  if (has_missing) grad_c2l_kernel<true, true><<<grid, 128, 0, st>>>(tiles, ntiles, ncell, nf, data, data_stride, packed, nullptr, nullptr, missing);
  else             grad_c2l_kernel<false, true><<<grid, 128, 0, st>>>(tiles, ntiles, ncell, nf, data, data_stride, packed, nullptr, nullptr, missing);
}

// =============================================================================================
// Field-transposed order-2 path (xgb_plan_regrid, round 2).  ncu on apply_packed_kernel: 17 % of HBM, LSU wavefronts 59 % —
// with field-major arrays every (exchange cell, field-level) is its own 32-byte gather.  Here the gradient kernel leaves, per
// SOURCE CELL, the values / grad_x / grad_y (/ grad_mask) of all field-levels next to each other: see rec_index below.
// The apply kernel puts the 32 lanes of a warp on 32 consecutive field-levels of ONE destination cell: an exchange-cell entry
// is three 256-byte coalesced loads for the whole warp, its weights are warp-uniform, every lane adds its field's terms in list
// order (the reference's order, conserve_interp.c:785-812), and a block's 32 x 32 (destination, field) results go through
// shared memory so the field-major output rows are written 256 bytes at a time.
// =============================================================================================
// record layout: rec[cell][f / 64][comp][f % 64] — per source cell and per chunk of 64 field-levels (what one warp of the apply
// kernel owns), the values, grad_x, grad_y (, grad_mask) are 512-byte rows one after the other.  A lane reaches everything it
// needs of a source cell from ONE address with constant offsets (before: three addresses nfp apart, fourteen integer
// instructions of 64-bit address arithmetic per entry), and every 16-byte-per-lane load of the warp is 512 contiguous bytes
// (a layout with the components of a field-level PAIR adjacent was tried first: same instruction count, but each load then
// touched 48 sectors for 16 sectors of data and the kernel stayed L1-bound at 2.14 ms).  nfp = nf rounded up to 64.
constexpr int kRecLane = 64;
__host__ __device__ __forceinline__ long long rec_index(long long cell, int f, int comp, int nc, int nfp)
{
  return ((cell * (nfp / kRecLane) + (f / kRecLane)) * nc + comp) * kRecLane + (f % kRecLane);
}
constexpr int kRecFields = 8;      // field-levels a gradient block transposes at a time
#ifndef XGB_REC_CHUNK
#define XGB_REC_CHUNK 16   // configs[1] regrid: 8: 2.288 ms, 16: 2.291, 32: 2.32, 64: 2.42, 128: 2.66
#endif
constexpr int kRecChunk = XGB_REC_CHUNK;      // field-levels per gradient block

#ifndef XGB_GRAD_UNROLL
#define XGB_GRAD_UNROLL 1
#endif
#ifndef XGB_GRAD_BLOCKS
#define XGB_GRAD_BLOCKS 4
#endif
#if XGB_GRAD_BLOCKS > 0
#define XGB_GRAD_BOUNDS __launch_bounds__(128, XGB_GRAD_BLOCKS)
#else
#define XGB_GRAD_BOUNDS __launch_bounds__(128)
#endif
constexpr int kGradUnroll = XGB_GRAD_UNROLL;
// XGB_GRAD_SPLIT: threads are dealt the INTERIOR cells of all tiles first (every corner of such a cell is the plain four-point
// average, gradient_c2l.c:163-166: nine loads, no case analysis), then the cells on the tiles' rims (a2b_ord2's edge and corner
// formulas).  With cells in storage order two of the three warps of a 96-cell row held one rim cell and ran both code paths
// for every field-level; ncu: 373 instructions per (cell, field-level), FP64 pipe 22 %.  configs[1] regrid (gradient + apply),
// results md5-identical: storage order 2.54 ms; interior first 2.40 ms (152 registers, 3 blocks / SM); interior first at
// __launch_bounds__(128, 4) (128 registers, no spill) 2.32 ms = the default; the field-level loop unrolled by 2 on top: 2.34
// (spills at 4 blocks) / 2.40 ms (168 registers, 3 blocks).
#ifndef XGB_GRAD_SPLIT
#define XGB_GRAD_SPLIT 1
#endif

// slot -> (tile, i, j): interior cells of tile 0, 1, ... then rim cells of tile 0, 1, ...; tiles narrower than 3 cells are all rim
__device__ __forceinline__ bool grad_slot_cell(const GradTile* __restrict__ tiles, int ntiles, long long s, int* t_out, int* i_out, int* j_out)
{
  long long rem = s;
  for (int t = 0; t < ntiles; ++t) {
    const int nx = tiles[t].nx, ny = tiles[t].ny;
    const long long cnt = (nx >= 3 && ny >= 3) ? (long long)(nx - 2) * (ny - 2) : 0;
    if (rem < cnt) { *t_out = t; *i_out = 1 + (int)(rem % (nx - 2)); *j_out = 1 + (int)(rem / (nx - 2)); return true; }
    rem -= cnt;
  }
  for (int t = 0; t < ntiles; ++t) {
    const int nx = tiles[t].nx, ny = tiles[t].ny;
    const bool thin = !(nx >= 3 && ny >= 3);
    const long long cnt = thin ? (long long)nx * ny : 2ll * nx + 2ll * ny - 4;
    if (rem < cnt) {
      *t_out = t;
      if (thin) { *i_out = (int)(rem % nx); *j_out = (int)(rem / nx); }
      else if (rem < nx) { *i_out = (int)rem; *j_out = 0; }
      else if (rem < 2ll * nx) { *i_out = (int)(rem - nx); *j_out = ny - 1; }
      else { const long long k = rem - 2ll * nx; *j_out = 1 + (int)(k >> 1); *i_out = (k & 1) ? nx - 1 : 0; }
      return false;
    }
    rem -= cnt;
  }
  *t_out = 0; *i_out = 0; *j_out = 0;
  return false;
}

template <bool MISSING>
__global__ void XGB_GRAD_BOUNDS
grad_c2l_rec_kernel(const GradTile* __restrict__ tiles, int ntiles, long long ncell, int nf, int nfp,
                    const double* __restrict__ data, long long data_stride, double* __restrict__ rec, double missing)
{
  constexpr int NC = MISSING ? 4 : 3;
  __shared__ double tile[NC][128][kRecFields + 1];
  __shared__ long long s_cell[128];
  const long long c0 = (long long)blockIdx.x * 128;
  long long c = c0 + threadIdx.x;                   // XGB_GRAD_SPLIT: a slot, turned into the cell it stands for below
  const bool live = c < ncell;
  GradTile g = tiles[0];
  int i = 0, j = 0, nx = 1, ny = 1, nxp = 2;
  long long lc = 0;
  bool interior = false;
  double dx_s = 0, dx_n = 0, dy_w = 0, dy_e = 0, area = 1;
  double en_s[3], en_nn[3], ee_w[3], ee_e[3], vlon[3], vlat[3];
  if (live) {
    int t = 0;
#if XGB_GRAD_SPLIT
    interior = grad_slot_cell(tiles, ntiles, c, &t, &i, &j);
    g = tiles[t];
    nx = g.nx; ny = g.ny; nxp = nx + 1;
    lc = (long long)j * nx + i;
    c = g.cell_off + lc;
#else
    while (t + 1 < ntiles && c >= tiles[t + 1].cell_off) ++t;
    g = tiles[t];
    nx = g.nx; ny = g.ny; nxp = nx + 1;
    lc = c - g.cell_off;
    i = (int)(lc % nx); j = (int)(lc / nx);
#endif
    // metrics of this cell (gradient_c2l.c:58-118), read once for all field-levels
    dx_s = g.dx[(long long)j * nx + i]; dx_n = g.dx[(long long)(j + 1) * nx + i];
    dy_w = g.dy[(long long)j * nxp + i]; dy_e = g.dy[(long long)j * nxp + i + 1];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      en_s[n] = g.en_n[3 * ((long long)j * nx + i) + n];
      en_nn[n] = g.en_n[3 * ((long long)(j + 1) * nx + i) + n];
      ee_w[n] = g.en_e[3 * ((long long)j * nxp + i) + n];
      ee_e[n] = g.en_e[3 * ((long long)j * nxp + i + 1) + n];
      vlon[n] = g.vlon[3 * lc + n];
      vlat[n] = g.vlat[3 * lc + n];
    }
    area = g.area[lc];
  }
  s_cell[threadIdx.x] = live ? c : -1;
  // blockIdx.y: a chunk of kRecChunk field-levels (the metrics above are read once per chunk)
  const int fbeg = blockIdx.y * kRecChunk, fend = (fbeg + kRecChunk < nf) ? fbeg + kRecChunk : nf;
  for (int f0 = fbeg; f0 < fend; f0 += kRecFields) {
    const int nk = (fend - f0 < kRecFields) ? fend - f0 : kRecFields;
    if (live) {
#pragma unroll kGradUnroll
      for (int k = 0; k < nk; ++k) {
        const double* q = data + (long long)(f0 + k) * data_stride + g.halo_off;
        double p00, p10, p01, p11;
        if (interior) {                            // corner (ci, cj) = 0.25 * (q[cj][ci] + q[cj][ci+1] + q[cj+1][ci] + q[cj+1][ci+1]), :163-166
          const double* r0 = q + (long long)j * (nx + 2) + i;
          const double* r1 = r0 + (nx + 2);
          const double* r2 = r1 + (nx + 2);
          const double a0 = r0[0], a1 = r0[1], a2 = r0[2], b0 = r1[0], b1 = r1[1], b2 = r1[2], d0 = r2[0], d1 = r2[1], d2 = r2[2];
          p00 = 0.25 * (a0 + a1 + b0 + b1); p10 = 0.25 * (a1 + a2 + b1 + b2);
          p01 = 0.25 * (b0 + b1 + d0 + d1); p11 = 0.25 * (b1 + b2 + d1 + d2);
        } else {
          p00 = corner_value(q, nx, ny, i, j, g); p10 = corner_value(q, nx, ny, i + 1, j, g);
          p01 = corner_value(q, nx, ny, i, j + 1, g); p11 = corner_value(q, nx, ny, i + 1, j + 1, g);
        }
        double g3[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          const double pdx_s = 0.5 * (p00 + p10) * dx_s * en_s[n];                // :86-92
          const double pdx_n = 0.5 * (p01 + p11) * dx_n * en_nn[n];
          const double pdy_w = 0.5 * (p00 + p01) * dy_w * ee_w[n];                // :94-99
          const double pdy_e = 0.5 * (p10 + p11) * dy_e * ee_e[n];
          g3[n] = pdx_n - pdx_s - pdy_w + pdy_e;                                  // :102-107
        }
        double vx = (vlon[0] * g3[0] + vlon[1] * g3[1] + vlon[2] * g3[2]) / area; // :110-117
        vx *= kRadius;
        double vy = (vlat[0] * g3[0] + vlat[1] * g3[1] + vlat[2] * g3[2]) / area;
        vy *= kRadius;
        const int w = nx + 2, ii = i + 1, jj = j + 1;
        tile[0][threadIdx.x][k] = q[(long long)jj * w + ii];
        tile[1][threadIdx.x][k] = vx;
        tile[2][threadIdx.x][k] = vy;
        if (MISSING) {                                                            // fregrid_util.c:2203-2216
          const int m = (q[(jj - 1) * w + ii - 1] == missing || q[(jj - 1) * w + ii] == missing || q[(jj - 1) * w + ii + 1] == missing ||
                         q[jj * w + ii - 1] == missing || q[jj * w + ii + 1] == missing || q[(jj + 1) * w + ii - 1] == missing ||
                         q[(jj + 1) * w + ii] == missing || q[(jj + 1) * w + ii + 1] == missing) ? 1 : 0;
          tile[NC - 1][threadIdx.x][k] = (double)m;
        }
      }
    }
    __syncthreads();
    // transposed write-out: kRecFields consecutive field-levels of a cell and a component are 64 contiguous bytes
    for (int idx = threadIdx.x; idx < 128 * kRecFields; idx += 128) {
      const int cl = idx / kRecFields, k = idx % kRecFields;
      const long long cc = s_cell[cl];
      if (cc >= 0 && k < nk) {
#pragma unroll
        for (int comp = 0; comp < NC; ++comp)
          rec[rec_index(cc, f0 + k, comp, NC, nfp)] = tile[comp][cl][k];
      }
    }
    __syncthreads();
  }
}

// A block = a tile of kTileD consecutive destination cells x up to 256 field-levels; warp w owns 64 of the field-levels (two per
// lane, 16-byte loads) and walks the tile's destination cells one after the other.  The tile's exchange-cell entries sit in
// shared memory as 32-byte records (read once per block), so an entry costs a warp two broadcast 16-byte reads, three 512-byte
// coalesced loads and twelve FP64 operations for 64 results.  Instruction counts per 64 results (ncu, configs[1]): first
// version (one destination cell x 32 fields per warp, entries from global memory) 350; with staged entries 215, of which 44
// were the both-ways select between staged and global entries and 40 the range checks of a shared-reciprocal division that
// did not pay at two quotients per reciprocal (removed here; it stays in the gradient kernel, 64 quotients per reciprocal).
#ifndef XGB_TILE_D
#define XGB_TILE_D 8
#endif
constexpr int kTileD = XGB_TILE_D;
constexpr int kTileE = 12 * XGB_TILE_D;   // entries staged per tile (mean 1.6 per destination cell on configs[1]); more: generic loop

struct __align__(16) TileEntry { double area, di, dj; long long off; };   // off: byte offset of the source cell's records

// a source cell's record for this lane's two field-levels
template <bool MISSING>
struct RecPair {
  double2 v, gx, gy, gm;
  // recf: this lane's pair record of source cell 0; off: byte offset of the source cell
  __device__ __forceinline__ void load(const char* __restrict__ recf, long long off)
  {
    const double2* rc = reinterpret_cast<const double2*>(recf + off);
    v = __ldg(rc);
    gx = __ldg(rc + kRecLane / 2);
    gy = __ldg(rc + kRecLane);
    if (MISSING) gm = __ldg(rc + 3 * kRecLane / 2);
  }
  __device__ __forceinline__ void add(double area, double di, double dj, double missing, double& acc0, double& acc1, double& as0,
                                      double& as1, bool& seen0, bool& seen1) const
  {
    if (MISSING) {
      if (v.x != missing) {                                               // :766
        if (gm.x != 0.0) acc0 += v.x * area;                              // :779
        else acc0 += (v.x + gx.x * di + gy.x * dj) * area;                // :782, :806
        as0 += area; seen0 = true;
      }
      if (v.y != missing) {
        if (gm.y != 0.0) acc1 += v.y * area;
        else acc1 += (v.y + gx.y * di + gy.y * dj) * area;
        as1 += area; seen1 = true;
      }
    } else {
      acc0 += (v.x + gx.x * di + gy.x * dj) * area;
      acc1 += (v.y + gx.y * di + gy.y * dj) * area;
      as0 += area;
    }
  }
};

#ifndef XGB_APPLY_BLOCKS
#define XGB_APPLY_BLOCKS 8
#endif

// XGB_TILE_R > 1 (measured, off): a block walks kTileR consecutive destination ROWS of its kTileD columns, one after the
// other — the source cells one row gathered from are the ones the next row asks for (a C96 cell spans four quarter-degree
// rows), so their records could come from the L1 instead of the L2 again (ncu on the one-row version: L1 hit rate 25 %,
// L2 75 %).  configs[1] regrid, results md5-identical: 1 row 2.32 ms, 2 rows 2.39, 4 rows 2.34, 8 rows 2.35, 16 rows 2.36 ms —
// where the records come from is not what the kernel waits for; it stays at one row.
#ifndef XGB_TILE_R
#define XGB_TILE_R 1
#endif
constexpr int kTileR = XGB_TILE_R;

// WF: warps of the block along the field-levels (64 each).  With few field-levels (a rank's share of a batch on 8 GPUs: 50)
// a block of four field-warps would have three of them idle; the block then takes 4 / WF destination tiles instead, one per
// group of WF warps, each group staging its own tile.
template <bool MISSING, int WF>
__global__ void __launch_bounds__(128, XGB_APPLY_BLOCKS)
apply_rec_kernel(ApplyCsr csr, long long ndst, int nx2, int nf, int nfp, const double* __restrict__ rec, double missing, int sum_mode,
                 double* __restrict__ out)
{
  constexpr int NC = MISSING ? 4 : 3;
  constexpr int kResRow = 64 + 4;                                 // row stride = 4 mod 16 doubles: the transposed reads below are conflict-free
  __shared__ __align__(16) double res[4][kTileD][kResRow];        // [warp][destination cell][field]
  constexpr int NT = 4 / WF;                                      // destination tiles per block
  constexpr int G = WF * 32;                                      // threads of one tile's group
  static_assert(kTileR == 1 || WF == 4, "row-walking patches take one tile per block");
  __shared__ uint32_t s_off_all[NT][kTileD + 1];
  __shared__ TileEntry s_ent_all[NT][kTileE];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tsub = wid / WF, wf = wid % WF, lt = threadIdx.x - tsub * G;
  uint32_t* s_off = s_off_all[tsub];
  TileEntry* s_ent = s_ent_all[tsub];
  const long long cbytes = (long long)NC * nfp * 8;                // bytes of one source cell's records
  const int fw = blockIdx.y * (WF * 64) + wf * 64;               // this warp's 64 field-levels
  const bool wactive = fw < nf;                                   // warp-uniform
  const int f = fw + 2 * lane;                                    // this lane's two: f, f + 1
  const bool l0 = wactive && f < nf;
  const int tiles_per_row = (nx2 + kTileD - 1) / kTileD;
  const long long ny2 = ndst / nx2;
#pragma unroll 1
  for (int row = 0; row < kTileR; ++row) {
  long long d0;
  int nd;
  if (kTileR == 1) {                                              // flat tiles over the destination index (rows may be ragged)
    d0 = ((long long)blockIdx.x * NT + tsub) * kTileD;
    nd = (ndst - d0 < kTileD) ? (int)(ndst - d0) : kTileD;
    if (nd < 0) nd = 0;                                           // a tile past the end (the last block's spare groups)
  } else {
    const long long j = (long long)(blockIdx.x / tiles_per_row) * kTileR + row;
    const int i0 = (int)(blockIdx.x % tiles_per_row) * kTileD;
    if (j >= ny2) break;                                          // block-uniform
    d0 = j * nx2 + i0;
    nd = (nx2 - i0 < kTileD) ? nx2 - i0 : kTileD;
    __syncthreads();                                              // the previous row's staging is no longer read
  }
  if (lt <= kTileD) s_off[lt] = (nd > 0) ? csr.off[d0 + (lt < nd ? lt : nd)] : 0u;
  __syncthreads();
  const uint32_t q0 = s_off[0];
  const int ne = (int)(s_off[nd] - q0);
  const bool staged = ne <= kTileE;
  if (staged)
    for (int k = lt; k < ne; k += G)
      s_ent[k] = TileEntry{csr.area[q0 + k], csr.di[q0 + k], csr.dj[q0 + k], (long long)csr.cell[q0 + k] * cbytes};
  __syncthreads();
  if (!wactive) continue;                                         // (still takes part in the barriers of the next row)
  const char* recf = reinterpret_cast<const char*>(rec + (long long)(fw / kRecLane) * NC * kRecLane + 2 * lane);
  // result of one destination cell from this lane's sums (conserve_interp.c:821-838)
  auto finish = [&](bool any, double acc0, double acc1, double as0, double as1, bool seen0, bool seen1) {
    double r0, r1;
    if (!MISSING) {
      if (sum_mode) { r0 = (as0 == 0) ? (any ? 0.0 : missing) : acc0; r1 = (as0 == 0) ? (any ? 0.0 : missing) : acc1; }   // :821-830
      else if (as0 > 0) { r0 = acc0 / as0; r1 = acc1 / as0; }                                                             // :833-834
      else { r0 = r1 = any ? 0.0 : missing; }                                                                             // :835-838
    } else if (sum_mode) {
      r0 = (as0 == 0) ? (seen0 ? 0.0 : missing) : acc0; r1 = (as1 == 0) ? (seen1 ? 0.0 : missing) : acc1;
    } else {
      r0 = (as0 > 0) ? acc0 / as0 : (seen0 ? 0.0 : missing);
      r1 = (as1 > 0) ? acc1 / as1 : (seen1 ? 0.0 : missing);
    }
    return make_double2(r0, r1);
  };
  // (tried, each slower than this plain per-cell loop at 1.87 ms, scripts/apply_variants.sh: the records of the two source
  // cells used last kept in registers, 1.97 ms — a third of the loads, half as many more instructions; one flat loop over the
  // tile's entries with the records of entry q + 1 requested before entry q is summed, 2.09 ms at 80 registers / 6 blocks,
  // 2.31 ms with spills at 64 / 8: the kernel lives on resident warps, not on loads in flight per warp; writing the results
  // out every 4 / 2 destination cells so that the staging shrinks and the L1 grows, 2.38 / 4.06 ms — the output rows want
  // segments of at least 64 bytes; 16-cell tiles with 128-byte segments, 2.36 ms at the 5 blocks their staging leaves)
  for (int dl = 0; dl < nd; ++dl) {
    const int b = (int)(s_off[dl] - q0), e = (int)(s_off[dl + 1] - q0);
    double acc0 = 0.0, acc1 = 0.0, as0 = 0.0, as1 = 0.0;
    bool seen0 = false, seen1 = false;
    if (l0) {
      if (staged) {
        for (int q = b; q < e; ++q) {
          const double2 p0 = *reinterpret_cast<const double2*>(&s_ent[q].area);       // (area, di)
          const double2 p1 = *reinterpret_cast<const double2*>(&s_ent[q].dj);         // (dj, offset)
          RecPair<MISSING> rr;
          rr.load(recf, __double_as_longlong(p1.y));
          rr.add(p0.x, p0.y, p1.x, missing, acc0, acc1, as0, as1, seen0, seen1);
        }
      } else {
        for (int q = b; q < e; ++q) {
          RecPair<MISSING> rr;
          rr.load(recf, (long long)csr.cell[q0 + q] * cbytes);
          rr.add(csr.area[q0 + q], csr.di[q0 + q], csr.dj[q0 + q], missing, acc0, acc1, as0, as1, seen0, seen1);
        }
      }
    }
    *reinterpret_cast<double2*>(&res[wid][dl][2 * lane]) = finish(e > b, acc0, acc1, as0, as1, seen0, seen1);
  }
  __syncwarp();
  // field-major rows: kTileD destination cells are contiguous; 32 / kTileD rows per pass
  constexpr int kRowsPerPass = 32 / kTileD;
  const int dl = lane % kTileD, sub = lane / kTileD;
  double* o = out + (long long)(fw + sub) * ndst + d0 + dl;
  const long long ostep = (long long)kRowsPerPass * ndst;
  if (fw + 64 <= nf && nd == kTileD) {
#pragma unroll 4
    for (int fr = sub; fr < 64; fr += kRowsPerPass, o += ostep) *o = res[wid][dl][fr];
  } else {
    for (int fr = sub; fr < 64; fr += kRowsPerPass, o += ostep)
      if (fw + fr < nf && dl < nd) *o = res[wid][dl][fr];
  }
  __syncwarp();                                                   // the next row overwrites this warp's staging
  }
}

// self-check of shared_div.cuh: quotients of a[i] / b[i] by SharedDiv against the compiler's division, bit for bit
__global__ void shared_div_check_kernel(long long n, const double* __restrict__ a, const double* __restrict__ b, unsigned long long* nbad)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SharedDiv sd(b[i]);
  const double q = sd.div(a[i]), w = a[i] / b[i];
  if (__double_as_longlong(q) != __double_as_longlong(w) && !(q != q && w != w)) atomicAdd(nbad, 1ull);
}

int shared_div_check(long long n, const double* a_host, const double* b_host, unsigned long long* nbad_host)
{
  double *a = nullptr, *b = nullptr;
  unsigned long long* nbad = nullptr;
  if (cudaMalloc(&a, n * 8) != cudaSuccess || cudaMalloc(&b, n * 8) != cudaSuccess || cudaMalloc(&nbad, 8) != cudaSuccess) return 1;
  cudaMemcpy(a, a_host, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(b, b_host, n * 8, cudaMemcpyHostToDevice);
  cudaMemset(nbad, 0, 8);
  shared_div_check_kernel<<<(unsigned)((n + 255) / 256), 256>>>(n, a, b, nbad);
  const cudaError_t e = cudaMemcpy(nbad_host, nbad, 8, cudaMemcpyDeviceToHost);
  cudaFree(a); cudaFree(b); cudaFree(nbad);
  return e == cudaSuccess ? 0 : 1;
}

size_t apply_rec_doubles(long long ncell, int nf, bool has_missing)
{
  const int nfp = (nf + kRecLane - 1) / kRecLane * kRecLane;
  return (size_t)ncell * (has_missing ? 4 : 3) * nfp;
}

void launch_regrid_rec(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride, double* rec,
                       bool has_missing, double missing, const ApplyCsr& csr, long long ndst, double apply_missing, int sum_mode,
                       double* out, cudaStream_t st, int nx2)
{
  if (ncell <= 0 || nf <= 0 || ndst <= 0) return;
  if (nx2 <= 0 || ndst % nx2) nx2 = (int)ndst;                     // no row structure known: one row
  const int nfp = (nf + kRecLane - 1) / kRecLane * kRecLane;
  const dim3 gblk((unsigned)((ncell + 127) / 128), (unsigned)((nf + kRecChunk - 1) / kRecChunk));
  const long long atiles = (kTileR == 1) ? (ndst + kTileD - 1) / kTileD
                                         : (long long)((nx2 + kTileD - 1) / kTileD) * ((ndst / nx2 + kTileR - 1) / kTileR);
  const int wf = (kTileR != 1 || nf > 128) ? 4 : (nf > 64 ? 2 : 1);     // warps along the field-levels
  const long long ablocks = (atiles + 4 / wf - 1) / (4 / wf);
  if (ablocks >= (1ll << 31)) return;
  const dim3 ablk((unsigned)ablocks, (unsigned)((nf + wf * 64 - 1) / (wf * 64)));
  g_launches += 2;
  if (has_missing) grad_c2l_rec_kernel<true><<<gblk, 128, 0, st>>>(tiles, ntiles, ncell, nf, nfp, data, data_stride, rec, missing);
  else             grad_c2l_rec_kernel<false><<<gblk, 128, 0, st>>>(tiles, ntiles, ncell, nf, nfp, data, data_stride, rec, missing);
#define XGB_APPLY_REC(M, W) apply_rec_kernel<M, W><<<ablk, 128, 0, st>>>(csr, ndst, nx2, nf, nfp, rec, apply_missing, sum_mode, out)
  if (has_missing) { if (wf == 4) XGB_APPLY_REC(true, 4); else if (wf == 2) XGB_APPLY_REC(true, 2); else XGB_APPLY_REC(true, 1); }
  else             { if (wf == 4) XGB_APPLY_REC(false, 4); else if (wf == 2) XGB_APPLY_REC(false, 2); else XGB_APPLY_REC(false, 1); }
#undef XGB_APPLY_REC
}

// =============================================================================================
// monotone limiter (conserve_interp.c:617-742): one field-level per launch set
// =============================================================================================
// order-preserving map double <-> uint64 so atomicMax/atomicMin on integers implement max/min of doubles
__device__ __forceinline__ unsigned long long dkey(double v)
{
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// f_bar_max / f_bar_min over the 3x3 neighbourhood (:630-640); f_max / f_min keys initialised (:626-629)
__global__ void __launch_bounds__(128)
monotone_bounds_kernel(const ApplyTile* __restrict__ tiles, int ntiles, long long ncell, const double* __restrict__ data,
                       double missing, double* __restrict__ fbmax, double* __restrict__ fbmin,
                       unsigned long long* __restrict__ fmax_key, unsigned long long* __restrict__ fmin_key)
{
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int t = 0;
  while (t + 1 < ntiles && c >= tiles[t + 1].cell_off) ++t;
  const int nx = tiles[t].nx, w = nx + 2;
  const long long lc = c - tiles[t].cell_off;
  const int i = (int)(lc % nx), j = (int)(lc / nx);
  const double* q = data + tiles[t].halo_off;
  double mx = -1.e20, mn = 1.e20;
  for (int jj = j - 1; jj <= j + 1; ++jj)
    for (int ii = i - 1; ii <= i + 1; ++ii) {
      const double v = q[(long long)(jj + 1) * w + ii + 1];
      if (v != missing) { if (v > mx) mx = v; if (v < mn) mn = v; }
    }
  fbmax[c] = mx; fbmin[c] = mn;
  fmax_key[c] = dkey(-1.e20); fmin_key[c] = dkey(1.e20);
}

// xdata[n] = f + grad_x*di + grad_y*dj (or f where grad_mask), f_max / f_min per source cell (:647-669)
__global__ void __launch_bounds__(256)
monotone_xdata_kernel(long long nxgrid, const int* __restrict__ t_in, const int* __restrict__ i_in, const int* __restrict__ j_in,
                      const double* __restrict__ di, const double* __restrict__ dj, const ApplyTile* __restrict__ tiles,
                      const double* __restrict__ data, const double* __restrict__ gx, const double* __restrict__ gy,
                      const int* __restrict__ gmask, double missing, double* __restrict__ xdata,
                      unsigned long long* __restrict__ fmax_key, unsigned long long* __restrict__ fmin_key)
{
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= nxgrid) return;
  const int t = t_in[n], nx = tiles[t].nx;
  const long long c = tiles[t].cell_off + (long long)j_in[n] * nx + i_in[n];
  const double f = data[tiles[t].halo_off + (long long)(j_in[n] + 1) * (nx + 2) + i_in[n] + 1];
  double v = missing;
  if (f != missing) {
    v = gmask[c] ? f : f + gx[c] * di[n] + gy[c] * dj[n];
    atomicMax(&fmax_key[c], dkey(v));
    atomicMin(&fmin_key[c], dkey(v));
  }
  xdata[n] = v;
}

// rescale towards the cell mean so no exchange-cell value leaves the neighbourhood bounds (:680-714)
__global__ void __launch_bounds__(256)
monotone_adjust_kernel(long long nxgrid, const int* __restrict__ t_in, const int* __restrict__ i_in, const int* __restrict__ j_in,
                       const ApplyTile* __restrict__ tiles, const double* __restrict__ data, double missing,
                       const double* __restrict__ fbmax, const double* __restrict__ fbmin,
                       const unsigned long long* __restrict__ fmax_key, const unsigned long long* __restrict__ fmin_key,
                       double* __restrict__ xdata, int* err)
{
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= nxgrid) return;
  double v = xdata[n];
  if (v == missing) return;
  const int t = t_in[n], nx = tiles[t].nx;
  const long long c = tiles[t].cell_off + (long long)j_in[n] * nx + i_in[n];
  const double f_bar = data[tiles[t].halo_off + (long long)(j_in[n] + 1) * (nx + 2) + i_in[n] + 1];
  const double f_max = dkey_inv(fmax_key[c]), f_min = dkey_inv(fmin_key[c]);
  const double bmax = fbmax[c], bmin = fbmin[c];
  if (f_max > bmax) {
    v = f_bar + ((v - f_bar) / (f_max - f_bar)) * (bmax - f_bar);
    if (v > bmax) {
      if (v - bmax < 1.e-10) v = bmax;                                       // TOLERANCE, conserve_interp.c:37
      if (v > bmax) atomicOr(err, kErrMonotoneMax);
    }
  } else if (f_min < bmin) {
    v = f_bar + ((v - f_bar) / (f_min - f_bar)) * (bmin - f_bar);
    if (v < bmin) {
      if (bmin - v < 1.e-10) v = bmin;
      if (v < bmin) atomicOr(err, kErrMonotoneMin);
    }
  }
  xdata[n] = v;
}

void launch_monotone(long long nxgrid, const int* t_in, const int* i_in, const int* j_in, const double* di, const double* dj,
                     const ApplyTile* tiles, int ntiles, long long ncell, const double* data, const double* gx, const double* gy,
                     const int* gmask, double missing, double* fbmax, double* fbmin, unsigned long long* fmax_key,
                     unsigned long long* fmin_key, double* xdata, int* err, cudaStream_t st)
{
  if (ncell <= 0) return;
  g_launches += 3;
  monotone_bounds_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(tiles, ntiles, ncell, data, missing, fbmax, fbmin, fmax_key, fmin_key);
  if (nxgrid <= 0) return;
  const unsigned blocks = (unsigned)((nxgrid + 255) / 256);
  monotone_xdata_kernel<<<blocks, 256, 0, st>>>(nxgrid, t_in, i_in, j_in, di, dj, tiles, data, gx, gy, gmask, missing, xdata, fmax_key, fmin_key);
  monotone_adjust_kernel<<<blocks, 256, 0, st>>>(nxgrid, t_in, i_in, j_in, tiles, data, missing, fbmax, fbmin, fmax_key, fmin_key, xdata, err);
}

}  // namespace xgb
