// Great-circle exchange-grid kernels for sm_100a: create_xgrid_great_circle (reference create_xgrid.c:1366-1466).
//
//   gc_cell_precompute : per cell the four corners as unit vectors (latlon2xyz, :1401-1402) in the reference's
//                        clockwise order, a bounding box in xyz grown by the bulge of the cell's great-circle sides,
//                        and the cell's spherical-excess area (get_grid_great_circle_area, :98-137)
//   gc_pyramid_level   : 2x2 union of destination boxes
//   gc_candidates<FILL>: per source cell depth-first walk of the box pyramid.  The reference tests every
//                        (source, destination) pair and rejects on coordinate ranges widened by 0.05 (318 km); a pair
//                        can only produce an exchange cell when the two cells overlap, and overlapping cells have
//                        intersecting boxes, so the tighter box test enumerates a superset of the accepted pairs
//                        and the emitted list is unchanged.
//   gc_clip            : one thread per candidate pair: clip_2dx2d_great_circle + great_circle_area + the
//                        area-ratio test (:1436-1452); counts accepted pairs per source cell
// The ordered compaction is the 2dx2d path's scatter kernel (the emission order is the same: source cells row-major,
// then destination index ascending).
#include "gc_clip.cuh"
#include "xgrid_gc.h"

namespace xgb {

extern long long g_launches;

__global__ void __launch_bounds__(128)
gc_cell_precompute_kernel(TileDesc tile, const double* __restrict__ lon, const double* __restrict__ lat, GcCells cells, int* err)
{
  const long long n = (long long)tile.nx * tile.ny;
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int i = (int)(c % tile.nx), j = (int)(c / tile.nx);
  const int nxp = tile.nx + 1;
  const double* lo = lon + tile.vert_off;
  const double* la = lat + tile.vert_off;
  const long long n0 = (long long)j * nxp + i, n1 = (long long)(j + 1) * nxp + i;
  const long long idx[4] = {n0, n1, n1 + 1, n0 + 1};                     // clockwise, create_xgrid.c:1421-1427
  gc::V3 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = gc::ll2xyz(lo[idx[k]], la[idx[k]]);
  const long long g = tile.cell_off + c;
  double bl[3] = {v[0].x, v[0].y, v[0].z}, bh[3] = {v[0].x, v[0].y, v[0].z};
  double chord2 = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cells.v[(long long)(3 * k + 0) * cells.ncell + g] = v[k].x;
    cells.v[(long long)(3 * k + 1) * cells.ncell + g] = v[k].y;
    cells.v[(long long)(3 * k + 2) * cells.ncell + g] = v[k].z;
    bl[0] = fmin(bl[0], v[k].x); bh[0] = fmax(bh[0], v[k].x);
    bl[1] = fmin(bl[1], v[k].y); bh[1] = fmax(bh[1], v[k].y);
    bl[2] = fmin(bl[2], v[k].z); bh[2] = fmax(bh[2], v[k].z);
    const gc::V3& w = v[(k + 1) & 3];
    const double dx = w.x - v[k].x, dy = w.y - v[k].y, dz = w.z - v[k].z;
    chord2 = fmax(chord2, dx * dx + dy * dy + dz * dz);
  }
  // Every point of the cell is a normalised convex combination q/|q| of its corners, |q| >= cos(phi/2) with phi the
  // cell's angular diameter <= twice its longest side; its coordinates leave the corners' range by at most
  // sec(phi/2) - 1 ~ phi^2/8 <= c^2/2 (c = longest chord).  0.75 c^2 covers the higher-order terms for c^2 < 1;
  // larger cells get the whole sphere.
  const double grow = (chord2 < 1.0) ? 0.75 * chord2 + 1e-13 : 2.0;
  Box3 b;
#pragma unroll
  for (int a = 0; a < 3; ++a) { b.lo[a] = bl[a] - grow; b.hi[a] = bh[a] + grow; }
  cells.box[g] = b;
  // gridArea of the ring after addEnd's duplicate suppression (pole cells of a lat-lon grid become triangles)
  gc::Ring r;
  r.len = 0; r.overflow = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) gc::ring_append_unique(r, v[k].x, v[k].y, v[k].z);
  cells.area[g] = gc::ring_area(r);
  (void)err;
}

void launch_gc_cell_precompute(const TileDesc& tile, const double* lon, const double* lat, GcCells cells, int* err, cudaStream_t st)
{
  const long long n = (long long)tile.nx * tile.ny;
  if (n <= 0) return;
  ++g_launches;
  gc_cell_precompute_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(tile, lon, lat, cells, err);
}

__device__ __forceinline__ Box3 load_box3(const Box3* p)
{
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  Box3 r;
  r.lo[0] = a.x; r.lo[1] = a.y; r.lo[2] = b.x; r.hi[0] = b.y; r.hi[1] = c.x; r.hi[2] = c.y;
  return r;
}

__global__ void __launch_bounds__(256)
gc_pyramid_level_kernel(Pyr3Level child, Box3* __restrict__ out, int nx, int ny)
{
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)nx * ny) return;
  const int ix = (int)(t % nx), iy = (int)(t / nx);
  Box3 r;
  for (int a = 0; a < 3; ++a) { r.lo[a] = 1e300; r.hi[a] = -1e300; }
  for (int dy = 0; dy < 2; ++dy) {
    const int cy = 2 * iy + dy;
    if (cy >= child.ny) continue;
    for (int dx = 0; dx < 2; ++dx) {
      const int cx = 2 * ix + dx;
      if (cx >= child.nx) continue;
      const Box3 b = load_box3(child.box + (long long)cy * child.nx + cx);
      for (int a = 0; a < 3; ++a) { r.lo[a] = fmin(r.lo[a], b.lo[a]); r.hi[a] = fmax(r.hi[a], b.hi[a]); }
    }
  }
  out[t] = r;
}

void launch_gc_pyramid_level(const Pyr3Level& child, Box3* out, int nx, int ny, cudaStream_t st)
{
  const long long n = (long long)nx * ny;
  ++g_launches;
  gc_pyramid_level_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(child, out, nx, ny);
}

__device__ __forceinline__ bool boxes_meet(const Box3& a, const Box3& b)
{
  return a.lo[0] <= b.hi[0] && b.lo[0] <= a.hi[0] && a.lo[1] <= b.hi[1] && b.lo[1] <= a.hi[1] &&
         a.lo[2] <= b.hi[2] && b.lo[2] <= a.hi[2];
}

constexpr int kGcStack = 3 * kMaxLevels + 8;

__device__ __forceinline__ void load_cell(const GcCells& c, long long g, gc::V3* v)
{
#pragma unroll
  for (int k = 0; k < 4; ++k)
    v[k] = gc::V3{c.v[(long long)(3 * k + 0) * c.ncell + g], c.v[(long long)(3 * k + 1) * c.ncell + g], c.v[(long long)(3 * k + 2) * c.ncell + g]};
}

// Second stage of the candidate search: of the pairs whose boxes meet, keep those with no side of either cell that has the
// whole other cell beyond it (gc::separated_by_side: such a pair cannot produce an exchange cell).  A quarter of the box
// candidates of configs[2] are neighbours that merely touch the box; dropping them here, one thread per pair, instead of at
// the head of the clip kernel keeps that kernel's lanes busy (clip 22.3 -> 16.0 ms), and doing it outside the per-source-cell
// walk keeps the walk cheap (inside the walk the test cost 1.4 ms in each of its two passes).
__global__ void __launch_bounds__(256)
gc_filter_kernel(GcCells src, GcCells dst, const int2* __restrict__ pairs, unsigned long long npairs, long long s0,
                 uint32_t* __restrict__ flag)
{
  const unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const int2 pr = pairs[p];
  gc::V3 a[4], b[4];
  load_cell(src, s0 + pr.x, a);
  load_cell(dst, pr.y, b);
  flag[p] = (gc::separated_by_side(a, b) || gc::separated_by_side(b, a)) ? 0u : 1u;
}

__global__ void __launch_bounds__(256)
gc_compact_kernel(const int2* __restrict__ pairs, unsigned long long npairs, const uint32_t* __restrict__ flag,
                  const uint32_t* __restrict__ pos, int2* __restrict__ kept)
{
  const unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (p < npairs && flag[p]) kept[pos[p]] = pairs[p];
}

// per source cell: where its surviving pairs start and how many they are (pos has npairs + 1 entries)
__global__ void __launch_bounds__(256)
gc_reoffset_kernel(long long ns, uint32_t* __restrict__ pair_off, uint32_t* __restrict__ pair_cnt, const uint32_t* __restrict__ pos)
{
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= ns) return;
  const uint32_t b = pos[pair_off[t]], e = pos[pair_off[t] + pair_cnt[t]];
  pair_off[t] = b;
  pair_cnt[t] = e - b;
}

void launch_gc_filter(const GcCells& src, const GcCells& dst, const int2* pairs, unsigned long long npairs, long long s0, uint32_t* flag,
                      cudaStream_t st)
{
  if (npairs == 0) return;
  ++g_launches;
  gc_filter_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(src, dst, pairs, npairs, s0, flag);
}

void launch_gc_compact(const int2* pairs, unsigned long long npairs, const uint32_t* flag, const uint32_t* pos, int2* kept, long long ns,
                       uint32_t* pair_off, uint32_t* pair_cnt, cudaStream_t st)
{
  g_launches += 2;
  if (npairs > 0) gc_compact_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(pairs, npairs, flag, pos, kept);
  gc_reoffset_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(ns, pair_off, pair_cnt, pos);
}

// The count pass leaves the first kGcSlots candidates of every source cell in `slots` (a 1/4 degree cell meets 1.7 one-degree
// cells on average); the fill pass copies them instead of walking the pyramid a second time and walks only for the cells
// that had more.  Same candidates in the same order.
constexpr int kGcSlots = 8;
template <bool FILL>
__global__ void __launch_bounds__(128)
gc_candidate_kernel(GcCells src, GcCells dst, long long s0, long long ns, const double* __restrict__ mask, Pyramid3 pyr,
                    const uint32_t* __restrict__ pair_off, uint32_t* __restrict__ cnt, int2* __restrict__ pairs,
                    int* __restrict__ slots, int* err)
{
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= ns) return;
  const long long s = s0 + t;
  uint32_t n = 0;
  const uint32_t base = FILL ? pair_off[t] : 0u;
  if (FILL && slots) {
    const uint32_t have = cnt[t];
    if (have <= (uint32_t)kGcSlots) {
      for (uint32_t k = 0; k < have; ++k) pairs[base + k] = make_int2((int)t, slots[t * kGcSlots + k]);
      return;
    }
  }
  if (mask == nullptr || mask[s] > kMaskThresh) {                          // create_xgrid.c:1419
    const Box3 sb = load_box3(src.box + s);
    unsigned long long stack[kGcStack];
    int sp = 0;
    const int top = pyr.nlev - 1;
    {
      const Pyr3Level& L = pyr.lev[top];
      for (int iy = 0; iy < L.ny; ++iy)
        for (int ix = 0; ix < L.nx; ++ix) {
          const long long q = (long long)iy * L.nx + ix;
          if (!boxes_meet(load_box3(L.box + q), sb)) continue;
          if (top == 0) { if (FILL) pairs[base + n] = make_int2((int)t, (int)q); else if (slots && n < (uint32_t)kGcSlots) slots[t * kGcSlots + n] = (int)q; ++n; }
          else stack[sp++] = ((unsigned long long)top << 58) | ((unsigned long long)iy << 29) | (unsigned long long)ix;
        }
    }
    while (sp > 0) {
      const unsigned long long e = stack[--sp];
      const int lev = (int)(e >> 58) - 1;
      const int py = (int)((e >> 29) & 0x1fffffffull), px = (int)(e & 0x1fffffffull);
      const Pyr3Level& L = pyr.lev[lev];
      for (int dy = 0; dy < 2; ++dy) {
        const int cy = 2 * py + dy;
        if (cy >= L.ny) continue;
        for (int dx = 0; dx < 2; ++dx) {
          const int cx = 2 * px + dx;
          if (cx >= L.nx) continue;
          const long long q = (long long)cy * L.nx + cx;
          if (!boxes_meet(load_box3(L.box + q), sb)) continue;
          if (lev == 0) { if (FILL) pairs[base + n] = make_int2((int)t, (int)q); else if (slots && n < (uint32_t)kGcSlots) slots[t * kGcSlots + n] = (int)q; ++n; }
          else if (sp < kGcStack) stack[sp++] = ((unsigned long long)lev << 58) | ((unsigned long long)cy << 29) | (unsigned long long)cx;
          else atomicOr(err, kErrStackOverflow);
        }
      }
    }
  }
  if (!FILL) cnt[t] = n;
}

size_t gc_slot_bytes(long long ns) { return (size_t)ns * kGcSlots * sizeof(int); }

void launch_gc_candidates(bool fill, const GcCells& src, const GcCells& dst, long long s0, long long ns, const double* mask, const Pyramid3& pyr,
                          const uint32_t* pair_off, uint32_t* cnt, int2* pairs, int* err, cudaStream_t st, int* slots)
{
  if (ns <= 0) return;
  const unsigned blocks = (unsigned)((ns + 127) / 128);
  ++g_launches;
  if (fill) gc_candidate_kernel<true><<<blocks, 128, 0, st>>>(src, dst, s0, ns, mask, pyr, pair_off, cnt, pairs, slots, err);
  else      gc_candidate_kernel<false><<<blocks, 128, 0, st>>>(src, dst, s0, ns, mask, pyr, pair_off, cnt, pairs, slots, err);
}

// Block size.  The kernel is 6 700 SASS instructions plus the double-double routines it calls, and ncu named "no instruction"
// its top stall by a factor of six: warps of many small blocks, started at different times, sit at different places in that
// code and evict each other's instruction-cache lines.  One large block per SM starts its warps together on consecutive pairs
// (same source cell, neighbouring destination cells: the same path through the clip), and that alone is worth a third
// (configs[2] clip phase, results md5-identical; threads x blocks/SM, registers):
//    64 x 8, 122: 10.74 ms      64 x 16, 64:  9.82      128 x 8, 64: 9.56      256 x 4, 64: 8.91      512 x 2, 64: 8.26
//   512 x 1, 124:  7.69        640 x 1, 96:  7.51      768 x 1, 80: 7.28      896 x 1, 72: 7.18     1024 x 1, 64: 7.43
// Small launches (a few blocks per SM at most) keep 128-thread blocks so that every SM has work.
// Then, with 896-thread blocks: the clip routine written as phases (gc_clip.cuh, clip_great_circle_t; same arithmetic) 7.17 ->
// 5.95 ms; one block barrier between the clip and the area code (XGB_GC_AREA_BAR: the warps enter the acos-heavy area routine
// together) 5.83 ms = the default.  Barriers between the phases inside the clip as well (XGB_GC_SYNC = 1 / 2 / 3: per phase /
// per side of cell 1 / per pair of sides) cost more in waiting than they save in fetches: 6.25 / 6.52 / 7.10 ms.
#ifndef XGB_GC_THREADS
#define XGB_GC_THREADS 896
#endif
#ifndef XGB_GC_SYNC
#define XGB_GC_SYNC 0
#endif
#ifndef XGB_GC_AREA_BAR
#define XGB_GC_AREA_BAR 1
#endif
constexpr int kGcBig = XGB_GC_THREADS, kGcSmall = 128;
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
gc_clip_kernel(GcCells src, GcCells dst, const double* __restrict__ mask, const int2* __restrict__ pairs,
               unsigned long long npairs, long long s0, double* __restrict__ parea, uint32_t* __restrict__ cnt, int* err)
{
  const unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
#if XGB_GC_SYNC > 0 || XGB_GC_AREA_BAR
  const bool live = p < npairs;
  const int2 pr = pairs[live ? p : npairs - 1];
#else
  if (p >= npairs) return;
  const bool live = true;
  const int2 pr = pairs[p];
#endif
  const long long s = s0 + pr.x, d = pr.y;
  gc::V3 a[4], b[4], out[gc::kPoly];
  load_cell(src, s, a);
  load_cell(dst, d, b);
  const int n_out = gc::clip_great_circle_t<XGB_GC_SYNC>(a, 4, b, 4, out, live);   // (pairs separated by a side never get here: gc_leaf)
#if XGB_GC_SYNC > 0 || XGB_GC_AREA_BAR
  __syncthreads();                                                // the block's warps enter the area code together
  if (!live) return;
#endif
  double keep = 0.0;
  if (n_out < 0) {
    atomicOr(err, n_out == gc::kErrNotConvex ? kErrGcNotConvex : (n_out == gc::kErrPool ? kErrGcNodePool : kErrGcWalk));
  } else if (n_out > 0) {
    const double m = mask ? mask[s] : 1.0;
    const double xarea = gc::great_circle_area(n_out, [&](int k) { return out[k]; }) * m;      // :1438
    const double a1 = src.area[s], a2 = dst.area[d];
    const double min_area = (a1 < a2) ? a1 : a2;                                               // :1439
    if (xarea / min_area > kAreaRatioThresh) keep = xarea;                                     // :1440
  }
  parea[p] = keep;
  if (keep > 0.0) atomicAdd(&cnt[pr.x], 1u);
}

void launch_gc_clip(const GcCells& src, const GcCells& dst, const double* mask, const int2* pairs, unsigned long long npairs,
                    long long s0, double* parea, uint32_t* cnt, int* err, cudaStream_t st)
{
  if (npairs == 0) return;
  ++g_launches;
  if (npairs >= 4ull * 148 * kGcBig)
    gc_clip_kernel<kGcBig><<<(unsigned)((npairs + kGcBig - 1) / kGcBig), kGcBig, 0, st>>>(src, dst, mask, pairs, npairs, s0, parea, cnt, err);
  else
    gc_clip_kernel<kGcSmall><<<(unsigned)((npairs + kGcSmall - 1) / kGcSmall), kGcSmall, 0, st>>>(src, dst, mask, pairs, npairs, s0, parea, cnt, err);
}

}  // namespace xgb

// host build of the same clip, for CPU tests against the compiled reference (include/xgrid_b200.h, Part 3)
extern "C" int xgb_gc_clip_host(const double* x1, const double* y1, const double* z1, int n1,
                                const double* x2, const double* y2, const double* z2, int n2,
                                double* xo, double* yo, double* zo, double* area)
{
  using namespace xgb;
  if (n1 > gc::kMaxIn || n2 > gc::kMaxIn) return gc::kErrPool;
  gc::V3 a[gc::kMaxIn], b[gc::kMaxIn], out[gc::kPoly];
  for (int k = 0; k < n1; ++k) a[k] = gc::V3{x1[k], y1[k], z1[k]};
  for (int k = 0; k < n2; ++k) b[k] = gc::V3{x2[k], y2[k], z2[k]};
  const int n = gc::clip_great_circle(a, n1, b, n2, out);
  for (int k = 0; k < n; ++k) { xo[k] = out[k].x; yo[k] = out[k].y; zo[k] = out[k].z; }
  if (area) *area = (n > 0) ? gc::great_circle_area(n, [&](int k) { return out[k]; }) : 0.0;
  return n;
}

// gc::gc_acos / gc_asin / gc_atan2 (gc_clip.cuh) on n arguments: host build and device build, for the tests.
// fn: 0 = acos (acosl rounded to double), 1 = asin, 2 = atan2(x[i], y[i])
namespace xgb {
__device__ __host__ inline double gc_math(int fn, double x, double y)
{
  return fn == 0 ? gc::gc_acos(x) : (fn == 1 ? gc::gc_asin(x) : gc::gc_atan2(x, y));
}
__global__ void gc_math_kernel(int fn, long long n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = gc_math(fn, x[i], y ? y[i] : 0.0);
}
}  // namespace xgb

extern "C" void xgb_gc_math_host(int fn, long long n, const double* x, const double* y, double* out)
{
  for (long long i = 0; i < n; ++i) out[i] = xgb::gc_math(fn, x[i], y ? y[i] : 0.0);
}

extern "C" int xgb_gc_math_device(int fn, long long n, const double* x_host, const double* y_host, double* out_host)
{
  double *x = nullptr, *y = nullptr, *o = nullptr;
  if (n <= 0) return 0;
  if (cudaMalloc(&x, n * 8) != cudaSuccess || cudaMalloc(&o, n * 8) != cudaSuccess || (y_host && cudaMalloc(&y, n * 8) != cudaSuccess)) {
    cudaFree(x); cudaFree(o); cudaFree(y);
    return 1;
  }
  cudaMemcpy(x, x_host, n * 8, cudaMemcpyHostToDevice);
  if (y_host) cudaMemcpy(y, y_host, n * 8, cudaMemcpyHostToDevice);
  xgb::gc_math_kernel<<<(unsigned)((n + 255) / 256), 256>>>(fn, n, x, y, o);
  const cudaError_t e = cudaMemcpy(out_host, o, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(x); cudaFree(y); cudaFree(o);
  return e == cudaSuccess ? 0 : 1;
}

extern "C" void xgb_gc_acos_host(long long n, const double* x, double* out) { xgb_gc_math_host(0, n, x, nullptr, out); }
extern "C" int xgb_gc_acos_device(long long n, const double* x_host, double* out_host) { return xgb_gc_math_device(0, n, x_host, nullptr, out_host); }
