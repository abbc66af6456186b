// make_coupler_mosaic's exchange grids on the device: atmosphere x land, atmosphere x ocean (sea ice) and land x ocean.
//
// Replaces the per-atmosphere-cell loop of the reference tool (tools/make_coupler_mosaic/make_coupler_mosaic.c:1250-1720),
// its land x ocean loop (:2556-2692) and the centroid / land_mask / ocean_mask sums after them (:1739-2030, :2694-2808), for the
// tool's default clip method (clip_2dx2d on longitude / latitude; --clip_method conserve_great_circle is refused).  The rules
// restated here, all of them the reference's:
//   * the outer cell (atmosphere; land in land x ocean) is fix_lon'd about pi, the inner cell (land, ocean) about the outer cell's
//     mean longitude (:1392-1398, :1466, :1595); clip_2dx2d(outer, inner);
//   * atm x lnd overlaps are kept WITH their vertices when area / min(area_lnd, area_atm) > area_ratio_thresh (:1485-1547);
//   * atm x ocn = clip(atm cell, ocean cell) * ocn_frac for ocean cells with ocn_frac > MIN_AREA_FRAC (:1604-1659);
//   * the area of an atm x lnd exchange cell is NOT its polygon's: it is the sum over ocean cells with lnd_frac = 1 - ocn_frac >
//     MIN_AREA_FRAC of clip(atm x lnd polygon, ocean cell) * lnd_frac, added in ocean-cell order, and the cell is written only if
//     that sum passes the area-ratio test again (:1662-1718); same for its centroid sums;
//   * land x ocean = clip(land cell, ocean cell) * ocn_frac over ocean cells with ocn_frac > MIN_AREA_FRAC (:2628-2690);
//   * order 2: tile1_distance / tile2_distance = centroid of the exchange cell minus the centroid of the parent cell, the parent's
//     being the area-weighted mean over ITS exchange cells, summed in list order (:1826-2017, :2694-2808); the nest tile's cells
//     stay out of the sums (:1856, :1894).
// Work decomposition: candidates = one warp per outer cell over row boxes -> 32-cell segment boxes -> cells of the inner mosaic
// (ballots keep the reference's inner-cell order, so no sort); clips = one thread per candidate pair with thread-private
// polygons; the ocean pass of an atm x lnd polygon = one thread per polygon walking its atmosphere cell's ocean candidates in
// order (the sum is sequential in the reference and has to stay so to be bit-identical); parent sums = per parent cell, its
// exchange cells in ascending list position.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/xgrid_b200.h"
#include "xgrid_internal.h"
#include "xgrid_plan.h"

namespace xgb { extern long long g_launches; }

namespace {
using namespace xgb;

constexpr double kMinAreaFrac = 1.e-4;     // make_coupler_mosaic.c:148
constexpr int kCap = 24;                   // thread-private polygon capacity (the reference's MV is 50; quads and pole cells need <= 16)
constexpr int kAxlCap = 16;                // stored atm x lnd polygon
constexpr unsigned kFull = 0xffffffffu;

enum : int { kErrCplPolygon = 1 << 20 };   // an atm x lnd polygon with more than kAxlCap vertices / a clip beyond kCap

template <class T>
struct Dev {
  T* p = nullptr;
  size_t n = 0;
  Dev() = default;
  Dev(const Dev&) = delete;
  Dev& operator=(const Dev&) = delete;
  ~Dev() { if (p) cudaFree(p); }
  bool alloc(size_t count)
  {
    if (p) { cudaFree(p); p = nullptr; }
    n = count;
    return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T)) == cudaSuccess;
  }
};

// -------------------------------------------------------------------------------------------------------------------
// inner cells: everything of fix_lon that does not depend on the outer cell (mosaic_util.c:672-725); xavg = x_sum / n
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
unshifted_kernel(TileDesc tile, const double* __restrict__ lon, const double* __restrict__ lat, CellSet cells, int* err)
{
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= (long long)tile.nx * tile.ny) return;
  const int i = (int)(c % tile.nx), j = (int)(c / tile.nx);
  const int nxp = tile.nx + 1;
  const double* lo = lon + tile.vert_off;
  const double* la = lat + tile.vert_off;
  const long long n0 = (long long)j * nxp + i, n3 = (long long)(j + 1) * nxp + i;
  double x[kMaxV + 2], y[kMaxV + 2];
  x[0] = lo[n0];     y[0] = la[n0];
  x[1] = lo[n0 + 1]; y[1] = la[n0 + 1];
  x[2] = lo[n3 + 1]; y[2] = la[n3 + 1];
  x[3] = lo[n3];     y[3] = la[n3];
  double ymin = y[0], ymax = y[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) { if (y[k] < ymin) ymin = y[k]; if (y[k] > ymax) ymax = y[k]; }
  double sum;
  int n = fix_lon_unshifted(x, y, 4, &sum);
  if (n <= 0 || n > kMaxV) { atomicOr(err, kErrTooManyVertices); n = 4; }
  double xmin = x[0], xmax = x[0];
  for (int k = 1; k < n; ++k) { if (x[k] < xmin) xmin = x[k]; if (x[k] > xmax) xmax = x[k]; }
  const long long g = tile.cell_off + c;
  cells.box[g] = Box{ymin, ymax, xmin, xmax};
  cells.xavg[g] = sum / n;
  cells.nv[g] = (unsigned char)n;
  for (int k = 0; k < n; ++k) {
    cells.vx[(long long)k * cells.ncell + g] = x[k];
    cells.vy[(long long)k * cells.ncell + g] = y[k];
  }
}

// boxes of 32-cell row segments, then of whole rows (they only prune; the reference's tests run on the cells)
__global__ void __launch_bounds__(128)
segbox_kernel(TileDesc tile, const Box* __restrict__ box, Box* __restrict__ segbox, int nseg)
{
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)tile.ny * nseg) return;
  const int lane = threadIdx.x & 31;
  const int row = (int)(w / nseg), seg = (int)(w % nseg);
  const int i = seg * 32 + lane;
  Box b{1e300, -1e300, 1e300, -1e300};
  if (i < tile.nx) b = box[tile.cell_off + (long long)row * tile.nx + i];
  for (int o = 16; o > 0; o >>= 1) {
    b.ymin = fmin(b.ymin, __shfl_xor_sync(kFull, b.ymin, o));
    b.ymax = fmax(b.ymax, __shfl_xor_sync(kFull, b.ymax, o));
    b.xmin = fmin(b.xmin, __shfl_xor_sync(kFull, b.xmin, o));
    b.xmax = fmax(b.xmax, __shfl_xor_sync(kFull, b.xmax, o));
  }
  if (lane == 0) segbox[w] = b;
}

__global__ void rowbox_kernel(const Box* __restrict__ segbox, int nseg, int ny, Box* __restrict__ rowbox)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ny) return;
  Box b{1e300, -1e300, 1e300, -1e300};
  for (int s = 0; s < nseg; ++s) {
    const Box q = segbox[(long long)r * nseg + s];
    b.ymin = fmin(b.ymin, q.ymin); b.ymax = fmax(b.ymax, q.ymax);
    b.xmin = fmin(b.xmin, q.xmin); b.xmax = fmax(b.xmax, q.xmax);
  }
  rowbox[r] = b;
}

// -------------------------------------------------------------------------------------------------------------------
// candidates
// -------------------------------------------------------------------------------------------------------------------
struct OuterBox { double ymin, ymax, xmin, xmax, xavg; };

// may the node hold a cell that passes cell_hit under one of the three shifts fix_lon can apply?  (x + 2pi rounds monotonically)
__device__ __forceinline__ bool node_hit(const Box& b, const OuterBox& s)
{
  if (b.ymin >= s.ymax || b.ymax <= s.ymin) return false;
  const double lo = b.xmin, hi = b.xmax;
  if (!(lo >= s.xmax || hi <= s.xmin)) return true;
  if (!(lo + kTwoPi >= s.xmax || hi + kTwoPi <= s.xmin)) return true;
  if (!(lo - kTwoPi >= s.xmax || hi - kTwoPi <= s.xmin)) return true;
  return false;
}

// the reference's tests on one (outer, inner) pair: latitude ranges of the raw corners, longitude ranges after the inner cell
// was fix_lon'd about the outer cell's mean (make_coupler_mosaic.c:1461-1473, :1595-1611, :2647-2657).  min / max of the
// shifted vertices are the shifted min / max.
__device__ __forceinline__ bool cell_hit(const CellSet& inner, long long q, const OuterBox& s)
{
  const Box b = inner.box[q];
  if (b.ymin >= s.ymax || b.ymax <= s.ymin) return false;
  double lo = b.xmin, hi = b.xmax;
  const double dx = inner.xavg[q] - s.xavg;
  if (dx < -kPi)     { lo += kTwoPi; hi += kTwoPi; }
  else if (dx > kPi) { lo -= kTwoPi; hi -= kTwoPi; }
  if (s.xmin >= hi || s.xmax <= lo) return false;
  return true;
}

struct InnerIndex {
  int ntiles;
  const TileDesc* tiles;
  const long long* row_off;   // first row box of tile t
  const long long* seg_off;   // first segment box of tile t
  const Box* rowbox;
  const Box* segbox;
};

// same_mode 1: the inner mosaic IS the outer one and only the cell itself counts (land on the atmosphere grid, :1474-1486);
//           2: only the inner tile with the outer tile's number is visited (ocean on the atmosphere mosaic, :1338-1345)
template <bool FILL>
__global__ void __launch_bounds__(128)
candidates_kernel(CellSet outer, const TileDesc* __restrict__ otiles, int notiles, CellSet inner, InnerIndex ix,
                  const double* __restrict__ inner_frac, int same_mode, const uint32_t* __restrict__ off,
                  uint32_t* __restrict__ cnt, int2* __restrict__ pairs)
{
  const long long c = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (c >= outer.ncell) return;
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const Box ob = outer.box[c];
  const OuterBox sb{ob.ymin, ob.ymax, ob.xmin, ob.xmax, outer.xavg[c]};
  const uint32_t base = FILL ? off[c] : 0u;
  uint32_t n = 0;
  if (same_mode == 1) {
    const bool hit = (lane == 0) && cell_hit(inner, c, sb) && (!inner_frac || inner_frac[c] > kMinAreaFrac);
    const unsigned votes = __ballot_sync(kFull, hit);
    if (FILL && hit) pairs[base] = make_int2((int)c, (int)c);
    n = (uint32_t)__popc(votes);
  } else {
    int otile = 0;
    for (int t = 1; t < notiles; ++t) if (c >= otiles[t].cell_off) otile = t;
    for (int t = 0; t < ix.ntiles; ++t) {
      if (same_mode == 2 && t != otile) continue;
      const TileDesc T = ix.tiles[t];
      const int nseg = (T.nx + 31) >> 5;
      for (int r0 = 0; r0 < T.ny; r0 += 32) {
        const int r = r0 + lane;
        const bool rh = (r < T.ny) && node_hit(ix.rowbox[ix.row_off[t] + r], sb);
        unsigned rows = __ballot_sync(kFull, rh);
        while (rows) {
          const int rr = r0 + __ffs(rows) - 1;
          rows &= rows - 1;
          for (int s0 = 0; s0 < nseg; s0 += 32) {
            const int s = s0 + lane;
            const bool sh = (s < nseg) && node_hit(ix.segbox[ix.seg_off[t] + (long long)rr * nseg + s], sb);
            unsigned segs = __ballot_sync(kFull, sh);
            while (segs) {
              const int ss = s0 + __ffs(segs) - 1;
              segs &= segs - 1;
              const int i = ss * 32 + lane;
              const long long q = T.cell_off + (long long)rr * T.nx + i;
              const bool hit = (i < T.nx) && cell_hit(inner, q, sb) && (!inner_frac || inner_frac[q] > kMinAreaFrac);
              const unsigned votes = __ballot_sync(kFull, hit);
              if (FILL && hit) pairs[base + n + (uint32_t)__popc(votes & below)] = make_int2((int)c, (int)q);
              n += (uint32_t)__popc(votes);
            }
          }
        }
      }
    }
  }
  if (!FILL && lane == 0) cnt[c] = n;
}

// -------------------------------------------------------------------------------------------------------------------
// geometry on thread-private polygons
// -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int load_outer(const CellSet& s, long long c, double* x, double* y)
{
  const int n = s.nv[c];
  for (int k = 0; k < n; ++k) { x[k] = s.vx[(long long)k * s.ncell + c]; y[k] = s.vy[(long long)k * s.ncell + c]; }
  return n;
}

// fix_lon(x, y, 4, tlon) of inner cell q from its tlon-independent part (mosaic_util.c:727-729)
__device__ __forceinline__ int load_inner(const CellSet& s, long long q, double tlon, double* x, double* y)
{
  const int n = s.nv[q];
  const double dx = s.xavg[q] - tlon;
  for (int k = 0; k < n; ++k) {
    double v = s.vx[(long long)k * s.ncell + q];
    if (dx < -kPi) v += kTwoPi; else if (dx > kPi) v -= kTwoPi;
    x[k] = v;
    y[k] = s.vy[(long long)k * s.ncell + q];
  }
  return n;
}

// clip_2dx2d (create_xgrid.c:1266-1341): polygon 1 cut by every side of polygon 2; result in xo / yo
__device__ int clip_2dx2d(const double* x1, const double* y1, int n1, const double* x2in, const double* y2in, int n2,
                          double* xo, double* yo, int* err)
{
  double tx[kCap], ty[kCap], x2[kMaxV + 2], y2[kMaxV + 2];
  bool wrap = false;
  for (int k = 0; k < n1; ++k) { tx[k] = x1[k]; ty[k] = y1[k]; if (tx[k] > kTwoPi || tx[k] < 0.0) wrap = true; }
  for (int k = 0; k < n2; ++k) { x2[k] = x2in[k]; y2[k] = y2in[k]; }
  if (wrap) {                                                                                  // pimod, :1279-1290, :1343-1349
    for (int k = 0; k < n1; ++k) { if (tx[k] < -kPi) tx[k] += kTwoPi; else if (tx[k] > kPi) tx[k] -= kTwoPi; }
    for (int k = 0; k < n2; ++k) { if (x2[k] < -kPi) x2[k] += kTwoPi; else if (x2[k] > kPi) x2[k] -= kTwoPi; }
  }
  int np = n1;
  double ex0 = x2[n2 - 1], ey0 = y2[n2 - 1];
  for (int e = 0; e < n2; ++e) {
    const double ex1 = x2[e], ey1 = y2[e];
    double px = tx[np - 1], py = ty[np - 1];
    bool was_in = inside_edge(ex0, ey0, ex1, ey1, px, py);
    int no = 0;
    for (int k = 0; k < np; ++k) {
      const double qx = tx[k], qy = ty[k];
      const bool is_in = inside_edge(ex0, ey0, ex1, ey1, qx, qy);
      if (is_in != was_in) {
        if (no >= kCap) { atomicOr(err, kErrCplPolygon); return 0; }
        const double dy1 = qy - py, dy2 = ey1 - ey0, dx1 = qx - px, dx2 = ex1 - ex0;
        const double ds1 = py * qx - qy * px, ds2 = ey0 * ex1 - ey1 * ex0;
        const double determ = dy2 * dx1 - dy1 * dx2;
        if (fabs(determ) < 1.0e-30) atomicOr(err, kErrParallelEdges);
        xo[no] = (dx2 * ds1 - dx1 * ds2) / determ;
        yo[no] = (dy2 * ds1 - dy1 * ds2) / determ;
        ++no;
      }
      if (is_in) {
        if (no >= kCap) { atomicOr(err, kErrCplPolygon); return 0; }
        xo[no] = qx; yo[no] = qy; ++no;
      }
      px = qx; py = qy; was_in = is_in;
    }
    np = no;
    if (np == 0) return 0;
    for (int k = 0; k < np; ++k) { tx[k] = xo[k]; ty[k] = yo[k]; }
    ex0 = ex1; ey0 = ey1;
  }
  return np;
}

__device__ __forceinline__ double lesser(double a, double b) { return a < b ? a : b; }   // the tool's min() macro

// -------------------------------------------------------------------------------------------------------------------
// atm x lnd polygons (make_coupler_mosaic.c:1431-1549)
// -------------------------------------------------------------------------------------------------------------------
struct AxlStore {
  long long npairs;
  double *px, *py;            // [kAxlCap][npairs]
  unsigned char* pn;          // vertex count, 0 = not kept
  Box* pbox;                  // (ymin, ymax, xmin, xmax) of the polygon
};

__global__ void __launch_bounds__(128)
axl_clip_kernel(CellSet atm, CellSet lnd, const double* __restrict__ area_lnd, const int2* __restrict__ pairs, int same,
                double thresh, AxlStore st, int* err)
{
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= st.npairs) return;
  const long long a = pairs[p].x, l = pairs[p].y;
  double xa[kMaxV + 2], ya[kMaxV + 2], xl[kMaxV + 2], yl[kMaxV + 2], xo[kCap], yo[kCap];
  const int na = load_outer(atm, a, xa, ya);
  const int nl = load_inner(lnd, l, atm.xavg[a], xl, yl);
  int n_out;
  if (same) { n_out = nl; for (int k = 0; k < nl; ++k) { xo[k] = xl[k]; yo[k] = yl[k]; } }
  else n_out = clip_2dx2d(xa, ya, na, xl, yl, nl, xo, yo, err);
  unsigned char keep = 0;
  if (n_out > 0) {
    const PolyView pv{xo, yo, 1};
    const double xarea = poly_area(pv, n_out);
    const double min_area = lesser(area_lnd[l], atm.area[a]);
    if (xarea / min_area > thresh) {
      if (n_out > kAxlCap) atomicOr(err, kErrCplPolygon);
      else {
        keep = (unsigned char)n_out;
        double xmin = xo[0], xmax = xo[0], ymin = yo[0], ymax = yo[0];
        for (int k = 0; k < n_out; ++k) {
          st.px[(long long)k * st.npairs + p] = xo[k];
          st.py[(long long)k * st.npairs + p] = yo[k];
          if (xo[k] < xmin) xmin = xo[k]; if (xo[k] > xmax) xmax = xo[k];
          if (yo[k] < ymin) ymin = yo[k]; if (yo[k] > ymax) ymax = yo[k];
        }
        st.pbox[p] = Box{ymin, ymax, xmin, xmax};
      }
    }
  }
  st.pn[p] = keep;
}

// atm x ocn (make_coupler_mosaic.c:1551-1660) and land x ocean (:2628-2690): clip(outer cell, ocean cell) * ocn_frac
template <int ORDER>
__global__ void __launch_bounds__(128)
ocean_clip_kernel(CellSet outer, CellSet ocn, const double* __restrict__ area_ocn, const double* __restrict__ omask,
                  const int2* __restrict__ pairs, long long npairs, double thresh, unsigned char* __restrict__ keep,
                  double* __restrict__ area, double* __restrict__ clon, double* __restrict__ clat, int* err)
{
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const long long a = pairs[p].x, o = pairs[p].y;
  const double ocn_frac = omask[o];
  unsigned char k = 0;
  if (ocn_frac > kMinAreaFrac) {
    double xa[kMaxV + 2], ya[kMaxV + 2], xc[kMaxV + 2], yc[kMaxV + 2], xo[kCap], yo[kCap];
    const double tlon = outer.xavg[a];
    const int na = load_outer(outer, a, xa, ya);
    const int no = load_inner(ocn, o, tlon, xc, yc);
    const int n_out = clip_2dx2d(xa, ya, na, xc, yc, no, xo, yo, err);
    if (n_out > 0) {
      const PolyView pv{xo, yo, 1};
      const double xarea = poly_area(pv, n_out) * ocn_frac;
      const double min_area = lesser(area_ocn[o], outer.area[a]);
      if (xarea / min_area > thresh) {
        k = 1;
        area[p] = xarea;
        if (ORDER == 2) {
          clon[p] = poly_ctrlon(pv, n_out, tlon) * ocn_frac;
          clat[p] = poly_ctrlat(pv, n_out) * ocn_frac;
        }
      }
    }
  }
  keep[p] = k;
}

// the land share of every ocean cell under an atm x lnd polygon (make_coupler_mosaic.c:1662-1718): one thread per polygon,
// ocean cells in the order of the atmosphere cell's candidate list (= the reference's no, jo, io loops)
template <int ORDER>
__global__ void __launch_bounds__(128)
axl_ocean_kernel(CellSet atm, CellSet ocn, const double* __restrict__ area_lnd, const double* __restrict__ omask,
                 const int2* __restrict__ lpairs, AxlStore st, const uint32_t* __restrict__ ooff, const int2* __restrict__ opairs,
                 double thresh, unsigned char* __restrict__ keep, double* __restrict__ area, double* __restrict__ clon,
                 double* __restrict__ clat, int* err)
{
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= st.npairs) return;
  const int nv = st.pn[p];
  if (nv == 0) { keep[p] = 0; return; }
  const long long a = lpairs[p].x, l = lpairs[p].y;
  double xp[kAxlCap], yp[kAxlCap], xc[kMaxV + 2], yc[kMaxV + 2], xo[kCap], yo[kCap];
  for (int k = 0; k < nv; ++k) { xp[k] = st.px[(long long)k * st.npairs + p]; yp[k] = st.py[(long long)k * st.npairs + p]; }
  const Box pb = st.pbox[p];
  const Box ab = atm.box[a];
  const double tlon = atm.xavg[a];
  const double min_area = lesser(area_lnd[l], atm.area[a]);
  double s_area = 0.0, s_clon = 0.0, s_clat = 0.0;
  // :1671 tests the polygon's latitude range against the ATMOSPHERE cell's (not the ocean cell's); kept as it is
  const bool lat_out = (pb.ymin >= ab.ymax || pb.ymax <= ab.ymin);
  for (uint32_t q = ooff[a]; q < ooff[a + 1]; ++q) {
    const long long o = opairs[q].y;
    const double lnd_frac = 1 - omask[o];
    if (!(lnd_frac > kMinAreaFrac)) continue;
    const int no = load_inner(ocn, o, tlon, xc, yc);
    double xo_min = xc[0], xo_max = xc[0];
    for (int k = 1; k < no; ++k) { if (xc[k] < xo_min) xo_min = xc[k]; if (xc[k] > xo_max) xo_max = xc[k]; }
    if (pb.xmin >= xo_max || pb.xmax <= xo_min || lat_out) continue;
    const int n_out = clip_2dx2d(xp, yp, nv, xc, yc, no, xo, yo, err);
    if (n_out > 0) {
      const PolyView pv{xo, yo, 1};
      const double xarea = poly_area(pv, n_out) * lnd_frac;
      if (xarea / min_area > thresh) {
        s_area += xarea;
        if (ORDER == 2) {
          s_clon += poly_ctrlon(pv, n_out, tlon) * lnd_frac;
          s_clat += poly_ctrlat(pv, n_out) * lnd_frac;
        }
      }
    }
  }
  const unsigned char k = (s_area / min_area > thresh) ? 1 : 0;
  keep[p] = k;
  area[p] = s_area;
  if (ORDER == 2) { clon[p] = s_clon; clat[p] = s_clat; }
}

// -------------------------------------------------------------------------------------------------------------------
// compaction of the kept pairs into the lists the tool writes
// -------------------------------------------------------------------------------------------------------------------
__global__ void flags_kernel(const unsigned char* __restrict__ keep, long long n, uint32_t* __restrict__ flag)
{
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n) flag[p] = keep[p] ? 1u : 0u;
}

struct DevList {
  long long n;
  int *outer, *inner;         // global cell numbers (tile-major)
  double *area, *clon, *clat;
  double *d1i, *d1j, *d2i, *d2j;
};

__global__ void emit_kernel(const unsigned char* __restrict__ keep, const uint32_t* __restrict__ pos, const int2* __restrict__ pairs,
                            long long npairs, const double* __restrict__ area, const double* __restrict__ clon,
                            const double* __restrict__ clat, DevList out)
{
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= npairs || !keep[p]) return;
  const uint32_t e = pos[p];
  out.outer[e] = pairs[p].x;
  out.inner[e] = pairs[p].y;
  out.area[e] = area[p];
  if (clon) { out.clon[e] = clon[p]; out.clat[e] = clat[p]; }
}

// -------------------------------------------------------------------------------------------------------------------
// parent-cell sums in list order.  acc = [area | clon | clat] x ncell; skip_lo..skip_hi = outer cells of the nest tile.
// -------------------------------------------------------------------------------------------------------------------
// the list is sorted by outer cell: the thread of the first entry of a run adds the run
__global__ void outer_sums_kernel(DevList L, double* __restrict__ acc, long long ncell, long long skip_lo, long long skip_hi)
{
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= L.n) return;
  const int c = L.outer[e];
  if (e > 0 && L.outer[e - 1] == c) return;
  if (c >= skip_lo && c < skip_hi) return;
  double a = acc[c], x = L.clon ? acc[ncell + c] : 0.0, y = L.clon ? acc[2 * ncell + c] : 0.0;
  for (long long k = e; k < L.n && L.outer[k] == c; ++k) {
    a += L.area[k];
    if (L.clon) { x += L.clon[k]; y += L.clat[k]; }
  }
  acc[c] = a;
  if (L.clon) { acc[ncell + c] = x; acc[2 * ncell + c] = y; }
}

__global__ void bucket_count_kernel(DevList L, uint32_t* __restrict__ cnt, long long skip_lo, long long skip_hi)
{
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= L.n) return;
  const int c = L.outer[e];
  if (c >= skip_lo && c < skip_hi) return;
  atomicAdd(&cnt[L.inner[e]], 1u);
}

__global__ void bucket_fill_kernel(DevList L, const uint32_t* __restrict__ off, uint32_t* __restrict__ cursor,
                                   uint32_t* __restrict__ bucket, long long skip_lo, long long skip_hi)
{
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= L.n) return;
  const int c = L.outer[e];
  if (c >= skip_lo && c < skip_hi) return;
  const int q = L.inner[e];
  bucket[off[q] + atomicAdd(&cursor[q], 1u)] = (uint32_t)e;
}

// per inner cell: its entries in ascending list position (insertion sort of the bucket), added one by one
__global__ void inner_sums_kernel(DevList L, const uint32_t* __restrict__ off, uint32_t* __restrict__ bucket,
                                  double* __restrict__ acc, long long ncell)
{
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const uint32_t b = off[c], n = off[c + 1] - b;
  if (n == 0) return;
  uint32_t* v = bucket + b;
  for (uint32_t i = 1; i < n; ++i) {
    const uint32_t key = v[i];
    uint32_t j = i;
    while (j > 0 && v[j - 1] > key) { v[j] = v[j - 1]; --j; }
    v[j] = key;
  }
  double a = acc[c], x = L.clon ? acc[ncell + c] : 0.0, y = L.clon ? acc[2 * ncell + c] : 0.0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t e = v[i];
    a += L.area[e];
    if (L.clon) { x += L.clon[e]; y += L.clat[e]; }
  }
  acc[c] = a;
  if (L.clon) { acc[ncell + c] = x; acc[2 * ncell + c] = y; }
}

// area-weighted centroid of every parent cell with exchange cells (make_coupler_mosaic.c:1938-1943)
__global__ void centroid_kernel(double* __restrict__ acc, long long ncell)
{
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const double a = acc[c];
  if (a > 0) { acc[ncell + c] /= a; acc[2 * ncell + c] /= a; }
}

// tile1_distance / tile2_distance (make_coupler_mosaic.c:1946-1958, :1972-1976, :1989-1994)
__global__ void distance_kernel(DevList L, const double* __restrict__ oacc, long long nouter, const double* __restrict__ iacc,
                                long long ninner)
{
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= L.n) return;
  const long long a = L.outer[e], q = L.inner[e];
  const double lon = L.clon[e] / L.area[e], lat = L.clat[e] / L.area[e];
  L.d1i[e] = lon - oacc[nouter + a];
  L.d1j[e] = lat - oacc[2 * nouter + a];
  L.d2i[e] = lon - iacc[ninner + q];
  L.d2j[e] = lat - iacc[2 * ninner + q];
}

// -------------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------------
inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct DevMosaic {
  std::vector<TileDesc> tiles;
  long long ncell = 0, nvert = 0;
  Dev<double> lon, lat;
  Dev<TileDesc> dtiles;
  Dev<Box> box, ubox, rowbox, segbox;
  Dev<double> xavg, area, vx, vy, uxavg, uvx, uvy;
  Dev<unsigned char> nv, unv;
  Dev<long long> row_off, seg_off;
  CellSet pi() { return CellSet{ncell, box.p, xavg.p, area.p, nv.p, vx.p, vy.p}; }
  CellSet un() { return CellSet{ncell, ubox.p, uxavg.p, nullptr, unv.p, uvx.p, uvy.p}; }
  InnerIndex index() { return InnerIndex{(int)tiles.size(), dtiles.p, row_off.p, seg_off.p, rowbox.p, segbox.p}; }

  bool setup(const xgb_mosaic_grid* g, bool as_inner, int* err, cudaStream_t st)
  {
    tiles.resize(g->ntiles);
    ncell = nvert = 0;
    long long nrow = 0, nsegs = 0;
    std::vector<long long> ro(g->ntiles), so(g->ntiles);
    for (int t = 0; t < g->ntiles; ++t) {
      if (g->nx[t] <= 0 || g->ny[t] <= 0) return false;
      tiles[t] = TileDesc{g->nx[t], g->ny[t], ncell, nvert};
      ro[t] = nrow; so[t] = nsegs;
      ncell += (long long)g->nx[t] * g->ny[t];
      nvert += (long long)(g->nx[t] + 1) * (g->ny[t] + 1);
      nrow += g->ny[t];
      nsegs += (long long)g->ny[t] * ((g->nx[t] + 31) / 32);
    }
    if (ncell >= 0x7fffffffLL) return false;
    bool ok = lon.alloc(nvert) && lat.alloc(nvert) && dtiles.alloc(tiles.size()) && box.alloc(ncell) && xavg.alloc(ncell) &&
              area.alloc(ncell) && nv.alloc(ncell) && vx.alloc((size_t)kMaxV * ncell) && vy.alloc((size_t)kMaxV * ncell);
    if (as_inner)
      ok = ok && ubox.alloc(ncell) && uxavg.alloc(ncell) && unv.alloc(ncell) && uvx.alloc((size_t)kMaxV * ncell) &&
           uvy.alloc((size_t)kMaxV * ncell) && rowbox.alloc(nrow) && segbox.alloc(nsegs) && row_off.alloc(tiles.size()) &&
           seg_off.alloc(tiles.size());
    if (!ok) return false;
    cudaMemcpyAsync(lon.p, g->lon, nvert * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(lat.p, g->lat, nvert * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dtiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, st);
    for (size_t t = 0; t < tiles.size(); ++t) launch_cell_precompute(tiles[t], lon.p, lat.p, pi(), err, st);
    if (as_inner) {
      cudaMemcpyAsync(row_off.p, ro.data(), ro.size() * sizeof(long long), cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(seg_off.p, so.data(), so.size() * sizeof(long long), cudaMemcpyHostToDevice, st);
      for (size_t t = 0; t < tiles.size(); ++t) {
        const TileDesc& T = tiles[t];
        const long long nc = (long long)T.nx * T.ny;
        const int nseg = (T.nx + 31) / 32;
        g_launches += 3;
        unshifted_kernel<<<nblocks(nc, 128), 128, 0, st>>>(T, lon.p, lat.p, un(), err);
        segbox_kernel<<<nblocks((long long)T.ny * nseg * 32, 128), 128, 0, st>>>(T, ubox.p, segbox.p + so[t], nseg);
        rowbox_kernel<<<nblocks(T.ny, 128), 128, 0, st>>>(segbox.p + so[t], nseg, T.ny, rowbox.p + ro[t]);
      }
      cudaStreamSynchronize(st);   // ro / so leave scope
    }
    return true;
  }
};

struct Pairs {
  Dev<uint32_t> cnt, off;
  Dev<int2> pairs;
  long long n = 0;
};

// count -> scan -> fill
bool find_pairs(DevMosaic& outer, DevMosaic& inner, const double* inner_frac, int same_mode, Pairs& P, cudaStream_t st)
{
  const long long nc = outer.ncell;
  Dev<unsigned long long> total;
  Dev<unsigned char> tmp;
  if (!P.cnt.alloc(nc + 1) || !P.off.alloc(nc + 1) || !total.alloc(1) || !tmp.alloc(scan_tmp_bytes(nc))) return false;
  ++g_launches;
  candidates_kernel<false><<<nblocks(nc * 32, 128), 128, 0, st>>>(outer.pi(), outer.dtiles.p, (int)outer.tiles.size(), inner.un(),
                                                                  inner.index(), inner_frac, same_mode, nullptr, P.cnt.p, nullptr);
  launch_exclusive_scan(P.cnt.p, P.off.p, nc, total.p, tmp.p, st);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, total.p, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return false;
  if (h >= 0xffffffffull) { xgb_set_error("xgb_make_coupler_xgrid: %llu candidate pairs exceed the 32-bit list offsets", h); return false; }
  P.n = (long long)h;
  if (!P.pairs.alloc((size_t)P.n)) return false;
  ++g_launches;
  candidates_kernel<true><<<nblocks(nc * 32, 128), 128, 0, st>>>(outer.pi(), outer.dtiles.p, (int)outer.tiles.size(), inner.un(),
                                                                 inner.index(), inner_frac, same_mode, P.off.p, nullptr, P.pairs.p);
  return true;
}

struct List {
  long long n = 0;
  Dev<int> outer, inner;
  Dev<double> area, clon, clat, d1i, d1j, d2i, d2j;
  DevList view(int order)
  {
    return DevList{n, outer.p, inner.p, area.p, order == 2 ? clon.p : nullptr, order == 2 ? clat.p : nullptr,
                   d1i.p, d1j.p, d2i.p, d2j.p};
  }
};

bool compact(const unsigned char* keep, const Pairs& P, const double* area, const double* clon, const double* clat, int order,
             List& L, cudaStream_t st)
{
  Dev<uint32_t> flag, pos;
  Dev<unsigned long long> total;
  Dev<unsigned char> tmp;
  if (!flag.alloc(P.n + 1) || !pos.alloc(P.n + 1) || !total.alloc(1) || !tmp.alloc(scan_tmp_bytes(P.n))) return false;
  if (P.n > 0) { ++g_launches; flags_kernel<<<nblocks(P.n, 256), 256, 0, st>>>(keep, P.n, flag.p); }
  launch_exclusive_scan(flag.p, pos.p, P.n, total.p, tmp.p, st);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, total.p, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess) return false;
  L.n = (long long)h;
  bool ok = L.outer.alloc(L.n) && L.inner.alloc(L.n) && L.area.alloc(L.n);
  if (order == 2)
    ok = ok && L.clon.alloc(L.n) && L.clat.alloc(L.n) && L.d1i.alloc(L.n) && L.d1j.alloc(L.n) && L.d2i.alloc(L.n) && L.d2j.alloc(L.n);
  if (!ok) return false;
  if (P.n > 0) {
    ++g_launches;
    emit_kernel<<<nblocks(P.n, 256), 256, 0, st>>>(keep, pos.p, P.pairs.p, P.n, area, order == 2 ? clon : nullptr,
                                                   order == 2 ? clat : nullptr, L.view(order));
  }
  return cudaStreamSynchronize(st) == cudaSuccess;    // flag / pos leave scope
}

bool inner_sums(List& L, int order, double* acc, long long ncell, long long skip_lo, long long skip_hi, cudaStream_t st)
{
  if (L.n == 0) return true;
  Dev<uint32_t> cnt, off, cursor, bucket;
  Dev<unsigned long long> total;
  Dev<unsigned char> tmp;
  if (!cnt.alloc(ncell + 1) || !off.alloc(ncell + 1) || !cursor.alloc(ncell) || !bucket.alloc(L.n) || !total.alloc(1) ||
      !tmp.alloc(scan_tmp_bytes(ncell)))
    return false;
  cudaMemsetAsync(cnt.p, 0, (ncell + 1) * sizeof(uint32_t), st);
  cudaMemsetAsync(cursor.p, 0, ncell * sizeof(uint32_t), st);
  g_launches += 3;
  bucket_count_kernel<<<nblocks(L.n, 256), 256, 0, st>>>(L.view(order), cnt.p, skip_lo, skip_hi);
  launch_exclusive_scan(cnt.p, off.p, ncell, total.p, tmp.p, st);
  bucket_fill_kernel<<<nblocks(L.n, 256), 256, 0, st>>>(L.view(order), off.p, cursor.p, bucket.p, skip_lo, skip_hi);
  inner_sums_kernel<<<nblocks(ncell, 128), 128, 0, st>>>(L.view(order), off.p, bucket.p, acc, ncell);
  return cudaStreamSynchronize(st) == cudaSuccess;
}

void outer_sums(List& L, int order, double* acc, long long ncell, long long skip_lo, long long skip_hi, cudaStream_t st)
{
  if (L.n == 0) return;
  ++g_launches;
  outer_sums_kernel<<<nblocks(L.n, 256), 256, 0, st>>>(L.view(order), acc, ncell, skip_lo, skip_hi);
}

// device list -> the caller's arrays: parent cells as (tile, i, j), 0-based
bool download(List& L, int order, const DevMosaic& mo, const DevMosaic& mi, xgb_coupler_list* out, cudaStream_t st)
{
  memset(out, 0, sizeof(*out));
  out->n = L.n;
  const size_t n = (size_t)L.n, cap = n ? n : 1;
  std::vector<int> ho(cap), hi(cap);
  int** ip[] = {&out->t1, &out->i1, &out->j1, &out->t2, &out->i2, &out->j2};
  for (int** q : ip) if (!(*q = (int*)malloc(cap * sizeof(int)))) return false;
  if (!(out->area = (double*)malloc(cap * sizeof(double)))) return false;
  cudaMemcpyAsync(ho.data(), L.outer.p, n * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(hi.data(), L.inner.p, n * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(out->area, L.area.p, n * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (order == 2) {
    double** dp[] = {&out->d1i, &out->d1j, &out->d2i, &out->d2j};
    double* src[] = {L.d1i.p, L.d1j.p, L.d2i.p, L.d2j.p};
    for (int k = 0; k < 4; ++k) {
      if (!(*dp[k] = (double*)malloc(cap * sizeof(double)))) return false;
      cudaMemcpyAsync(*dp[k], src[k], n * sizeof(double), cudaMemcpyDeviceToHost, st);
    }
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return false;
  auto split = [](const DevMosaic& m, int g, int* t, int* i, int* j) {
    int k = 0;
    for (int q = 1; q < (int)m.tiles.size(); ++q) if (g >= m.tiles[q].cell_off) k = q;
    const long long c = g - m.tiles[k].cell_off;
    *t = k; *i = (int)(c % m.tiles[k].nx); *j = (int)(c / m.tiles[k].nx);
  };
  for (size_t e = 0; e < n; ++e) {
    split(mo, ho[e], &out->t1[e], &out->i1[e], &out->j1[e]);
    split(mi, hi[e], &out->t2[e], &out->i2[e], &out->j2[e]);
  }
  return true;
}

void free_list(xgb_coupler_list* l)
{
  void* p[] = {l->t1, l->i1, l->j1, l->t2, l->i2, l->j2, l->area, l->d1i, l->d1j, l->d2i, l->d2j};
  for (void* q : p) free(q);
  memset(l, 0, sizeof(*l));
}

double* download_doubles(const double* d, long long n, cudaStream_t st)
{
  double* h = (double*)malloc((n ? n : 1) * sizeof(double));
  if (h) cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, st);
  return h;
}

}  // namespace

extern "C" void xgb_coupler_result_free(xgb_coupler_result* r)
{
  if (!r) return;
  free_list(&r->atmxlnd); free_list(&r->atmxocn); free_list(&r->lndxocn);
  void* p[] = {r->area_atm, r->area_lnd, r->area_ocn, r->lnd_xarea, r->ocn_xarea};
  for (void* q : p) free(q);
  memset(r, 0, sizeof(*r));
}

extern "C" int xgb_make_coupler_xgrid(int device, int interp_order, double area_ratio_thresh, int tile_nest, int lnd_same_as_atm,
                                      int ocn_same_as_atm, const xgb_mosaic_grid* atm, const xgb_mosaic_grid* lnd,
                                      const xgb_mosaic_grid* ocn, const double* omask, xgb_coupler_result* out)
{
  if (!atm || !ocn || !omask || !out || (!lnd && !lnd_same_as_atm) || (interp_order != 1 && interp_order != 2)) {
    xgb_set_error("xgb_make_coupler_xgrid: bad arguments");
    return 1;
  }
  memset(out, 0, sizeof(*out));
  if (cudaSetDevice(device) != cudaSuccess) { xgb_set_error("xgb_make_coupler_xgrid: cudaSetDevice(%d) failed", device); return 1; }
  cudaStream_t st;
  if (cudaStreamCreate(&st) != cudaSuccess) { xgb_set_error("xgb_make_coupler_xgrid: cannot create a stream"); return 1; }
  const int order = interp_order;
  int rc = 1;
  {
    Dev<int> err;
    DevMosaic A, Lown, O;
    Dev<double> mask;
    auto fail = [&](const char* what) {
      const cudaError_t e = cudaGetLastError();
      xgb_set_error("xgb_make_coupler_xgrid: %s%s%s", what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    };
    do {
      if (!err.alloc(1)) { fail("out of device memory"); break; }
      cudaMemsetAsync(err.p, 0, sizeof(int), st);
      if (!A.setup(atm, lnd_same_as_atm != 0, err.p, st)) { fail("atmosphere mosaic: bad sizes or out of device memory"); break; }
      if (!lnd_same_as_atm && !Lown.setup(lnd, true, err.p, st)) { fail("land mosaic: bad sizes or out of device memory"); break; }
      DevMosaic& L = lnd_same_as_atm ? A : Lown;
      if (!O.setup(ocn, true, err.p, st)) { fail("ocean mosaic: bad sizes or out of device memory"); break; }
      if (ocn_same_as_atm && O.tiles.size() != A.tiles.size()) { fail("ocn_same_as_atm with a different number of tiles"); break; }
      if (!mask.alloc(O.ncell)) { fail("out of device memory"); break; }
      cudaMemcpyAsync(mask.p, omask, O.ncell * sizeof(double), cudaMemcpyHostToDevice, st);
      long long skip_lo = -1, skip_hi = -1;
      if (tile_nest >= 0 && tile_nest < (int)A.tiles.size()) {
        skip_lo = A.tiles[tile_nest].cell_off;
        skip_hi = skip_lo + (long long)A.tiles[tile_nest].nx * A.tiles[tile_nest].ny;
      }

      // ---- atmosphere x land polygons, atmosphere x ocean
      Pairs PL, PO;
      if (!find_pairs(A, L, nullptr, lnd_same_as_atm ? 1 : 0, PL, st) || !find_pairs(A, O, nullptr, ocn_same_as_atm ? 2 : 0, PO, st)) {
        fail("candidate search failed");
        break;
      }
      Dev<double> px, py, larea, lclon, lclat, oarea, oclon, oclat;
      Dev<unsigned char> pn, lkeep, okeep;
      Dev<Box> pbox;
      bool ok = px.alloc((size_t)kAxlCap * PL.n) && py.alloc((size_t)kAxlCap * PL.n) && pn.alloc(PL.n) && pbox.alloc(PL.n) &&
                lkeep.alloc(PL.n) && larea.alloc(PL.n) && okeep.alloc(PO.n) && oarea.alloc(PO.n);
      if (order == 2) ok = ok && lclon.alloc(PL.n) && lclat.alloc(PL.n) && oclon.alloc(PO.n) && oclat.alloc(PO.n);
      if (!ok) { fail("out of device memory"); break; }
      AxlStore store{PL.n, px.p, py.p, pn.p, pbox.p};
      if (PL.n > 0) {
        ++g_launches;
        axl_clip_kernel<<<nblocks(PL.n, 128), 128, 0, st>>>(A.pi(), L.un(), L.area.p, PL.pairs.p, lnd_same_as_atm ? 1 : 0,
                                                            area_ratio_thresh, store, err.p);
      }
      if (PO.n > 0) {
        ++g_launches;
        if (order == 2)
          ocean_clip_kernel<2><<<nblocks(PO.n, 128), 128, 0, st>>>(A.pi(), O.un(), O.area.p, mask.p, PO.pairs.p, PO.n, area_ratio_thresh,
                                                                   okeep.p, oarea.p, oclon.p, oclat.p, err.p);
        else
          ocean_clip_kernel<1><<<nblocks(PO.n, 128), 128, 0, st>>>(A.pi(), O.un(), O.area.p, mask.p, PO.pairs.p, PO.n, area_ratio_thresh,
                                                                   okeep.p, oarea.p, nullptr, nullptr, err.p);
      }
      if (PL.n > 0) {
        ++g_launches;
        if (order == 2)
          axl_ocean_kernel<2><<<nblocks(PL.n, 128), 128, 0, st>>>(A.pi(), O.un(), L.area.p, mask.p, PL.pairs.p, store, PO.off.p, PO.pairs.p,
                                                                  area_ratio_thresh, lkeep.p, larea.p, lclon.p, lclat.p, err.p);
        else
          axl_ocean_kernel<1><<<nblocks(PL.n, 128), 128, 0, st>>>(A.pi(), O.un(), L.area.p, mask.p, PL.pairs.p, store, PO.off.p, PO.pairs.p,
                                                                  area_ratio_thresh, lkeep.p, larea.p, nullptr, nullptr, err.p);
      }
      List AXL, AXO, LXO;
      if (!compact(lkeep.p, PL, larea.p, lclon.p, lclat.p, order, AXL, st) || !compact(okeep.p, PO, oarea.p, oclon.p, oclat.p, order, AXO, st)) {
        fail("compaction failed");
        break;
      }

      // ---- land x ocean (only when the land model has its own mosaic, make_coupler_mosaic.c:2484)
      Pairs PX;
      Dev<double> xarea, xclon, xclat;
      Dev<unsigned char> xkeep;
      if (!lnd_same_as_atm) {
        if (!find_pairs(L, O, mask.p, 0, PX, st)) { fail("candidate search failed"); break; }
        ok = xkeep.alloc(PX.n) && xarea.alloc(PX.n);
        if (order == 2) ok = ok && xclon.alloc(PX.n) && xclat.alloc(PX.n);
        if (!ok) { fail("out of device memory"); break; }
        if (PX.n > 0) {
          ++g_launches;
          if (order == 2)
            ocean_clip_kernel<2><<<nblocks(PX.n, 128), 128, 0, st>>>(L.pi(), O.un(), O.area.p, mask.p, PX.pairs.p, PX.n, area_ratio_thresh,
                                                                     xkeep.p, xarea.p, xclon.p, xclat.p, err.p);
          else
            ocean_clip_kernel<1><<<nblocks(PX.n, 128), 128, 0, st>>>(L.pi(), O.un(), O.area.p, mask.p, PX.pairs.p, PX.n, area_ratio_thresh,
                                                                     xkeep.p, xarea.p, nullptr, nullptr, err.p);
        }
        if (!compact(xkeep.p, PX, xarea.p, xclon.p, xclat.p, order, LXO, st)) { fail("compaction failed"); break; }
      }

      // ---- parent sums: land_mask / ocean_mask numerators, order-2 centroid distances
      const int ncomp = (order == 2) ? 3 : 1;
      Dev<double> aacc, lacc, oacc, l2acc, o2acc;
      ok = aacc.alloc((size_t)ncomp * A.ncell) && lacc.alloc((size_t)ncomp * L.ncell) && oacc.alloc((size_t)ncomp * O.ncell);
      if (!lnd_same_as_atm && order == 2) ok = ok && l2acc.alloc((size_t)3 * L.ncell) && o2acc.alloc((size_t)3 * O.ncell);
      if (!ok) { fail("out of device memory"); break; }
      cudaMemsetAsync(aacc.p, 0, (size_t)ncomp * A.ncell * sizeof(double), st);
      cudaMemsetAsync(lacc.p, 0, (size_t)ncomp * L.ncell * sizeof(double), st);
      cudaMemsetAsync(oacc.p, 0, (size_t)ncomp * O.ncell * sizeof(double), st);
      if (!inner_sums(AXL, order, lacc.p, L.ncell, skip_lo, skip_hi, st) || !inner_sums(AXO, order, oacc.p, O.ncell, skip_lo, skip_hi, st)) {
        fail("parent sums failed");
        break;
      }
      // the masks need the raw area sums: copy them out before the centroid division touches the other two planes (it leaves plane 0)
      if (order == 2) {
        outer_sums(AXL, order, aacc.p, A.ncell, skip_lo, skip_hi, st);
        outer_sums(AXO, order, aacc.p, A.ncell, skip_lo, skip_hi, st);
        g_launches += 3;
        centroid_kernel<<<nblocks(A.ncell, 256), 256, 0, st>>>(aacc.p, A.ncell);
        centroid_kernel<<<nblocks(L.ncell, 256), 256, 0, st>>>(lacc.p, L.ncell);
        centroid_kernel<<<nblocks(O.ncell, 256), 256, 0, st>>>(oacc.p, O.ncell);
        if (AXL.n > 0) { ++g_launches; distance_kernel<<<nblocks(AXL.n, 256), 256, 0, st>>>(AXL.view(2), aacc.p, A.ncell, lacc.p, L.ncell); }
        if (AXO.n > 0) { ++g_launches; distance_kernel<<<nblocks(AXO.n, 256), 256, 0, st>>>(AXO.view(2), aacc.p, A.ncell, oacc.p, O.ncell); }
        if (!lnd_same_as_atm) {
          cudaMemsetAsync(l2acc.p, 0, (size_t)3 * L.ncell * sizeof(double), st);
          cudaMemsetAsync(o2acc.p, 0, (size_t)3 * O.ncell * sizeof(double), st);
          outer_sums(LXO, order, l2acc.p, L.ncell, -1, -1, st);
          if (!inner_sums(LXO, order, o2acc.p, O.ncell, -1, -1, st)) { fail("parent sums failed"); break; }
          g_launches += 2;
          centroid_kernel<<<nblocks(L.ncell, 256), 256, 0, st>>>(l2acc.p, L.ncell);
          centroid_kernel<<<nblocks(O.ncell, 256), 256, 0, st>>>(o2acc.p, O.ncell);
          if (LXO.n > 0) { ++g_launches; distance_kernel<<<nblocks(LXO.n, 256), 256, 0, st>>>(LXO.view(2), l2acc.p, L.ncell, o2acc.p, O.ncell); }
        }
      }

      // ---- results
      if (!download(AXL, order, A, L, &out->atmxlnd, st) || !download(AXO, order, A, O, &out->atmxocn, st) ||
          !download(LXO, order, L, O, &out->lndxocn, st)) {
        fail("cannot copy the lists to the host");
        break;
      }
      out->ncell_atm = A.ncell; out->ncell_lnd = L.ncell; out->ncell_ocn = O.ncell;
      out->area_atm = download_doubles(A.area.p, A.ncell, st);
      out->area_lnd = download_doubles(L.area.p, L.ncell, st);
      out->area_ocn = download_doubles(O.area.p, O.ncell, st);
      out->lnd_xarea = download_doubles(lacc.p, L.ncell, st);
      out->ocn_xarea = download_doubles(oacc.p, O.ncell, st);
      int herr = 0;
      cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess || !out->area_atm || !out->area_lnd || !out->area_ocn || !out->lnd_xarea || !out->ocn_xarea) {
        fail("cannot copy the results to the host");
        break;
      }
      if (herr & kErrParallelEdges) {
        xgb_set_error("the line between <x1_0,y1_0> and  <x1_1,y1_1> should not parallel to the line between <x2_0,y2_0> and  <x2_1,y2_1>");
        break;
      }
      if (herr) { xgb_set_error("xgb_make_coupler_xgrid: a cell or an overlap has more vertices than the kernels hold (error bits 0x%x)", herr); break; }
      rc = 0;
    } while (0);
  }
  cudaStreamDestroy(st);
  if (rc) xgb_coupler_result_free(out);
  return rc;
}
