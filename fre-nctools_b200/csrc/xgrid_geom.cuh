// Device-side spherical polygon geometry for the exchange-grid kernels (sm_100a).
//
// Every routine states the arithmetic of the reference routine it replaces in the same
// association order and is compiled with -fmad=false, because each accept/reject decision of
// the exchange grid is a floating-point predicate (inside_edge <= 1e-12, area ratio > 1e-6,
// pole tests in fix_lon) and the integer cell lists have to match the reference bit for bit.
// Citations are relative to the reference tree (mlee03/FRE-NCtools).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "ref_trig.cuh"

namespace xgb {

constexpr double kPi     = 3.14159265358979323846;
constexpr double kTwoPi  = 2.0 * kPi;
constexpr double kHalfPi = 0.5 * kPi;
constexpr double kRadius = 6371000.0;          // constant.h:23
constexpr double kSmall  = 1.e-10;             // mosaic_util.h:34 SMALL_VALUE
constexpr double kPoleTol = 1.e-6;             // mosaic_util.c:35 TOLORENCE
constexpr double kAreaRatioThresh = 1.e-6;     // create_xgrid.c:27
constexpr double kMaskThresh = 0.5;            // create_xgrid.c:28
// The same numbers in constant memory for the hot device code (xgrid_kernels.cu maps the names onto these after its includes):
// an FP64 instruction takes a constant-bank operand for free, a 64-bit literal is two instructions to materialise (ref_trig.cuh).
#if defined(__CUDACC__)
static __constant__ double dev_kPi = 3.14159265358979323846, dev_kTwoPi = 2.0 * 3.14159265358979323846,
                           dev_kHalfPi = 0.5 * 3.14159265358979323846, dev_kSmall = 1.e-10, dev_kAreaRatioThresh = 1.e-6,
                           dev_kInsideTol = 1.e-12, dev_kEps30 = 1.0e-30;
#endif
constexpr double kInsideTol = 1.e-12;          // inside_edge, create_xgrid.c:2349
constexpr double kEps30 = 1.0e-30;             // EPSLN30, create_xgrid.c:1314
constexpr int    kMaxV = 8;                    // create_xgrid.c:627 MAX_V (vertices of a fix_lon'd cell)
constexpr int    kMaxClip = 16;                // capacity of a clipped polygon (<= n1 + n2 for convex cells)

// ---------------------------------------------------------------------------------------------
// fix_lon (mosaic_util.c:667-738) on a thread-private 4-vertex cell; returns the vertex count.
// x/y must have room for kMaxV + 2 entries.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int vtx_remove(double* x, double* y, int n, int at) {
  for (int k = at; k < n - 1; ++k) { x[k] = x[k + 1]; y[k] = y[k + 1]; }
  return n - 1;
}
__device__ __forceinline__ int vtx_insert(double* x, double* y, int n, int at, double lon, double lat) {
  for (int k = n - 1; k >= at; --k) { x[k + 1] = x[k]; y[k + 1] = y[k]; }
  x[at] = lon; y[at] = lat;
  return n + 1;
}

// everything of fix_lon that does not depend on tlon: pole vertices, twin poles, unwrap (mosaic_util.c:672-725); *sum_out = x_sum
__device__ inline int fix_lon_unshifted(double* x, double* y, int n, double* sum_out) {
  const double near_pole = kHalfPi - kPoleTol;
  int nn = n;
  *sum_out = 0.0;
  bool any_pole = false;
  for (int i = 0; i < nn; ++i) any_pole |= (fabs(y[i]) >= near_pole);
  if (any_pole) {
    // pole vertices must come in pairs (mosaic_util.c:679-692)
    for (int i = 0; i < nn; ++i) {
      if (fabs(y[i]) >= near_pole) {
        int prev = (i + nn - 1) % nn, next = (i + 1) % nn;
        if (y[prev] == y[i] && y[next] == y[i]) { nn = vtx_remove(x, y, nn, i); --i; }
        else if (y[prev] != y[i] && y[next] != y[i]) {
          if (nn >= kMaxV + 2) return -1;
          nn = vtx_insert(x, y, nn, i, x[i], y[i]); ++i;
        }
      }
    }
    // first/second of a pole pair take the neighbouring longitudes (:693-700)
    for (int i = 0; i < nn; ++i) {
      if (fabs(y[i]) >= near_pole) {
        int prev = (i + nn - 1) % nn, next = (i + 1) % nn;
        if (y[prev] != y[i]) x[i] = x[prev];
        if (y[next] != y[i]) x[i] = x[next];
      }
    }
  }
  // a side through a pole gets twin pole vertices (:702-717)
  for (int i = 0; i < nn; ++i) {
    int prev = (i + nn - 1) % nn;
    double d = x[i] - x[prev];
    if (fabs(d + kPi) < kSmall || fabs(d - kPi) < kSmall) {
      if (nn + 2 > kMaxV + 2) return -1;
      double xa = x[prev], xb = x[i];
      double yp = (y[i] < 0.0) ? -kHalfPi : kHalfPi;
      nn = vtx_insert(x, y, nn, i, xb, yp);
      nn = vtx_insert(x, y, nn, i, xa, yp);
      break;
    }
  }
  if (nn == 0) return 0;
  // unwrap (:718-725)
  double sum = x[0];
  for (int i = 1; i < nn; ++i) {
    double d = x[i] - x[i - 1];
    if (d < -kPi) d = d + kTwoPi;
    else if (d > kPi) d = d - kTwoPi;
    x[i] = x[i - 1] + d;
    sum += x[i];
  }
  *sum_out = sum;
  return nn;
}

__device__ inline int fix_lon(double* x, double* y, int n, double tlon) {
  double sum;
  const int nn = fix_lon_unshifted(x, y, n, &sum);
  if (nn <= 0) return nn;
  // mean within pi of tlon (:727-729)
  double shift = (sum / nn) - tlon;
  if (shift < -kPi)      { for (int i = 0; i < nn; ++i) x[i] += kTwoPi; }
  else if (shift > kPi)  { for (int i = 0; i < nn; ++i) x[i] -= kTwoPi; }
  return nn;
}

// ---------------------------------------------------------------------------------------------
// Strided polygon view: vertex k lives at p[k * stride].  stride = blockDim.x for the
// shared-memory staging used by the clip kernel (bank-conflict free), 1 for private arrays.
// ---------------------------------------------------------------------------------------------
struct PolyView {
  const double* x; const double* y; int stride;
  XGB_HD double X(int k) const { return x[k * stride]; }
  XGB_HD double Y(int k) const { return y[k * stride]; }
};

// poly_area_main (mosaic_util.c:417-459) -> m^2
__device__ __forceinline__ double poly_area(const PolyView& p, int n) {
  double acc = 0.0;
  double xi = p.X(0), yi = p.Y(0);
  const double x0 = xi, y0 = yi;
  for (int i = 0; i < n; ++i) {
    double xn, yn;
    if (i + 1 < n) { xn = p.X(i + 1); yn = p.Y(i + 1); } else { xn = x0; yn = y0; }
    double dx = xn - xi;
    double lat1 = yn, lat2 = yi;
    if (dx > kPi)  dx = dx - 2.0 * kPi;
    if (dx < -kPi) dx = dx + 2.0 * kPi;
    if (fabs(dx + kPi) < kSmall || fabs(dx - kPi) < kSmall) {
      acc += kPi;                                  // side through a pole (:434-437)
    } else if (fabs(lat1 - lat2) < kSmall) {
      acc -= dx * ref_sin(0.5 * (lat1 + lat2));
    } else {
      double dy = 0.5 * (lat1 - lat2);
      double dat = ref_sin(dy) / dy;
      acc -= dx * ref_sin(0.5 * (lat1 + lat2)) * dat;
    }
    xi = xn; yi = yn;
  }
  return (acc < 0) ? -acc * kRadius * kRadius : acc * kRadius * kRadius;
}

// poly_ctrlat (create_xgrid.c:2096-2121)
__device__ __forceinline__ double poly_ctrlat(const PolyView& p, int n) {
  double acc = 0.0;
  double xi = p.X(0), yi = p.Y(0);
  const double x0 = xi, y0 = yi;
  for (int i = 0; i < n; ++i) {
    double xn, yn;
    if (i + 1 < n) { xn = p.X(i + 1); yn = p.Y(i + 1); } else { xn = x0; yn = y0; }
    double dx = xn - xi;
    double lat1 = yn, lat2 = yi;
    double dy = lat2 - lat1;
    double hdy = dy * 0.5;
    double avg_y = (lat1 + lat2) * 0.5;
    if (dx != 0.0) {
      if (dx > kPi)   dx = dx - 2.0 * kPi;
      if (dx <= -kPi) dx = dx + 2.0 * kPi;
      // the reference binary evaluates cos(avg_y)/sin(avg_y) with one sincos() call, cos(lat1) and
      // sin(hdy) with cos()/sin() (gcc -O2); ref_* reproduce exactly those entry points
      double sa, ca;
      ref_sincos(avg_y, &sa, &ca);
      if (fabs(hdy) < kSmall)
        acc -= dx * (2 * ca + lat2 * sa - ref_cos(lat1));
      else
        acc -= dx * ((ref_sin(hdy) / hdy) * (2 * ca + lat2 * sa) - ref_cos(lat1));
    }
    xi = xn; yi = yn;
  }
  return acc * kRadius * kRadius;
}

// poly_ctrlon (create_xgrid.c:2170-2217)
__device__ __forceinline__ double poly_ctrlon(const PolyView& p, int n, double clon) {
  double acc = 0.0;
  double xi = p.X(0), yi = p.Y(0);
  const double x0 = xi, y0 = yi;
  for (int i = 0; i < n; ++i) {
    double xn, yn;
    if (i + 1 < n) { xn = p.X(i + 1); yn = p.Y(i + 1); } else { xn = x0; yn = y0; }
    double phi1 = xn, phi2 = xi, lat1 = yn, lat2 = yi;
    double dphi = phi1 - phi2;
    if (dphi != 0.0) {
      double s1, c1, s2, c2;
      ref_sincos(lat1, &s1, &c1);
      ref_sincos(lat2, &s2, &c2);
      double f1 = 0.5 * (c1 * s1 + lat1);
      double f2 = 0.5 * (c2 * s2 + lat2);
      if (dphi > kPi)  dphi = dphi - 2.0 * kPi;
      if (dphi < -kPi) dphi = dphi + 2.0 * kPi;
      double dphi1 = phi1 - clon;
      if (dphi1 > kPi)  dphi1 -= 2.0 * kPi;
      if (dphi1 < -kPi) dphi1 += 2.0 * kPi;
      double dphi2 = phi2 - clon;
      if (dphi2 > kPi)  dphi2 -= 2.0 * kPi;
      if (dphi2 < -kPi) dphi2 += 2.0 * kPi;
      if (fabs(dphi2 - dphi1) < kPi) {
        acc -= dphi * (dphi1 * f1 + dphi2 * f2) / 2.0;
      } else {
        double fac = (dphi1 > 0.0) ? kPi : -kPi;
        double fint = f1 + (f2 - f1) * (fac - dphi1) / fabs(dphi);
        acc -= 0.5 * dphi1 * (dphi1 - fac) * f1 - 0.5 * dphi2 * (dphi2 + fac) * f2 + 0.5 * fac * (dphi1 + dphi2) * fint;
      }
    }
    xi = xn; yi = yn;
  }
  return acc * kRadius * kRadius;
}

// ---------------------------------------------------------------------------------------------
// poly_area, poly_ctrlon and poly_ctrlat of one polygon in a single pass over its edges.
// Each of the three sums accumulates exactly the terms of its reference routine, in the same
// order; what is shared is only trig that is provably the same number:
//   * sincos(lat) of a vertex serves the edge before and the edge after it (the reference
//     evaluates it twice) and doubles as poly_ctrlat's cos(lat1) (cos() and sincos() agree bit
//     for bit: both are do_cos / do_sin of the same reduced argument, see ref_trig.cuh);
//   * poly_area's sin(dy), dy = (lat1-lat2)/2, is -sin(hdy) of poly_ctrlat (odd function, exact), so
//     sin(dy)/dy == sin(hdy)/hdy;
//   * poly_area's sin((lat1+lat2)/2) equals the sin half of poly_ctrlat's sincos for |x| < 0.855469
//     (both are do_sin(x, 0)); above that sin() takes another path and is evaluated separately.
// ORDER == 1 evaluates poly_area only.
// ---------------------------------------------------------------------------------------------
// WARP: called by all 32 lanes of a converged warp (lanes without a polygon pass n = 0).  The edge loop then runs to the
// warp's largest vertex count and the lanes are re-converged explicitly after every data-dependent section (the
// sin/cos routines branch on the argument range), which the compiler does not do on its own across the inlined code.
// LEAN: the trig goes through ref_trig_site / ref_sin_small (ref_trig.cuh): one reduction and one sin-type + one cos-type
// evaluation serve both table ranges, the large-argument reduction (never taken by latitudes) becomes a call, and sin of
// the half-difference takes its Taylor path inline — same bits, about a third of the instructions per site.  The clip
// kernel is bound by instruction fetch (profiles/r01*), so this is worth more than the arithmetic it saves.
template <int ORDER, bool WARP = false, bool LEAN = false>
__device__ __forceinline__ void poly_moments(const PolyView& p, int n, double clon,
                                             double* area_out, double* ctrlon_out, double* ctrlat_out,
                                             const double* T = nullptr) {
  auto sincos_ = [&](double a, double* s, double* c) {
    if (LEAN) ref_trig_site(a, false, s, c, T); else ref_sincos(a, s, c);
  };
  auto sin_half_ = [&](double a) { return LEAN ? ref_sin_small(a) : ref_sin(a); };
  auto sin_own_ = [&](double a) {
    if (!LEAN) return ref_sin(a);
    double s, c;
    ref_trig_site(a, true, &s, &c, T);
    return s;
  };
  double aacc = 0.0, lonacc = 0.0, latacc = 0.0;
  double xi = 0.0, yi = 0.0;
  if (n > 0) { xi = p.X(0); yi = p.Y(0); }
  const double x0 = xi, y0 = yi;
  double si = 0.0, ci = 0.0;
  if (ORDER == 2 && n > 0) sincos_(yi, &si, &ci);
  int nloop = n;
  if (WARP) {
#if defined(__CUDA_ARCH__)
    __syncwarp();
    nloop = __reduce_max_sync(0xffffffffu, n);
#endif
  }
  const double s0 = si, c0 = ci;
  for (int i = 0; i < nloop; ++i) {
    const bool act = !WARP || i < n;
    double xn = x0, yn = y0, sn = s0, cn = c0;
    if (act && i + 1 < n) { xn = p.X(i + 1); yn = p.Y(i + 1); if (ORDER == 2) sincos_(yn, &sn, &cn); }
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp();
#endif
    const double lat1 = yn, lat2 = yi;
    const double dx_raw = xn - xi;                       // x[ip]-x[i] == phi1-phi2
    double dxa = dx_raw;                                 // poly_area's wrapped dx (mosaic_util.c:429-432)
    if (dxa > kPi)  dxa = dxa - 2.0 * kPi;
    if (dxa < -kPi) dxa = dxa + 2.0 * kPi;
    const bool pole_edge = (fabs(dxa + kPi) < kSmall || fabs(dxa - kPi) < kSmall);
    const bool flat_area = (fabs(lat1 - lat2) < kSmall);
    const double avg = 0.5 * (lat1 + lat2);
    const double dy = 0.5 * (lat1 - lat2);               // hdy of poly_ctrlat is -dy
    const bool moving = (ORDER == 2) && (dx_raw != 0.0); // poly_ctrlon / poly_ctrlat skip dx == 0 edges
    const bool flat_lat = (fabs(dy) < kSmall);           // fabs(hdy) < SMALL_VALUE (create_xgrid.c:2114)

    double s_avg = 0.0, c_avg = 0.0;
    if (act && moving) sincos_(avg, &s_avg, &c_avg);
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp();
#endif
    double dat = 0.0;
    if (act && ((!pole_edge && !flat_area) || (moving && !flat_lat))) dat = sin_half_(dy) / dy;
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp();
#endif
    const uint32_t hi = (uint32_t)(trig::bits(avg) >> 32) & 0x7fffffffu;
    const bool own_sin = act && !pole_edge && !(moving && hi < 0x3feb6000u);
    double sin_avg = s_avg;
    if (own_sin) sin_avg = sin_own_(avg);
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp();
#endif
    if (act) {
      if (pole_edge) {
        aacc += kPi;                                     // mosaic_util.c:434-437
      } else {
        if (flat_area) aacc -= dxa * sin_avg;
        else           aacc -= dxa * sin_avg * dat;
      }
      if (moving) {
        // poly_ctrlat (create_xgrid.c:2100-2118)
        double dxl = dx_raw;
        if (dxl > kPi)   dxl = dxl - 2.0 * kPi;
        if (dxl <= -kPi) dxl = dxl + 2.0 * kPi;
        if (flat_lat) latacc -= dxl * (2 * c_avg + lat2 * s_avg - cn);
        else          latacc -= dxl * (dat * (2 * c_avg + lat2 * s_avg) - cn);
        // poly_ctrlon (create_xgrid.c:2176-2215)
        const double f1 = 0.5 * (cn * sn + lat1);
        const double f2 = 0.5 * (ci * si + lat2);
        double dphi = dx_raw;
        if (dphi > kPi)  dphi = dphi - 2.0 * kPi;
        if (dphi < -kPi) dphi = dphi + 2.0 * kPi;
        double dphi1 = xn - clon;
        if (dphi1 > kPi)  dphi1 -= 2.0 * kPi;
        if (dphi1 < -kPi) dphi1 += 2.0 * kPi;
        double dphi2 = xi - clon;
        if (dphi2 > kPi)  dphi2 -= 2.0 * kPi;
        if (dphi2 < -kPi) dphi2 += 2.0 * kPi;
        if (fabs(dphi2 - dphi1) < kPi) {
          lonacc -= dphi * (dphi1 * f1 + dphi2 * f2) / 2.0;
        } else {
          const double fac = (dphi1 > 0.0) ? kPi : -kPi;
          const double fint = f1 + (f2 - f1) * (fac - dphi1) / fabs(dphi);
          lonacc -= 0.5 * dphi1 * (dphi1 - fac) * f1 - 0.5 * dphi2 * (dphi2 + fac) * f2 + 0.5 * fac * (dphi1 + dphi2) * fint;
        }
      }
      xi = xn; yi = yn; si = sn; ci = cn;
    }
#if defined(__CUDA_ARCH__)
    if (WARP) __syncwarp();
#endif
  }
  *area_out = (aacc < 0) ? -aacc * kRadius * kRadius : aacc * kRadius * kRadius;
  if (ORDER == 2) { *ctrlon_out = lonacc * kRadius * kRadius; *ctrlat_out = latacc * kRadius * kRadius; }
}

// ---------------------------------------------------------------------------------------------
// The same three sums with ONE trig evaluation site (ref_trig_site) and rolled loops: the clip kernel is bound by
// instruction fetch, and poly_moments above inlines five copies of the trig range paths.  Edge i runs up to three
// evaluations through the site: sincos(lat of the next vertex) [order 2], sincos(avg) [moving edges, order 2] and sin(avg)
// with ref_sin's own reduction where poly_area's sin() differs from sincos's (see above).  Iteration -1 only evaluates
// vertex 0.  Every accumulated term is the expression of poly_moments, so the results are bit-identical to it
// (tests/test_capi_cpu.py pins this routine against the oracle's poly_area / poly_ctrlon / poly_ctrlat on the host).
// T = the 440-entry sin/cos table (shared memory in the kernel).
// ---------------------------------------------------------------------------------------------
template <int ORDER>
XGB_HD void poly_moments_site(const PolyView& p, int n, double clon, const double* T,
                              double* area_out, double* ctrlon_out, double* ctrlat_out) {
  double aacc = 0.0, lonacc = 0.0, latacc = 0.0;
  double xi = 0.0, yi = 0.0, si = 0.0, ci = 0.0, s0 = 0.0, c0 = 0.0;
#pragma unroll 1
  for (int i = -1; i < n; ++i) {
    const bool edge = (i >= 0);
    const bool fresh = (i + 1 < n);                      // the next vertex is not vertex 0 again
    const int kn = fresh ? i + 1 : 0;
    const double xn = p.X(kn), yn = p.Y(kn);
    const double lat1 = yn, lat2 = yi;
    const double dx_raw = xn - xi;                       // x[ip]-x[i] == phi1-phi2
    double dxa = dx_raw;                                 // poly_area's wrapped dx (mosaic_util.c:429-432)
    if (dxa > kPi)  dxa = dxa - 2.0 * kPi;
    if (dxa < -kPi) dxa = dxa + 2.0 * kPi;
    const bool pole_edge = (fabs(dxa + kPi) < kSmall || fabs(dxa - kPi) < kSmall);
    const bool flat_area = (fabs(lat1 - lat2) < kSmall);
    const double avg = 0.5 * (lat1 + lat2);
    const double dy = 0.5 * (lat1 - lat2);               // hdy of poly_ctrlat is -dy
    const bool moving = (ORDER == 2) && edge && (dx_raw != 0.0);
    const bool flat_lat = (fabs(dy) < kSmall);
    const uint32_t hi = (uint32_t)(trig::bits(avg) >> 32) & 0x7fffffffu;
    const bool own_sin = edge && !pole_edge && !(moving && hi < 0x3feb6000u);
    double sn = s0, cn = c0, s_avg = 0.0, c_avg = 0.0, sin_avg = 0.0;
#pragma unroll 1
    for (int t = (ORDER == 2) ? 0 : 2; t < 3; ++t) {
      const bool doit = (t == 0) ? ((ORDER == 2) && fresh) : (t == 1) ? moving : own_sin;
      if (doit) {
        double s_, c_;
        ref_trig_site((t == 0) ? yn : avg, t == 2, &s_, &c_, T);
        if (t == 0) { sn = s_; cn = c_; }
        else if (t == 1) { s_avg = s_; c_avg = c_; }
        else sin_avg = s_;
      }
    }
    if (!own_sin) sin_avg = s_avg;
    if (edge) {
      double dat = 0.0;
      if ((!pole_edge && !flat_area) || (moving && !flat_lat)) dat = ref_sin_small(dy) / dy;
      if (pole_edge) {
        aacc += kPi;                                     // mosaic_util.c:434-437
      } else {
        if (flat_area) aacc -= dxa * sin_avg;
        else           aacc -= dxa * sin_avg * dat;
      }
      if (moving) {
        // poly_ctrlat (create_xgrid.c:2100-2118)
        double dxl = dx_raw;
        if (dxl > kPi)   dxl = dxl - 2.0 * kPi;
        if (dxl <= -kPi) dxl = dxl + 2.0 * kPi;
        if (flat_lat) latacc -= dxl * (2 * c_avg + lat2 * s_avg - cn);
        else          latacc -= dxl * (dat * (2 * c_avg + lat2 * s_avg) - cn);
        // poly_ctrlon (create_xgrid.c:2176-2215)
        const double f1 = 0.5 * (cn * sn + lat1);
        const double f2 = 0.5 * (ci * si + lat2);
        double dphi = dx_raw;
        if (dphi > kPi)  dphi = dphi - 2.0 * kPi;
        if (dphi < -kPi) dphi = dphi + 2.0 * kPi;
        double dphi1 = xn - clon;
        if (dphi1 > kPi)  dphi1 -= 2.0 * kPi;
        if (dphi1 < -kPi) dphi1 += 2.0 * kPi;
        double dphi2 = xi - clon;
        if (dphi2 > kPi)  dphi2 -= 2.0 * kPi;
        if (dphi2 < -kPi) dphi2 += 2.0 * kPi;
        if (fabs(dphi2 - dphi1) < kPi) {
          lonacc -= dphi * (dphi1 * f1 + dphi2 * f2) / 2.0;
        } else {
          const double fac = (dphi1 > 0.0) ? kPi : -kPi;
          const double fint = f1 + (f2 - f1) * (fac - dphi1) / fabs(dphi);
          lonacc -= 0.5 * dphi1 * (dphi1 - fac) * f1 - 0.5 * dphi2 * (dphi2 + fac) * f2 + 0.5 * fac * (dphi1 + dphi2) * fint;
        }
      }
    } else { s0 = sn; c0 = cn; }
    xi = xn; yi = yn; si = sn; ci = cn;
  }
  *area_out = (aacc < 0) ? -aacc * kRadius * kRadius : aacc * kRadius * kRadius;
  if (ORDER == 2) { *ctrlon_out = lonacc * kRadius * kRadius; *ctrlat_out = latacc * kRadius * kRadius; }
}

// inside_edge (create_xgrid.c:2342-2350)
__device__ __forceinline__ bool inside_edge(double x0, double y0, double x1, double y1, double x, double y) {
  double product = (x - x0) * (y1 - y0) + (x0 - x1) * (y - y0);
  return product <= 1.e-12;
}

}  // namespace xgb
