// sin / cos / sincos that round exactly like the host libm the reference is linked against.
//
// Why: the reference's poly_area (mosaic_util.c:417-459) is a line integral that cancels
// catastrophically for small exchange cells — terms of size dlon*sin(lat) add up to a result of
// size dlon*dlat*cos(lat) — so a 1-ulp difference in sin() shows up as 1e-12..1e-9 relative in
// xgrid_area and far worse in tile1_distance.  Matching the reference to its stated 1e-12 therefore
// needs the same sin/cos bits, not merely an accurate sin/cos.
//
// What: the algorithm of the IBM Accurate Mathematical Library as built into glibc 2.39 for x86-64
// (sysdeps/ieee754/dbl-64/s_sin.c: do_sin / do_cos / TAYLOR_SIN over a 1/128-spaced table), with the
// multiply-add contractions of its FMA build (__sin_fma / __cos_fma / __sincos_fma, the variant the
// dynamic linker selects on every FMA-capable x86-64 CPU) written out as explicit fma() calls.
// |x| < 2.426265 (all latitudes and half-differences the 2dx2d path feeds to trig) takes the table paths;
// 2.426265 <= |x| < 105414350 (longitudes, great-circle path) takes reduce_sincos + do_sincos; beyond that the
// toolchain's sin/cos is used (never reached by regridding inputs).
//
// The same code compiles for host and device; tests/test_trig_cpu.py checks the host build
// bit-for-bit against libm on millions of arguments, and tests/test_xgrid_gpu.py checks the device
// build against the host build.  If a different libm is in use the regridding results still agree
// to rounding level, just not bit-for-bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define XGB_HD __host__ __device__ __forceinline__
#else
#define XGB_HD inline
#endif

namespace xgb {

#if defined(__CUDACC__)
static __device__ __align__(16) const double kSinCosTabDev[440] = {
#include "sincostab.inc"
};
#endif
static const double kSinCosTabHost[440] = {
#include "sincostab.inc"
};

namespace trig {
// The constants of s_sin.c.  Device code reads them from constant memory (XGB_K): an FP64 instruction takes a constant-bank
// operand for free, while a 64-bit literal costs two UMOV / IMAD.MOV to materialise each time it is used — a tenth of all
// instructions of the moments kernel before this (ncu r02t: UMOV 10.4 %, IMAD 13.8 % of the warp instructions).
#define XGB_TRIG_CONSTS(X)                                                                              \
  X(big, 0x1.8000000000000p+45)                                                                         \
  X(hp0, 0x1.921fb54442d18p+0)   /* pi/2 high part */                                                   \
  X(hp1, 0x1.1a62633145c07p-54)  /* pi/2 low part */                                                    \
  X(s1, -0x1.5555555555555p-3) X(s2, 0x1.1111111110ecep-7) X(s3, -0x1.a01a019db08b8p-13)                \
  X(s4, 0x1.71de27b9a7ed9p-19) X(s5, -0x1.addffc2fcdf59p-26)                                            \
  X(sn3, -0x1.5555555555515p-3) X(sn5, 0x1.11110e829872fp-7)                                            \
  X(cs4, -0x1.5555555555535p-5) X(cs6, 0x1.6c16bedd9e239p-10)                                           \
  X(hpinv, 0x1.45F306DC9C883p-1) X(toint, 0x1.8000000000000p52)                                         \
  X(mp1, 0x1.921FB58000000p0) X(mp2, -0x1.DDE973C000000p-27)                                            \
  X(pp3, -0x1.CB3B398000000p-55) X(pp4, -0x1.d747f23e32ed7p-83)                                         \
  X(c126, 0.126) X(c01588, 0.01588)
#define XGB_X_HOST(name, value) constexpr double name = value;
XGB_TRIG_CONSTS(XGB_X_HOST)
#undef XGB_X_HOST
constexpr double cs2 = 0.5;
#if defined(__CUDACC__)
#define XGB_X_DEV(name, value) static __constant__ double dev_##name = value;
XGB_TRIG_CONSTS(XGB_X_DEV)
#undef XGB_X_DEV
#endif
#if defined(__CUDA_ARCH__) && !defined(XGB_LITERAL_CONSTS)
#define XGB_K(name) (::xgb::trig::dev_##name)
#else
#define XGB_K(name) (::xgb::trig::name)
#endif

XGB_HD uint64_t bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
XGB_HD double tab(int i) {
#if defined(__CUDA_ARCH__)
  return __ldg(&kSinCosTabDev[i]);
#else
  return kSinCosTabHost[i];
#endif
}
XGB_HD double mag_with_sign_of(double mag, double sgn) { return copysign(fabs(mag), sgn); }

// TAYLOR_SIN(x*x, x, dx)
XGB_HD double taylor_sin(double x, double dx) {
  const double xx = x * x;
  double p = fma(xx, XGB_K(s5), XGB_K(s4));
  p = fma(xx, p, XGB_K(s3));
  p = fma(xx, p, XGB_K(s2));
  p = fma(xx, p, XGB_K(s1));
  const double h = dx * 0.5;
  double w = fma(p, x, -h);
  w = fma(xx, w, dx);
  return x + w;
}

struct Tab { double sn, ssn, cs, ccs; };
// u = big + |x| puts round(|x|*128) in the low word; returns |x| - k/128 and loads row k
XGB_HD double reduce(double ax, Tab* t) {
  const double u = ax + XGB_K(big);
  const int k4 = (int)((uint32_t)bits(u) << 2);
  t->sn = tab(k4); t->ssn = tab(k4 + 1); t->cs = tab(k4 + 2); t->ccs = tab(k4 + 3);
  return ax - (u - XGB_K(big));
}

XGB_HD double sin_core(double xr, double dx, const Tab& t) {       // do_sin after reduction
  const double xx = xr * xr;
  const double p = fma(xx, XGB_K(sn5), XGB_K(sn3));
  const double s = xr + fma(xr * xx, p, dx);
  double q = fma(xx, XGB_K(cs6), XGB_K(cs4));
  q = fma(q, xx, cs2);
  const double c = fma(xr, dx, xx * q);
  double e = fma(s, t.ccs, t.ssn);
  e = fma(-c, t.sn, e);
  const double cor = fma(s, t.cs, e);
  return t.sn + cor;
}

XGB_HD double cos_core(double xr, const Tab& t) {                  // do_cos after reduction (dx already added)
  const double xx = xr * xr;
  const double p = fma(xx, XGB_K(sn5), XGB_K(sn3));
  const double s = fma(xr * xx, p, xr);
  double q = fma(xx, XGB_K(cs6), XGB_K(cs4));
  q = fma(q, xx, cs2);
  const double c = xx * q;
  double e = fma(-s, t.ssn, t.ccs);
  e = fma(-c, t.cs, e);
  const double cor = fma(-s, t.sn, e);
  return t.cs + cor;
}

XGB_HD double do_sin(double x, double dx) {
  const double ax = fabs(x);
  if (ax < XGB_K(c126)) return taylor_sin(x, dx);
  if (x <= 0) dx = -dx;
  Tab t;
  const double xr = reduce(ax, &t);
  return mag_with_sign_of(sin_core(xr, dx, t), x);
}

XGB_HD double do_cos(double x, double dx) {
  if (x < 0) dx = -dx;
  Tab t;
  const double xr = reduce(fabs(x), &t) + dx;
  return cos_core(xr, t);
}
// reduce_sincos (s_sin.c): x = n*pi/2 + (a + da), |a| <= pi/4.  Only the y step is contracted in the FMA build
// (checked against this image's libm on 4e8 arguments; contracting the XGB_K(pp4) step as well is indistinguishable).

XGB_HD int reduce_sincos(double x, double* a, double* da) {
  const double t = x * XGB_K(hpinv) + XGB_K(toint);
  const double xn = t - XGB_K(toint);
  const int n = (int)(bits(t) & 3u);
  const double y = fma(-xn, XGB_K(mp2), fma(-xn, XGB_K(mp1), x));
  double t1 = xn * XGB_K(pp3);
  const double t2 = y - t1;
  double db = (y - t2) - t1;
  t1 = xn * XGB_K(pp4);
  const double b = t2 - t1;
  db += (t2 - b) - t1;
  *a = b; *da = db;
  return n;
}

XGB_HD double do_sincos(double a, double da, int n) {
  double r;
  if (n & 1) r = do_cos(a, da);
  else {
    const double xx = a * a;
    if (xx < XGB_K(c01588)) r = taylor_sin(a, da);
    else r = mag_with_sign_of(do_sin(a, da), a);
  }
  return (n & 2) ? -r : r;
}
}  // namespace trig

XGB_HD double ref_sin(double x) {
  const uint32_t k = (uint32_t)(trig::bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;
  if (k < 0x3feb6000u) return trig::do_sin(x, 0.0);
  if (k < 0x400368fdu) return trig::mag_with_sign_of(trig::do_cos(XGB_K(hp0) - fabs(x), XGB_K(hp1)), x);
  if (k < 0x419921FBu) { double a, da; const int n = trig::reduce_sincos(x, &a, &da); return trig::do_sincos(a, da, n); }
  return sin(x);
}

XGB_HD double ref_cos(double x) {
  const uint32_t k = (uint32_t)(trig::bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;
  if (k < 0x3feb6000u) return trig::do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = XGB_K(hp0) - fabs(x);
    const double a = y + XGB_K(hp1);
    const double da = (y - a) + XGB_K(hp1);
    return trig::do_sin(a, da);
  }
  if (k < 0x419921FBu) { double a, da; const int n = trig::reduce_sincos(x, &a, &da); return trig::do_sincos(a, da, n + 1); }
  return cos(x);
}

XGB_HD void ref_sincos(double x, double* sn, double* cs) {
  const uint32_t k = (uint32_t)(trig::bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) { *sn = x; *cs = 1.0; return; }
  if (k < 0x3feb6000u) {
    const double ax = fabs(x);
    trig::Tab t;
    const double xr0 = trig::reduce(ax, &t);
    // sin: do_sin(x, 0); cos: do_cos(x, 0) — both share one table row
    if (ax < XGB_K(c126)) *sn = trig::taylor_sin(x, 0.0);
    else *sn = trig::mag_with_sign_of(trig::sin_core(xr0, (x > 0) ? 0.0 : -0.0, t), x);
    *cs = trig::cos_core(xr0 + ((x >= 0) ? 0.0 : -0.0), t);
    return;
  }
  if (k < 0x400368fdu) {
    const double y = XGB_K(hp0) - fabs(x);
    const double a = y + XGB_K(hp1);
    const double da = (y - a) + XGB_K(hp1);
    const double aa = fabs(a);
    trig::Tab t;
    const double xr0 = trig::reduce(aa, &t);
    *sn = trig::mag_with_sign_of(trig::cos_core(xr0 + ((a < 0) ? -da : da), t), x);
    if (aa < XGB_K(c126)) *cs = trig::taylor_sin(a, da);
    else *cs = trig::mag_with_sign_of(trig::sin_core(xr0, (a <= 0) ? -da : da, t), a);
    return;
  }
  if (k < 0x419921FBu) {
    double a, da;
    const int n = trig::reduce_sincos(x, &a, &da);
    *sn = trig::do_sincos(a, da, n); *cs = trig::do_sincos(a, da, n + 1);
    return;
  }
  *sn = sin(x); *cs = cos(x);
}

// ---------------------------------------------------------------------------------------------
// One evaluation site for everything the clip kernel asks of trig (the kernel is instruction-fetch bound: five inlined
// copies of the three range paths cost more than the arithmetic).  T is the 440-entry table (shared-memory copy in the
// kernel, host table in the CPU tests).
//   sin_only == false: (*sn, *cs) = ref_sincos(x)
//   sin_only == true : *sn = ref_sin(x), *cs unspecified
// Both table ranges reduce to "sin-type" and "cos-type" evaluations of one reduced argument (a, da):
//   |x| < 0.855469        a = x, da = 0:                  sin = sin-type, cos = cos-type           (ref_sin == ref_sincos.sin)
//   0.855469 <= |x| < 2.426265   y = pi/2 - |x|:          sin = +-cos-type, cos = sin-type, with
//        ref_sincos: a = y + XGB_K(hp1), da = (y - a) + XGB_K(hp1);     ref_sin: a = y, da = XGB_K(hp1)  (do_cos(y, XGB_K(hp1)))
// The operation order of every path is that of ref_sin / ref_sincos above; tests pin all three bit for bit against libm.
// ---------------------------------------------------------------------------------------------
namespace trig {
XGB_HD double reduce_t(double ax, Tab* t, const double* T) {
  const double u = ax + XGB_K(big);
  const int k4 = (int)((uint32_t)bits(u) << 2);
#if defined(__CUDA_ARCH__)
  const double2 r0 = *reinterpret_cast<const double2*>(T + k4);       // rows are 32 bytes: two 16-byte loads
  const double2 r1 = *reinterpret_cast<const double2*>(T + k4 + 2);
  t->sn = r0.x; t->ssn = r0.y; t->cs = r1.x; t->ccs = r1.y;
#else
  t->sn = T[k4]; t->ssn = T[k4 + 1]; t->cs = T[k4 + 2]; t->ccs = T[k4 + 3];
#endif
  return ax - (u - XGB_K(big));
}
}  // namespace trig

#if defined(__CUDACC__)
static __device__ __noinline__ void ref_sincos_call(double x, double* sn, double* cs) { ref_sincos(x, sn, cs); }
static __device__ __noinline__ double ref_sin_call(double x) { return ref_sin(x); }
#endif

XGB_HD void ref_trig_site(double x, bool sin_only, double* sn, double* cs, const double* T) {
  const uint32_t k = (uint32_t)(trig::bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) { *sn = x; *cs = 1.0; return; }
  if (k >= 0x400368fdu) {                       // never reached by latitudes
#if defined(__CUDA_ARCH__)
    if (sin_only) *sn = ref_sin_call(x); else ref_sincos_call(x, sn, cs);
#else
    if (sin_only) *sn = ref_sin(x); else ref_sincos(x, sn, cs);
#endif
    return;
  }
  const bool swap = (k >= 0x3feb6000u);
  double a = x, da = 0.0;
  if (swap) {
    const double y = XGB_K(hp0) - fabs(x);
    if (sin_only) { a = y; da = XGB_K(hp1); }
    else { a = y + XGB_K(hp1); da = (y - a) + XGB_K(hp1); }
  }
  const double aa = fabs(a);
  trig::Tab t;
  const double xr0 = trig::reduce_t(aa, &t, T);
  double S = 0.0, Cc = 0.0;
  if (!(swap && sin_only)) {                    // sin-type of (a, da)
    if (aa < XGB_K(c126)) S = trig::taylor_sin(a, da);
    else S = trig::mag_with_sign_of(trig::sin_core(xr0, (a <= 0) ? -da : da, t), a);
  }
  if (swap || !sin_only)                        // cos-type of (a, da)
    Cc = trig::cos_core(xr0 + ((a < 0) ? -da : da), t);
  if (swap) { *sn = trig::mag_with_sign_of(Cc, x); *cs = S; }
  else      { *sn = S; *cs = Cc; }
}

// sin(x) for the half-differences of latitudes (|x| < 0.126 on any realistic grid): Taylor path inline, the rest by call
XGB_HD double ref_sin_small(double x) {
  const uint32_t k = (uint32_t)(trig::bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;
  if (fabs(x) < XGB_K(c126)) return trig::taylor_sin(x, 0.0);
#if defined(__CUDA_ARCH__)
  return ref_sin_call(x);
#else
  return ref_sin(x);
#endif
}

XGB_HD const double* ref_trig_table() {
#if defined(__CUDA_ARCH__)
  return kSinCosTabDev;
#else
  return kSinCosTabHost;
#endif
}

}  // namespace xgb
