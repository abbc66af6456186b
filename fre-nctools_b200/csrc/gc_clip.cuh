// Great-circle polygon clipping and spherical-excess area for the exchange-grid kernels (sm_100a; the same code
// also compiles for the host so tests can run it against the compiled reference on the CPU).
//
// Replaces clip_2dx2d_great_circle (reference create_xgrid.c:1479-1908), line_intersect_2D_3D (:1919-2081),
// intersect_tri_with_line / invert_matrix_3x3 (mosaic_util.c:967-1043), insidePolygon (:1487-1530),
// great_circle_area / spherical_angle (:763-838) and the list primitives addEnd / addIntersect / insertIntersect /
// setInbound / getFirstInbound (:1095-1470).
//
// The reference walks singly linked lists carved from a global 100-node pool; here each thread keeps its four
// lists as small fixed arrays in local memory (two vertex rings of at most kGcRing nodes, the intersection list, the
// output polygon) and an "insert after" is an array insertion.  Every decision (EPSLN8 snapping of the edge
// parameters, EPSLN10 point identity, exact-coordinate node identity, inbound/outbound bookkeeping, the walk
// itself) is restated one for one: which vertex ends up where decides the integer cell lists.
//
// Two places cannot be bit-identical to an x86-64 build of the reference, which evaluates them in x87 extended
// precision: the 3x3 solve of intersect_tri_with_line (long double) and acosl() in spherical_angle.  The solve is
// done here in double-double arithmetic (106 bits, so the rounded result equals the x87 one except when the x87
// result itself sits within 2^-11 ulp of a rounding boundary); acos uses gc_acos below.  Areas therefore agree with the
// reference to a few 1e-16 steradian (absolute: the spherical excess is a difference of O(1) angles), the integer
// lists exactly.
#pragma once
#include <math.h>
#include "ref_trig.cuh"

// The great-circle kernel is bound by instruction fetch (ncu r02v: 11 488 SASS instructions, "no instruction" the top stall by
// 6x, 14 of 32 lanes active): the large routines are calls on the device, one copy each, instead of inlined at every site.
#if defined(__CUDACC__)
#define XGB_HD_CALL static __host__ __device__ __noinline__
#else
#define XGB_HD_CALL static inline
#endif

namespace xgb {
namespace gc {

constexpr double kRange = 0.05;     // RANGE_CHECK_CRITERIA, mosaic_util.h:26
constexpr double kEps8 = 1.e-8, kEps10 = 1.e-10, kEps15 = 1.e-15, kEps30 = 1.e-30;
constexpr double kPiD = 3.14159265358979323846;
constexpr double kR = 6371000.0;
constexpr int kRing = 12;           // vertices (<= 4) + inserted intersections (<= 8 between two convex quadrilaterals)
constexpr int kInter = 12;
constexpr int kPoly = 12;           // the overlap of two convex quadrilaterals has at most 8 vertices
constexpr int kMaxIn = 4;           // corners of an input cell
constexpr int kErrNotConvex = -1, kErrWalk = -2, kErrPool = -3;

// ---- double-double helpers (explicit fma: the file is compiled with -fmad=false) --------------------------
struct dd { double hi, lo; };
XGB_HD dd two_sum(double a, double b) { const double s = a + b, bb = s - a; return dd{s, (a - (s - bb)) + (b - bb)}; }
XGB_HD dd two_prod(double a, double b) { const double p = a * b; return dd{p, fma(a, b, -p)}; }
XGB_HD dd dd_norm(double hi, double lo) { const double s = hi + lo; return dd{s, lo - (s - hi)}; }
XGB_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_norm(s.hi, s.lo);
  s.lo += t.lo;
  return dd_norm(s.hi, s.lo);
}
XGB_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
XGB_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
XGB_HD dd dd_mul_d(dd a, double b) { dd p = two_prod(a.hi, b); p.lo = fma(a.lo, b, p.lo); return dd_norm(p.hi, p.lo); }
XGB_HD dd dd_mul(dd a, dd b) { dd p = two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return dd_norm(p.hi, p.lo); }
XGB_HD dd dd_div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = dd_sub(a, dd_mul_d(b, q1));
  const double q2 = r.hi / b.hi;
  r = dd_sub(r, dd_mul_d(b, q2));
  const double q3 = r.hi / b.hi;
  dd q = dd_norm(q1, q2);
  return dd_add(q, dd{q3, 0.0});
}
// a*b - c*d with doubles, exact products
XGB_HD dd det2(double a, double b, double c, double d) { return dd_sub(two_prod(a, b), two_prod(c, d)); }

// ---- acosl(x) rounded to double, as spherical_angle (mosaic_util.c:834) produces it ------------------------------------
// On x86-64 glibc's acosl is the x87 sequence fpatan(sqrt((1 - x)(1 + x)), x) (sysdeps/i386/fpu/e_acosl.c): the angle comes
// out with a 64-bit mantissa and the assignment to `double angle` rounds it a second time.  great_circle_area subtracts
// (n - 2) pi from the sum of the angles, so one unit in the last place of an angle is 1e-11 .. 1e-10 of a quarter-degree
// cell's area: the device acos() (1-2 ulp) cannot meet the 1e-12 the areas are held to.  Here the same formula runs in
// double-double (atan by a 65-row table of atan(k/64) and a Taylor tail, ~2^-100 relative), and the result is rounded twice
// the way the hardware does it: to 64 bits, then to 53, ties to even.  That reproduces the reference bit for bit wherever
// fpatan's own result is the correctly rounded one (Intel documents < 1 ulp of the 64-bit format; tests/test_gc_cpu.py counts
// the differences against this machine's acosl over millions of arguments).
#if defined(__CUDACC__)
static __device__ const double kAtanTabDev[65][2] = {
#include "atantab.inc"
};
#endif
static const double kAtanTabHost[65][2] = {
#include "atantab.inc"
};
XGB_HD dd atan_row(int k) {
#if defined(__CUDA_ARCH__)
  return dd{kAtanTabDev[k][0], kAtanTabDev[k][1]};
#else
  return dd{kAtanTabHost[k][0], kAtanTabHost[k][1]};
#endif
}
XGB_HD dd dd_sqrt(dd a) {
  if (!(a.hi > 0.0)) return dd{0.0, 0.0};
  const double s = sqrt(a.hi);
  const dd r = dd_sub(a, two_prod(s, s));                 // one Newton step from a 53-bit root
  return dd_norm(s, r.hi / (2.0 * s));
}
// atan(t), 0 <= t <= 1: atan(c) + atan((t - c) / (1 + t c)), c = k / 64 next to t, the second angle below 1 / 128
XGB_HD dd dd_atan01(dd t) {
  const int k = (int)(t.hi * 64.0 + 0.5);
  const double c = k * 0.015625;
  const dd r = dd_div(dd_add(t, dd{-c, 0.0}), dd_add(dd_mul_d(t, c), dd{1.0, 0.0}));
  const dd r2 = dd_mul(r, r);
  const double z = r2.hi;
  // atan(r) / r = 1 - z/3 + z^2 (1/5 - z/7 + z^2/9 - z^3/11 + z^4/13 - z^5/15), z <= 2^-14: the bracket in double is enough
  double p = fma(z, -1.0 / 15.0, 1.0 / 13.0);
  p = fma(z, p, -1.0 / 11.0);
  p = fma(z, p, 1.0 / 9.0);
  p = fma(z, p, -1.0 / 7.0);
  p = fma(z, p, 1.0 / 5.0);
  p = (z * z) * p;
  const dd third{0x1.5555555555555p-2, 0x1.5555555555555p-56};
  const dd w = dd_add(dd_neg(dd_mul(r2, third)), dd{p, 0.0});
  return dd_add(atan_row(k), dd_add(r, dd_mul(r, w)));
}
// v = hi + lo, hi = fl(v): v rounded to a 64-bit mantissa, then to double (both to nearest even).  The second rounding
// changes hi only when the first one lands exactly half way between hi and its neighbour on lo's side.
XGB_HD double round_through_x87(dd v) {
  const double hi = v.hi, lo = v.lo;
  if (lo == 0.0 || hi == 0.0) return hi;
#if defined(__CUDA_ARCH__)
  const double nb = __longlong_as_double(__double_as_longlong(hi) + (((lo > 0.0) == (hi > 0.0)) ? 1 : -1));
#else
  const double nb = nextafter(hi, (lo > 0.0) ? INFINITY : -INFINITY);
#endif
  const double gap = fabs(nb - hi);                       // spacing of the doubles on that side, exact
  const double half = 0.5 * gap, ulp64 = gap * (1.0 / 2048.0);
  if (fabs(lo) > half - 0.5 * ulp64) {                    // rounds to the half-way point at 64 bits: tie for the double
    const bool hi_even = (trig::bits(hi) & 1u) == 0u;
    return hi_even ? hi : nb;
  }
  return hi;
}
XGB_HD double gc_acos(double x) {
  const dd pi{0x1.921fb54442d18p+1, 0x1.1a62633145c07p-53}, half_pi{0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54};
  if (x >= 1.0) return 0.0;
  if (x <= -1.0) return kPiD;
  const double ax = fabs(x);
  const dd s = dd_sqrt(dd_mul(two_sum(1.0, -x), two_sum(1.0, x)));     // sin of the angle; 1 -+ x are exact as double-doubles
  dd a;
  if (s.hi <= ax) {                                                    // within 45 degrees of a pole of the formula
    a = dd_atan01(dd_div(s, dd{ax, 0.0}));
    if (x < 0.0) a = dd_sub(pi, a);
  } else {
    a = dd_atan01(dd_div(dd{ax, 0.0}, s));
    a = (x < 0.0) ? dd_add(half_pi, a) : dd_sub(half_pi, a);
  }
  return round_through_x87(a);
}

// asin and atan2 as glibc returns them.  The IBM Accurate Mathematical Library routines behind glibc's asin / atan2 aim at the
// correctly rounded result (and reach it on all but a few arguments per million); the same double-double atan, rounded once,
// therefore gives their bits.  Used by calc_c2l_grid_info's great_circle_distance and xyz2latlon (mosaic_util.c:228-252,
// :754-757) so that the order-2 gradient metrics computed on the device equal the reference's.
XGB_HD double gc_asin(double x) {
  const dd half_pi{0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54};
  const double ax = fabs(x);
  if (!(ax <= 1.0)) return asin(x);                                    // NaN / domain error: the toolchain's answer
  if (ax < 0x1p-27) return x;                                          // x + x^3/6 rounds to x
  if (ax == 1.0) return copysign(half_pi.hi, x);
  const dd s = dd_sqrt(dd_mul(two_sum(1.0, -ax), two_sum(1.0, ax)));   // cos of the angle
  dd a;
  if (ax <= s.hi) a = dd_atan01(dd_div(dd{ax, 0.0}, s));
  else a = dd_sub(half_pi, dd_atan01(dd_div(s, dd{ax, 0.0})));
  return copysign(a.hi, x);
}
XGB_HD double gc_atan2(double y, double x) {
  const dd pi{0x1.921fb54442d18p+1, 0x1.1a62633145c07p-53}, half_pi{0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54};
  const double ax = fabs(x), ay = fabs(y);
  if (!(ax < INFINITY) || !(ay < INFINITY) || (ax == 0.0 && ay == 0.0) || ay < ax * 0x1p-60 || ax < ay * 0x1p-60)
    return atan2(y, x);                                                // zeros, infinities, NaN, axis-hugging: the toolchain's answer
  dd a;
  if (ay <= ax) a = dd_atan01(dd_div(dd{ay, 0.0}, dd{ax, 0.0}));
  else a = dd_sub(half_pi, dd_atan01(dd_div(dd{ax, 0.0}, dd{ay, 0.0})));
  if (x < 0.0) a = dd_sub(pi, a);
  return copysign(a.hi, y);
}

struct V3 { double x, y, z; };

XGB_HD bool same_point(double x1, double y1, double z1, double x2, double y2, double z2) {   // mosaic_util.c:1193
  return !(fabs(x1 - x2) > kEps10 || fabs(y1 - y2) > kEps10 || fabs(z1 - z2) > kEps10);
}

XGB_HD V3 cross(const V3& a, const V3& b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// spherical_angle (mosaic_util.c:800-838), plain double as an autotools build compiles it.  exact: the angle as the reference
// rounds it (gc_acos); otherwise the toolchain's acos, within 2 ulp of it, for sums that are only compared with a tolerance.
XGB_HD_CALL double spherical_angle_x(const V3& v1, const V3& v2, const V3& v3, bool exact) {
  const V3 p = cross(v1, v2), q = cross(v1, v3);
  double ddd = (p.x * p.x + p.y * p.y + p.z * p.z) * (q.x * q.x + q.y * q.y + q.z * q.z);
  if (ddd <= 0.0) return 0.;
  ddd = (p.x * q.x + p.y * q.y + p.z * q.z) / sqrt(ddd);
  if (fabs(ddd - 1) < kEps30) ddd = 1;
  if (fabs(ddd + 1) < kEps30) ddd = -1;
  if (ddd > 1. || ddd < -1.) return (ddd < 0.) ? kPiD : 0.;
  return exact ? gc_acos(ddd) : acos(ddd);
}
XGB_HD double spherical_angle(const V3& v1, const V3& v2, const V3& v3) { return spherical_angle_x(v1, v2, v3, true); }

// great_circle_area (mosaic_util.c:763-790) of n vertices v[0..n)
template <class Get>
XGB_HD double great_circle_area(int n, Get get) {
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const int i1 = (i + 1 < n) ? i + 1 : i + 1 - n;
    const int i2 = (i + 2 < n) ? i + 2 : i + 2 - n;
    sum += spherical_angle(get(i1), get(i2), get(i));
  }
  return (sum - (n - 2.) * kPiD) * kR * kR;
}

// vertex ring node / intersection node
struct RNode { double x, y, z, u; signed char intersect, inbound, inside; };   // 40 bytes: the rings live in local memory
struct INode { double x, y, z, u, u_clip; signed char inbound, subj_index, clip_index; };
struct Ring { RNode n[kRing]; int len; bool overflow; };

XGB_HD void ring_append_unique(Ring& l, double x, double y, double z) {          // addEnd(list, x, y, z, 0, 0, 0, -1)
  for (int k = 0; k < l.len; ++k) if (same_point(l.n[k].x, l.n[k].y, l.n[k].z, x, y, z)) return;
  if (l.len >= kRing) { l.overflow = true; return; }
  l.n[l.len++] = RNode{x, y, z, 0.0, 0, 0, -1};
}

// insidePolygon (mosaic_util.c:1487-1530): |sum of the angles the sides subtend - 2 pi| < 1e-8.  The sum is taken with the
// toolchain's acos first (a seventh of the instructions of gc_acos); it is within 2e-14 of the reference's, so only a sum
// within 1e-12 of the threshold is taken again with the reference's own roundings.  Same decision in every case.
XGB_HD_CALL bool inside_polygon(const RNode& q, const Ring& l) {
  const V3 p0{q.x, q.y, q.z};
  for (int pass = 0; pass < 2; ++pass) {
    double sum = 0;
    for (int k = 0; k < l.len; ++k) {
      const RNode& a = l.n[k];
      const RNode& b = l.n[(k + 1 < l.len) ? k + 1 : 0];
      if (same_point(p0.x, p0.y, p0.z, a.x, a.y, a.z)) return true;
      sum += spherical_angle_x(p0, V3{b.x, b.y, b.z}, V3{a.x, a.y, a.z}, pass == 1);
    }
    const double dev = fabs(sum - 2 * kPiD);
    if (pass == 1 || fabs(dev - kEps8) > 1.e-12) return dev < kEps8;
  }
  return false;
}

XGB_HD double ring_area(const Ring& l) {                                          // gridArea, mosaic_util.c:1364
  return great_circle_area(l.len, [&](int k) { return V3{l.n[k].x, l.n[k].y, l.n[k].z}; });
}

// gridArea(l) > 0 (mosaic_util.c:1364; the convexity test of create_xgrid.c:1575): sign of the spherical excess, settled by
// the fast angles unless the excess is below 1e-12 rad
XGB_HD bool ring_area_positive(const Ring& l) {
  for (int pass = 0; pass < 2; ++pass) {
    double sum = 0.0;
    const int n = l.len;
    for (int i = 0; i < n; ++i) {
      const int i1 = (i + 1 < n) ? i + 1 : i + 1 - n;
      const int i2 = (i + 2 < n) ? i + 2 : i + 2 - n;
      sum += spherical_angle_x(V3{l.n[i1].x, l.n[i1].y, l.n[i1].z}, V3{l.n[i2].x, l.n[i2].y, l.n[i2].z}, V3{l.n[i].x, l.n[i].y, l.n[i].z}, pass == 1);
    }
    const double excess = sum - (n - 2.) * kPiD;
    if (pass == 1 || fabs(excess) > 1.e-12) return excess * kR * kR > 0;
  }
  return false;
}

// insertIntersect (mosaic_util.c:1291-1362); false if vertex v is not in the ring (the reference aborts)
XGB_HD_CALL bool ring_insert(Ring& l, const V3& p, double u1, double u2, int inbound, const V3& v) {
  int a = -1;
  for (int k = 0; k < l.len; ++k) if (l.n[k].x == v.x && l.n[k].y == v.y && l.n[k].z == v.z) { a = k; break; }
  if (a < 0) return false;
  double ucur = u1;
  if (u1 == 1) { ucur = 0; a = (a + 1 < l.len) ? a + 1 : 0; }
  if (ucur == 0) {
    RNode& t = l.n[a];
    t.intersect = 2; t.inside = 1; t.u = ucur; t.x = p.x; t.y = p.y; t.z = p.z;
    return true;
  }
  if (u2 != 0 && u2 != 1) {
    if (inbound == 1) {
      int b = (a + 1 < l.len) ? a + 1 : 0;
      int guard = 0;
      while (l.n[b].intersect && guard++ < kRing) b = (b + 1 < l.len) ? b + 1 : 0;
      l.n[b].inside = 0;
    } else if (inbound == 2) l.n[a].inside = 0;
  }
  int b = a + 1;
  while (b < l.len) {
    if (l.n[b].intersect == 1) { if (l.n[b].u > ucur) break; }
    else break;
    a = b; ++b;
  }
  if (l.len >= kRing) { l.overflow = true; return true; }
  for (int k = l.len; k > a + 1; --k) l.n[k] = l.n[k - 1];
  ++l.len;
  l.n[a + 1] = RNode{p.x, p.y, p.z, ucur, 1, (signed char)inbound, 1};
  return true;
}

// parameter t where the line l1 + t (l2 - l1) meets the plane through a, b and the origin
// (intersect_tri_with_line, mosaic_util.c:967-1008: M = [l1-l2 | b-a | 0-a], x = M^-1 (l1-a), t = x[0]); double-double
XGB_HD_CALL bool plane_line_param(const V3& a, const V3& b, const V3& l1, const V3& l2, double* t) {
  const double m0 = l1.x - l2.x, m1 = b.x - a.x, m2 = 0.0 - a.x;
  const double m3 = l1.y - l2.y, m4 = b.y - a.y, m5 = 0.0 - a.y;
  const double m6 = l1.z - l2.z, m7 = b.z - a.z, m8 = 0.0 - a.z;
  {
    // Plain-double estimate first.  The caller only asks whether t lies in [0, 1] (after snapping within 1e-8) unless it
    // does; an estimate that is off by far less than 1e-4 settles the many side pairs that do not meet, and the
    // double-double solve below is kept for those that (nearly) do and for ill-conditioned systems.
    const double e0 = m4 * m8 - m5 * m7, e1 = m3 * m8 - m5 * m6, e2 = m3 * m7 - m4 * m6;
    const double det_d = m0 * e0 - m1 * e1 + m2 * e2;
    const double scale = fabs(m0 * e0) + fabs(m1 * e1) + fabs(m2 * e2);
    if (fabs(det_d) > 1e-9 * scale && fabs(det_d) > 1e-12) {
      const double w0 = l1.x - a.x, w1 = l1.y - a.y, w2 = l1.z - a.z;
      const double t_d = (e0 * w0 + (m2 * m7 - m1 * m8) * w1 + (m1 * m5 - m2 * m4) * w2) / det_d;
      if (t_d < -1e-4 || t_d > 1.0 + 1e-4) { *t = t_d; return true; }
    }
  }
  const dd c0 = det2(m4, m8, m5, m7);                 // cofactors of the first column
  const dd c1 = det2(m3, m8, m5, m6);
  const dd c2 = det2(m3, m7, m4, m6);
  const dd det = dd_add(dd_sub(dd_mul_d(c0, m0), dd_mul_d(c1, m1)), dd_mul_d(c2, m2));
  if (fabs(det.hi) < kEps15) return false;
  const dd r1 = det2(m2, m7, m1, m8), r2 = det2(m1, m5, m2, m4);   // first row of the adjugate: c0, r1, r2
  const double v0 = l1.x - a.x, v1 = l1.y - a.y, v2 = l1.z - a.z;
  const dd num = dd_add(dd_add(dd_mul_d(c0, v0), dd_mul_d(r1, v1)), dd_mul_d(r2, v2));
  const dd q = dd_div(num, det);
  *t = q.hi + q.lo;
  return true;
}

// line_intersect_2D_3D (create_xgrid.c:1919-2081)
XGB_HD bool line_intersect(const V3& a1, const V3& a2, const V3& q1, const V3& q2, const V3& q3,
                           V3* p, double* ua, double* uq, int* inbound) {
  *inbound = 0;
  if (same_point(a1.x, a1.y, a1.z, q1.x, q1.y, q1.z)) { *ua = 0; *uq = 0; *p = a1; return true; }
  if (same_point(a1.x, a1.y, a1.z, q2.x, q2.y, q2.z)) { *ua = 0; *uq = 1; *p = a1; return true; }
  if (same_point(a2.x, a2.y, a2.z, q1.x, q1.y, q1.z)) { *ua = 1; *uq = 0; *p = a2; return true; }
  if (same_point(a2.x, a2.y, a2.z, q2.x, q2.y, q2.z)) { *ua = 1; *uq = 1; *p = a2; return true; }
  if (!plane_line_param(q1, q2, a1, a2, ua)) return false;
  if (fabs(*ua) < kEps8) *ua = 0;
  if (fabs(*ua - 1) < kEps8) *ua = 1;
  if (*ua < 0 || *ua > 1) return false;
  if (!plane_line_param(a1, a2, q1, q2, uq)) return false;
  if (fabs(*uq) < kEps8) *uq = 0;
  if (fabs(*uq - 1) < kEps8) *uq = 1;
  if (*uq < 0 || *uq > 1) return false;
  const double u = *ua;
  const V3 c3 = cross(cross(a1, a2), cross(q1, q2));
  if (fabs(sqrt(c3.x * c3.x + c3.y * c3.y + c3.z * c3.z)) < kEps30) return false;
  V3 r{a1.x + u * (a2.x - a1.x), a1.y + u * (a2.y - a1.y), a1.z + u * (a2.z - a1.z)};
  const double norm = sqrt(r.x * r.x + r.y * r.y + r.z * r.z);
  r.x /= norm; r.y /= norm; r.z /= norm;
  *p = r;
  if (*uq != 0 && *uq != 1) {
    const V3 d{a2.x - a1.x, a2.y - a1.y, a2.z - a1.z}, v1{q2.x - q1.x, q2.y - q1.y, q2.z - q1.z}, v2{q3.x - q2.x, q3.y - q2.y, q3.z - q2.z};
    const V3 c1 = cross(v1, v2), c2 = cross(v1, d);
    *inbound = (c1.x * c2.x + c1.y * c2.y + c1.z * c2.z > 0) ? 2 : 1;
  }
  return true;
}

XGB_HD int ring_find(const Ring& l, double x, double y, double z) {
  for (int k = 0; k < l.len; ++k) if (l.n[k].x == x && l.n[k].y == y && l.n[k].z == z) return k;
  return -1;
}

struct Poly { V3* v; int len; bool overflow; };      // a view of the caller's output array (capacity kPoly)
XGB_HD void poly_append_unique(Poly& l, double x, double y, double z) {
  for (int k = 0; k < l.len; ++k) if (same_point(l.v[k].x, l.v[k].y, l.v[k].z, x, y, z)) return;
  if (l.len >= kPoly) { l.overflow = true; return; }
  l.v[l.len++] = V3{x, y, z};
}

// clip_2dx2d_great_circle: vertices of cell 1 (subject) and cell 2 (clip) in the reference's clockwise order.
// Returns the vertex count of the overlap polygon written to out (capacity kPoly), 0 if none, < 0 where the
// reference would abort.
//
// The routine is written as phases — ring construction, inside flags (angle sums), the (side of 1) x (side of 2) intersection
// steps, the walk — so that a thread block can pass through them TOGETHER: with SYNC > 0 every thread of the block calls it
// (live = false for padding threads) and meets the others at a barrier before each phase (SYNC >= 2: before each side of cell
// 1, SYNC >= 3: before each pair of sides).  The kernel is bound by instruction fetch; warps that execute the same phase at
// the same time share its cache lines instead of evicting each other's (xgrid_gc_kernels.cu, gc_clip_kernel).  SYNC = 0 is
// the plain routine (host build, small launches); the arithmetic and its order are the same in every mode.
struct ClipState {
  Ring g1, g2;
  V3 pt1[kMaxIn], pt2[kMaxIn];
  INode inter[kInter];
  int ninter, npts1, npts2;
  bool inter_overflow;
};

// one (side i1 of cell 1, side i2 of cell 2) step of :1606-1670; false where the reference would abort
XGB_HD bool clip_step(ClipState& S, int i1, int i2) {
  Ring& g1 = S.g1; Ring& g2 = S.g2;
  V3* pt1 = S.pt1; V3* pt2 = S.pt2;
  INode* inter = S.inter;
  int& ninter = S.ninter;
  bool& inter_overflow = S.inter_overflow;
  const int npts1 = S.npts1, npts2 = S.npts2;
  const int i1p = (i1 + 1 < npts1) ? i1 + 1 : 0;
  const int i2p = (i2 + 1 < npts2) ? i2 + 1 : i2 + 1 - npts2;
  const int i2p2 = (i2 + 2 < npts2) ? i2 + 2 : i2 + 2 - npts2;
  V3 p;
  double u1, u2;
  int inbound;
  if (!line_intersect(pt1[i1], pt1[i1p], pt2[i2], pt2[i2p], pt2[i2p2], &p, &u1, &u2, &inbound)) return true;
  {                                                                                      // addIntersect :1133-1186
    double u1c = u1, u2c = u2;
    int s = i1, c = i2;
    if (u1c == 1) { u1c = 0; s = i1p; }
    if (u2c == 1) { u2c = 0; c = i2p; }
    bool dup = false;
    for (int k = 0; k < ninter; ++k)
      if ((inter[k].u == u1c && inter[k].subj_index == s) || (inter[k].u_clip == u2c && inter[k].clip_index == c)) { dup = true; break; }
    if (dup) return true;
    if (ninter >= kInter) { inter_overflow = true; return true; }
    inter[ninter++] = INode{p.x, p.y, p.z, u1c, u2c, (signed char)inbound, (signed char)s, (signed char)c};
  }
  if (u1 == 1) { if (!ring_insert(g1, p, 0.0, u2, inbound, pt1[i1p])) return false; }
  else         { if (!ring_insert(g1, p, u1, u2, inbound, pt1[i1])) return false; }
  if (u1 == 1) pt1[i1p] = p; else if (u1 == 0) pt1[i1] = p;
  if (u2 == 1) { if (!ring_insert(g2, p, 0.0, u1, 0, pt2[i2p])) return false; }
  else         { if (!ring_insert(g2, p, u2, u1, 0, pt2[i2])) return false; }
  if (u2 == 1) pt2[i2p] = p; else if (u2 == 0) pt2[i2] = p;

  return true;
}

// the walk over the two rings from the first inbound intersection (:1676-1904)
XGB_HD int clip_walk(ClipState& S, V3* out) {
  Ring& g1 = S.g1; Ring& g2 = S.g2;
  INode* inter = S.inter;
  const int ninter = S.ninter, npts1 = S.npts1, npts2 = S.npts2;
  int nint = ninter, first = -1;                                                             // :1676-1693
  if (nint > 1) for (int k = 0; k < ninter; ++k) if (inter[k].inbound == 2) { first = k; break; }
  if (first < 0 && nint > 1) {                                                               // setInbound :1437-1470
    for (int k = 0; k < ninter; ++k) {
      if (inter[k].inbound) continue;
      const int a = ring_find(g1, inter[k].x, inter[k].y, inter[k].z);
      if (a < 0) return kErrWalk;
      const RNode& prev = g1.n[a > 0 ? a - 1 : g1.len - 1];
      const RNode& next = g1.n[a + 1 < g1.len ? a + 1 : 0];
      inter[k].inbound = (prev.inside == 0 && next.inside == 1) ? 2 : 1;
    }
    for (int k = 0; k < ninter; ++k) if (inter[k].inbound == 2) { first = k; break; }
  }

  int n_out = 0;
  if (first >= 0) {                                                                          // :1697-1838
    Poly poly;
    poly.v = out; poly.len = 0; poly.overflow = false;
    const V3 f{inter[first].x, inter[first].y, inter[first].z};
    if (ring_find(g1, f.x, f.y, f.z) < 0) return kErrWalk;
    poly_append_unique(poly, f.x, f.y, f.z);
    --nint;
    V3 cur = f;
    const int maxiter1 = ninter;
    int iter1 = 0;
    bool found1 = false, on2 = false;
    while (iter1 < maxiter1) {
      const Ring& cl = on2 ? g2 : g1;
      const int a = ring_find(cl, cur.x, cur.y, cur.z);
      if (a < 0) return kErrWalk;
      int b = (a + 1 < cl.len) ? a + 1 : 0;
      bool found2 = false;
      for (int iter2 = 0; iter2 < cl.len; ++iter2) {
        bool b_is_x = false;
        const RNode& nb = cl.n[b];
        if (nb.intersect) {
          if (nb.x == f.x && nb.y == f.y && nb.z == f.z) { found1 = true; break; }
          const RNode& nc = cl.n[(b + 1 < cl.len) ? b + 1 : 0];
          found2 = true; b_is_x = true;
          if (nc.intersect || nc.inside == 1) found2 = false;
        }
        if (found2) { cur = V3{nb.x, nb.y, nb.z}; break; }
        poly_append_unique(poly, nb.x, nb.y, nb.z);
        if (b_is_x) --nint;
        b = (b + 1 < cl.len) ? b + 1 : 0;
      }
      if (found1) break;
      if (!found2) return kErrWalk;
      if (cur.x == f.x && cur.y == f.y && cur.z == f.z) { found1 = true; break; }
      poly_append_unique(poly, cur.x, cur.y, cur.z);
      --nint;
      on2 = !on2;
      ++iter1;
    }
    if (!found1 || nint > 0) return kErrWalk;
    if (poly.overflow) return kErrPool;
    n_out = poly.len;
    if (n_out < 3) n_out = 0;
  }
  if (n_out == 0) {                                                                          // :1841-1871
    int c = 0;
    for (int k = 0; k < g1.len; ++k) if (g1.n[k].intersect != 1 && g1.n[k].inside == 1) ++c;
    if (c == npts1) {
      for (int k = 0; k < npts1; ++k) out[k] = V3{g1.n[k].x, g1.n[k].y, g1.n[k].z};
      return npts1;
    }
  }
  if (n_out == 0) {                                                                          // :1874-1904
    int c = 0;
    for (int k = 0; k < g2.len; ++k) if (g2.n[k].intersect != 1 && g2.n[k].inside == 1) ++c;
    if (c == npts2) {
      for (int k = 0; k < npts2; ++k) out[k] = V3{g2.n[k].x, g2.n[k].y, g2.n[k].z};
      n_out = npts2;
    }
  }
  return n_out;
}

#if defined(__CUDA_ARCH__)
#define XGB_GC_BAR(level) do { if (SYNC >= (level)) __syncthreads(); } while (0)
#else
#define XGB_GC_BAR(level) do { } while (0)
#endif

template <int SYNC>
XGB_HD int clip_great_circle_t(const V3* c1, int n1, const V3* c2, int n2, V3* out, bool live) {
  int status = live ? 1 : 0;                                     // 1: still working; otherwise the value to return
  if (status == 1 && (n1 > kMaxIn || n2 > kMaxIn)) status = kErrPool;
  if (status == 1) {                                             // six range rejections (:1508-1528)
    double lo1[3], hi1[3], lo2[3], hi2[3];
    lo1[0] = hi1[0] = c1[0].x; lo1[1] = hi1[1] = c1[0].y; lo1[2] = hi1[2] = c1[0].z;
    for (int k = 1; k < n1; ++k) {
      lo1[0] = fmin(lo1[0], c1[k].x); hi1[0] = fmax(hi1[0], c1[k].x);
      lo1[1] = fmin(lo1[1], c1[k].y); hi1[1] = fmax(hi1[1], c1[k].y);
      lo1[2] = fmin(lo1[2], c1[k].z); hi1[2] = fmax(hi1[2], c1[k].z);
    }
    lo2[0] = hi2[0] = c2[0].x; lo2[1] = hi2[1] = c2[0].y; lo2[2] = hi2[2] = c2[0].z;
    for (int k = 1; k < n2; ++k) {
      lo2[0] = fmin(lo2[0], c2[k].x); hi2[0] = fmax(hi2[0], c2[k].x);
      lo2[1] = fmin(lo2[1], c2[k].y); hi2[1] = fmax(hi2[1], c2[k].y);
      lo2[2] = fmin(lo2[2], c2[k].z); hi2[2] = fmax(hi2[2], c2[k].z);
    }
    for (int a = 0; a < 3; ++a)
      if (lo1[a] >= hi2[a] + kRange || lo2[a] >= hi1[a] + kRange) status = 0;
  }
  ClipState S;
  Ring& g1 = S.g1; Ring& g2 = S.g2;
  g1.len = g2.len = 0; g1.overflow = g2.overflow = false;
  S.ninter = 0; S.inter_overflow = false; S.npts1 = S.npts2 = 0;
  if (status == 1) {
    for (int k = 0; k < n1; ++k) ring_append_unique(g1, c1[k].x, c1[k].y, c1[k].z);
    for (int k = 0; k < n2; ++k) ring_append_unique(g2, c2[k].x, c2[k].y, c2[k].z);
    S.npts1 = g1.len; S.npts2 = g2.len;
  }
  XGB_GC_BAR(1);
  if (status == 1) {
    for (int k = 0; k < g1.len; ++k) g1.n[k].inside = inside_polygon(g1.n[k], g2) ? 1 : 0;   // :1549-1568
    for (int k = 0; k < g2.len; ++k) g2.n[k].inside = inside_polygon(g2.n[k], g1) ? 1 : 0;
    if (!ring_area_positive(g1) || !ring_area_positive(g2)) status = kErrNotConvex;          // :1575-1578
  }
  if (status == 1) {
    for (int k = 0; k < S.npts1; ++k) S.pt1[k] = V3{g1.n[k].x, g1.n[k].y, g1.n[k].z};
    for (int k = 0; k < S.npts2; ++k) S.pt2[k] = V3{g2.n[k].x, g2.n[k].y, g2.n[k].z};
  }
  XGB_GC_BAR(1);
  const int lim1 = SYNC ? kMaxIn : S.npts1, lim2 = SYNC ? kMaxIn : S.npts2;
  for (int i1 = 0; i1 < lim1; ++i1) {                                                        // :1606-1670
    XGB_GC_BAR(2);
    for (int i2 = 0; i2 < lim2; ++i2) {
      XGB_GC_BAR(3);
      if (status == 1 && i1 < S.npts1 && i2 < S.npts2 && !clip_step(S, i1, i2)) status = kErrWalk;
    }
  }
  if (status == 1 && (g1.overflow || g2.overflow || S.inter_overflow)) status = kErrPool;
  XGB_GC_BAR(1);
  return (status == 1) ? clip_walk(S, out) : status;
}

XGB_HD int clip_great_circle(const V3* c1, int n1, const V3* c2, int n2, V3* out) {
  return clip_great_circle_t<0>(c1, n1, c2, n2, out, true);
}

// Cheap proof that two cells cannot produce an exchange cell: some side of one cell has every corner of the other on its
// outer side by more than kSepMargin (sine of the angular distance to the side's great circle; the inner side is where
// the cell's own opposite corner lies, so the test does not depend on the orientation of the grid).  The spherical cap
// n.x >= margin is convex, so it then holds the whole other cell; the cells are at least ~1e-5 rad apart, far beyond the
// reference's tolerances (EPSLN8 on edge parameters, EPSLN10 on point identity, 1e-8 on insidePolygon's angle sum, whose
// deficit for a point 1e-5 outside a side of length L is 8e-5/L), so the reference finds no intersection and no inside
// vertex and returns 0 for such a pair.  Degenerate sides (pole triangles) have n = 0 and never separate.
constexpr double kSepMargin = 1.e-5;
XGB_HD bool separated_by_side(const V3* a, const V3* b) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const V3 n = cross(a[i], a[(i + 1) & 3]);
    const V3& o2 = a[(i + 2) & 3];
    const V3& o3 = a[(i + 3) & 3];
    const double in2 = n.x * o2.x + n.y * o2.y + n.z * o2.z, in3 = n.x * o3.x + n.y * o3.y + n.z * o3.z;
    const double inner = (fabs(in2) > fabs(in3)) ? in2 : in3;      // the cell's own far corner marks the inner side
    const double len2 = n.x * n.x + n.y * n.y + n.z * n.z;
    if (!(len2 > 1e-24)) continue;                                 // coincident corners (pole triangle)
    const double lim = kSepMargin * sqrt(len2);
    // a far corner within the margin of the side's circle (sliver cell), or the two far corners on opposite sides
    // (twisted cell), give no reliable orientation: do not use this side
    if (!(fabs(inner) > 10.0 * lim) || in2 * inner < -lim * fabs(inner) || in3 * inner < -lim * fabs(inner)) continue;
    bool out = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double dk = n.x * b[k].x + n.y * b[k].y + n.z * b[k].z;
      out = out && ((inner > 0.0) ? (dk < -lim) : (dk > lim));
    }
    if (out) return true;
  }
  return false;
}

// latlon2xyz (mosaic_util.c:212-222) with the sin/cos entry points the reference binary calls
XGB_HD V3 ll2xyz(double lon, double lat) {
  const double cl = ref_cos(lat);
  return V3{cl * ref_cos(lon), cl * ref_sin(lon), ref_sin(lat)};
}

}  // namespace gc
}  // namespace xgb
