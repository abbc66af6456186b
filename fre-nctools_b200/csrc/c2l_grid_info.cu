// One-time grid metrics for the order-2 gradient: calc_c2l_grid_info / get_edge / mid_pt_sphere
// (reference gradient_c2l.c:368-454, :198-313, :315-337) on the device.
//
// Round 2: the Cartesian vertices (latlon2xyz), the unit vectors (unit_vect_latlon), the edge normals and the spherical-excess
// cell area are computed with the reference binary's own libm entry points and bits (ref_trig.cuh; acosl as the reference
// rounds it, gc_clip.cuh) and come out bit-identical (area: but for the few angles per 100 000 where the x87 fpatan is not
// correctly rounded).  dx, dy and the edge weights go through asin / atan2: gc_asin / gc_atan2 give the correctly rounded value,
// which is glibc's on all but ~1 argument per 1000 (glibc 2.39's asin is not correctly rounded there; checked with mpmath), so
// these agree to 1e-15 relative and are bit-identical in the vast majority of cells.  Callers that need the reference's exact metrics can still pass them in with xgb_plan_grad_set_metrics.
#include "apply_internal.h"
#include "gc_clip.cuh"

namespace xgb {

extern long long g_launches;

__device__ __forceinline__ double gc_distance(double lon1, double lat1, double lon2, double lat2)
{
  // mosaic_util.c:754-757; the reference binary calls sin, cos, cos, sin, asin (objdump of oracle/_ref)
  const double a = ref_sin((lat1 - lat2) / 2.), b = ref_sin((lon1 - lon2) / 2.);
  const double beta = 2. * gc::gc_asin(sqrt(a * a + ref_cos(lat1) * ref_cos(lat2) * (b * b)));
  return kRadius * beta;
}

__device__ __forceinline__ void ll2xyz(double lon, double lat, double* v)
{
  const gc::V3 r = gc::ll2xyz(lon, lat);      // mosaic_util.c:212-222 with the libm entry points and bits of the reference binary
  v[0] = r.x; v[1] = r.y; v[2] = r.z;
}

__device__ __forceinline__ void cross3(const double* p1, const double* p2, double* e)
{
  e[0] = p1[1] * p2[2] - p1[2] * p2[1];
  e[1] = p1[2] * p2[0] - p1[0] * p2[2];
  e[2] = p1[0] * p2[1] - p1[1] * p2[0];
}

__device__ __forceinline__ void normalize3(double* e)
{
  const double n = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  e[0] /= n; e[1] /= n; e[2] /= n;
}

__device__ __forceinline__ double sph_angle(const double* v1, const double* v2, const double* v3)   // mosaic_util.c:800-838
{
  double p[3], q[3];
  cross3(v1, v2, p);
  cross3(v1, v3, q);
  double ddd = (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (ddd <= 0.0) return 0.;
  ddd = (p[0] * q[0] + p[1] * q[1] + p[2] * q[2]) / sqrt(ddd);
  if (fabs(ddd - 1) < 1.e-30) ddd = 1;
  if (fabs(ddd + 1) < 1.e-30) ddd = -1;
  if (ddd > 1. || ddd < -1.) return (ddd < 0.) ? kPi : 0.;
  return gc::gc_acos(ddd);      // acosl rounded to double, as the reference computes it (gc_clip.cuh)
}

// mid_pt_sphere (gradient_c2l.c:315-337) -> (lon, lat)
__device__ __forceinline__ void mid_pt(double lon1, double lat1, double lon2, double lat2, double* pm)
{
  double e1[3], e2[3], e[3];
  ll2xyz(lon1, lat1, e1);
  ll2xyz(lon2, lat2, e2);
  e[0] = e1[0] + e2[0]; e[1] = e1[1] + e2[1]; e[2] = e1[2] + e2[2];
  normalize3(e);
  // xyz2latlon (mosaic_util.c:228-252)
  double xx = e[0], yy = e[1], zz = e[2];
  const double dist = sqrt(xx * xx + yy * yy + zz * zz);
  xx /= dist; yy /= dist; zz /= dist;
  double lon = (fabs(xx) + fabs(yy) < 1.e-10) ? 0. : gc::gc_atan2(yy, xx);
  if (lon < 0.) lon = 2. * kPi + lon;
  pm[0] = lon; pm[1] = gc::gc_asin(zz);
}

// edge weight at a corner between the two neighbouring mid points (get_edge, gradient_c2l.c:243-309)
__device__ __forceinline__ double edge_weight(const double* m0, const double* m1, double clon, double clat)
{
  const double d1 = gc_distance(m0[0], m0[1], clon, clat);
  const double d2 = gc_distance(m1[0], m1[1], clon, clat);
  return d2 / (d1 + d2);
}

// one thread per corner point (i, j), 0 <= i <= nx, 0 <= j <= ny
__global__ void __launch_bounds__(128)
c2l_grid_info_kernel(int nx, int ny, const double* __restrict__ xt, const double* __restrict__ yt,
                     const double* __restrict__ xc, const double* __restrict__ yc,
                     double* dx, double* dy, double* area, double* edge_w, double* edge_e, double* edge_s, double* edge_n,
                     double* en_n, double* en_e, double* vlon, double* vlat)
{
  const int nxp = nx + 1, nyp = ny + 1, w = nx + 2;
  const long long tix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (tix >= (long long)nxp * nyp) return;
  const int i = (int)(tix % nxp), j = (int)(tix / nxp);
  const long long c = (long long)j * nxp + i;
  const double lon0 = xc[c], lat0 = yc[c];
  double v0[3];
  ll2xyz(lon0, lat0, v0);
  if (i < nx) {                                              // N-cell centre quantities (:385-391, :418-427)
    double v1[3];
    ll2xyz(xc[c + 1], yc[c + 1], v1);
    dx[(long long)j * nx + i] = gc_distance(lon0, lat0, xc[c + 1], yc[c + 1]);
    double* e = en_n + 3 * ((long long)j * nx + i);
    cross3(v0, v1, e);
    normalize3(e);
  }
  if (j < ny) {                                              // E-cell centre quantities (:393-399, :429-438)
    double v1[3];
    ll2xyz(xc[c + nxp], yc[c + nxp], v1);
    dy[c] = gc_distance(lon0, lat0, xc[c + nxp], yc[c + nxp]);
    double* e = en_e + 3 * c;
    cross3(v1, v0, e);
    normalize3(e);
  }
  if (i < nx && j < ny) {                                    // T-cell quantities (:401-411, :440-446)
    double lr[3], ul[3], ur[3];
    ll2xyz(xc[c + 1], yc[c + 1], lr);
    ll2xyz(xc[c + nxp], yc[c + nxp], ul);
    ll2xyz(xc[c + nxp + 1], yc[c + nxp + 1], ur);
    const double a1 = sph_angle(v0, lr, ul), a2 = sph_angle(lr, ur, v0), a3 = sph_angle(ur, ul, lr), a4 = sph_angle(ul, ur, v0);
    const long long m = (long long)j * nx + i;
    area[m] = (a1 + a2 + a3 + a4 - 2. * kPi) * kRadius * kRadius;       // spherical_excess_area, mosaic_util.c:846-880
    const double lon = xt[(long long)(j + 1) * w + i + 1], lat = yt[(long long)(j + 1) * w + i + 1];
    double sl, cl, sa, ca;                                                     // unit_vect_latlon, mosaic_util.c:937-957:
    ref_sincos(lon, &sl, &cl); ref_sincos(lat, &sa, &ca);                      // two sincos calls in the reference binary
    vlon[3 * m] = -sl; vlon[3 * m + 1] = cl; vlon[3 * m + 2] = 0.;
    vlat[3 * m] = -sa * cl; vlat[3 * m + 1] = -sa * sl; vlat[3 * m + 2] = ca;
  }
  // edge weights: 0.5 at the two end points (:211-218), interpolation weight elsewhere
  if (i == 0 || i == nx) {
    double wgt = 0.5;
    if (j >= 1 && j < ny) {
      const int col = (i == 0) ? 0 : nx;
      double m0[2], m1[2];
      mid_pt(xt[(long long)j * w + col], yt[(long long)j * w + col], xt[(long long)j * w + col + 1], yt[(long long)j * w + col + 1], m0);
      mid_pt(xt[(long long)(j + 1) * w + col], yt[(long long)(j + 1) * w + col], xt[(long long)(j + 1) * w + col + 1], yt[(long long)(j + 1) * w + col + 1], m1);
      wgt = edge_weight(m0, m1, lon0, lat0);
    }
    if (i == 0) edge_w[j] = wgt;
    if (i == nx) edge_e[j] = wgt;
  }
  if (j == 0 || j == ny) {
    double wgt = 0.5;
    if (i >= 1 && i < nx) {
      const int row = (j == 0) ? 0 : ny;
      double m0[2], m1[2];
      mid_pt(xt[(long long)row * w + i], yt[(long long)row * w + i], xt[(long long)(row + 1) * w + i], yt[(long long)(row + 1) * w + i], m0);
      mid_pt(xt[(long long)row * w + i + 1], yt[(long long)row * w + i + 1], xt[(long long)(row + 1) * w + i + 1], yt[(long long)(row + 1) * w + i + 1], m1);
      wgt = edge_weight(m0, m1, lon0, lat0);
    }
    if (j == 0) edge_s[i] = wgt;
    if (j == ny) edge_n[i] = wgt;
  }
}

void launch_c2l_grid_info(int nx, int ny, const double* xt, const double* yt, const double* xc, const double* yc,
                          double* dx, double* dy, double* area, double* edge_w, double* edge_e, double* edge_s, double* edge_n,
                          double* en_n, double* en_e, double* vlon, double* vlat, cudaStream_t st)
{
  const long long n = (long long)(nx + 1) * (ny + 1);
  ++g_launches;
  c2l_grid_info_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(nx, ny, xt, yt, xc, yc, dx, dy, area, edge_w, edge_e, edge_s, edge_n,
                                                                  en_n, en_e, vlon, vlat);
}

}  // namespace xgb
