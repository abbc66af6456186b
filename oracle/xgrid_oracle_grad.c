/* TEST INFRASTRUCTURE — CPU oracle, never shipped or measured as the product.
 *
 * Order-2 gradient terms of fregrid's conservative path, restated from the reference:
 *   orc_grad_c2l            gradient_c2l.c:58-118 (grad_c2l) with a2b_ord2 :124-195, all on_*_edge flags set
 *                           the way fregrid passes them (fregrid_util.c:2197-2200)
 *   orc_grad_mask           fregrid_util.c:2203-2216
 *   orc_calc_c2l_grid_info  gradient_c2l.c:368-454 with get_edge :198-313, mid_pt_sphere :315-337 and the
 *                           mosaic_util.c helpers great_circle_distance :747, spherical_excess_area :846,
 *                           spherical_angle :800 (plain double + acosl, see orc_spherical_angle),
 *                           latlon2xyz :212, xyz2latlon :228, unit_vect_latlon :937
 * Pinned against the compiled reference (oracle/_ref) and golden vectors by tests/test_apply_cpu.py.
 */
#include <math.h>
#include <stdlib.h>
#include "xgrid_oracle.h"

#define ORC_RADIUS 6371000.0

/* ---- a2b_ord2: cell centres (with halo) -> corners ------------------------------------------- */
static void orc_corners(int nx, int ny, const double *q, const double *ew, const double *ee,
                        const double *es, const double *en, double *pb)
{
  const int w = nx + 2, nxp = nx + 1, nyp = ny + 1;
  const double third = 1./3.;
  int i, j;
  for (j = 1; j < ny; j++) for (i = 1; i < nx; i++)                                  /* :163-166 */
    pb[j*nxp+i] = 0.25*(q[j*w+i] + q[j*w+i+1] + q[(j+1)*w+i] + q[(j+1)*w+i+1]);
  pb[0]         = third*(q[w+1]    + q[w]        + q[1]);                            /* :169-172 */
  pb[nx]        = third*(q[w+nx]   + q[nx]       + q[w+nxp]);
  pb[ny*nxp+nx] = third*(q[ny*w+nx] + q[ny*w+nxp] + q[nyp*w+nx]);
  pb[ny*nxp]    = third*(q[ny*w+1]  + q[ny*w]     + q[nyp*w+1]);
  for (j = 1; j < ny; j++) {                                                         /* :175-184 */
    double a = 0.5*(q[j*w] + q[j*w+1]),       b = 0.5*(q[(j+1)*w] + q[(j+1)*w+1]);
    double c = 0.5*(q[j*w+nx] + q[j*w+nxp]),  d = 0.5*(q[(j+1)*w+nx] + q[(j+1)*w+nxp]);
    pb[j*nxp]    = ew[j]*a + (1-ew[j])*b;
    pb[j*nxp+nx] = ee[j]*c + (1-ee[j])*d;
  }
  for (i = 1; i < nx; i++) {                                                         /* :187-196 */
    double a = 0.5*(q[i] + q[w+i]),            b = 0.5*(q[i+1] + q[w+i+1]);
    double c = 0.5*(q[ny*w+i] + q[nyp*w+i]),   d = 0.5*(q[ny*w+i+1] + q[nyp*w+i+1]);
    pb[i]        = es[i]*a + (1-es[i])*b;
    pb[ny*nxp+i] = en[i]*c + (1-en[i])*d;
  }
}

void orc_grad_c2l(int nx, int ny, const double *pin, const double *dx, const double *dy,
                  const double *area, const double *edge_w, const double *edge_e,
                  const double *edge_s, const double *edge_n, const double *en_n, const double *en_e,
                  const double *vlon, const double *vlat, double *grad_x, double *grad_y)
{
  const int nxp = nx + 1, nyp = ny + 1;
  double *pb = (double *)malloc((size_t)nxp*nyp*sizeof(double));
  int i, j, n;
  orc_corners(nx, ny, pin, edge_w, edge_e, edge_s, edge_n, pb);
  for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
    const int m = j*nx + i;
    double g3[3];
    for (n = 0; n < 3; n++) {
      double fs = 0.5*(pb[j*nxp+i]     + pb[j*nxp+i+1])    *dx[m]        *en_n[3*m+n];            /* :86-92 */
      double fn = 0.5*(pb[(j+1)*nxp+i] + pb[(j+1)*nxp+i+1])*dx[(j+1)*nx+i]*en_n[3*((j+1)*nx+i)+n];
      double fw = 0.5*(pb[j*nxp+i]     + pb[(j+1)*nxp+i])  *dy[j*nxp+i]  *en_e[3*(j*nxp+i)+n];    /* :94-99 */
      double fe = 0.5*(pb[j*nxp+i+1]   + pb[(j+1)*nxp+i+1])*dy[j*nxp+i+1]*en_e[3*(j*nxp+i+1)+n];
      g3[n] = fn - fs - fw + fe;                                                                   /* :102-107 */
    }
    grad_x[m] = (vlon[3*m]*g3[0] + vlon[3*m+1]*g3[1] + vlon[3*m+2]*g3[2])/area[m];                 /* :110-117 */
    grad_x[m] *= ORC_RADIUS;
    grad_y[m] = (vlat[3*m]*g3[0] + vlat[3*m+1]*g3[1] + vlat[3*m+2]*g3[2])/area[m];
    grad_y[m] *= ORC_RADIUS;
  }
  free(pb);
}

void orc_grad_mask(int nx, int ny, const double *pin, double missing, int *mask)
{
  const int w = nx + 2;
  int i, j, di, dj;
  for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
    int m = 0;
    for (dj = 0; dj <= 2; dj++) for (di = 0; di <= 2; di++) {
      if (di == 1 && dj == 1) continue;
      if (pin[(j+dj)*w + i+di] == missing) m = 1;
    }
    mask[j*nx+i] = m;
  }
}

/* ---- metrics ---------------------------------------------------------------------------------- */
static void orc_ll2xyz(double lon, double lat, double v[3])
{
  v[0] = cos(lat)*cos(lon); v[1] = cos(lat)*sin(lon); v[2] = sin(lat);
}

static double orc_gc_distance(const double p1[2], const double p2[2])
{
  double beta = 2.*asin(sqrt(sin((p1[1]-p2[1])/2.)*sin((p1[1]-p2[1])/2.) +
                             cos(p1[1])*cos(p2[1])*(sin((p1[0]-p2[0])/2.)*sin((p1[0]-p2[0])/2.))));
  return ORC_RADIUS*beta;
}

static void orc_cross(const double a[3], const double b[3], double e[3])
{
  e[0] = a[1]*b[2] - a[2]*b[1];
  e[1] = a[2]*b[0] - a[0]*b[2];
  e[2] = a[0]*b[1] - a[1]*b[0];
}

static void orc_unit(double e[3])
{
  double n = sqrt(e[0]*e[0] + e[1]*e[1] + e[2]*e[2]);
  int k;
  for (k = 0; k < 3; k++) e[k] /= n;
}

double orc_spherical_angle(const double v1[3], const double v2[3], const double v3[3])
{
  /* plain double: HAVE_LONG_DOUBLE_WIDER never reaches mosaic_util.c (it does not include config.h); only the
   * final acosl() runs in x87 extended precision and is rounded back to double (mosaic_util.c:834) */
  double px = v1[1]*v2[2] - v1[2]*v2[1], py = v1[2]*v2[0] - v1[0]*v2[2], pz = v1[0]*v2[1] - v1[1]*v2[0];
  double qx = v1[1]*v3[2] - v1[2]*v3[1], qy = v1[2]*v3[0] - v1[0]*v3[2], qz = v1[0]*v3[1] - v1[1]*v3[0];
  double ddd = (px*px + py*py + pz*pz)*(qx*qx + qy*qy + qz*qz);
  double angle;
  if (ddd <= 0.0) return 0.;
  ddd = (px*qx + py*qy + pz*qz)/sqrt(ddd);
  if (fabs(ddd - 1) < 1.e-30) ddd = 1;
  if (fabs(ddd + 1) < 1.e-30) ddd = -1;
  if (ddd > 1. || ddd < -1.) angle = (ddd < 0.) ? M_PI : 0.;
  else angle = acosl(ddd);
  return angle;
}

static double orc_excess_area(const double ll[2], const double ul[2], const double lr[2], const double ur[2])
{
  double a[3], b[3], c[3], ang1, ang2, ang3, ang4;
  orc_ll2xyz(ll[0], ll[1], a); orc_ll2xyz(lr[0], lr[1], b); orc_ll2xyz(ul[0], ul[1], c);
  ang1 = orc_spherical_angle(a, b, c);
  orc_ll2xyz(lr[0], lr[1], a); orc_ll2xyz(ur[0], ur[1], b); orc_ll2xyz(ll[0], ll[1], c);
  ang2 = orc_spherical_angle(a, b, c);
  orc_ll2xyz(ur[0], ur[1], a); orc_ll2xyz(ul[0], ul[1], b); orc_ll2xyz(lr[0], lr[1], c);
  ang3 = orc_spherical_angle(a, b, c);
  orc_ll2xyz(ul[0], ul[1], a); orc_ll2xyz(ur[0], ur[1], b); orc_ll2xyz(ll[0], ll[1], c);
  ang4 = orc_spherical_angle(a, b, c);
  return (ang1 + ang2 + ang3 + ang4 - 2.*M_PI)*ORC_RADIUS*ORC_RADIUS;
}

static void orc_mid_point(double lon1, double lat1, double lon2, double lat2, double pm[2])
{
  double a[3], b[3], e[3], d;
  orc_ll2xyz(lon1, lat1, a); orc_ll2xyz(lon2, lat2, b);
  e[0] = a[0] + b[0]; e[1] = a[1] + b[1]; e[2] = a[2] + b[2];
  orc_unit(e);
  d = sqrt(e[0]*e[0] + e[1]*e[1] + e[2]*e[2]);                                    /* xyz2latlon renormalises */
  e[0] /= d; e[1] /= d; e[2] /= d;
  pm[0] = (fabs(e[0]) + fabs(e[1]) < 1.e-10) ? 0 : atan2(e[1], e[0]);
  pm[1] = asin(e[2]);
  if (pm[0] < 0.) pm[0] = 2.*M_PI + pm[0];
}

void orc_calc_c2l_grid_info(int nx, int ny, const double *xt, const double *yt,
                            const double *xc, const double *yc,
                            double *dx, double *dy, double *area,
                            double *edge_w, double *edge_e, double *edge_s, double *edge_n,
                            double *en_n, double *en_e, double *vlon, double *vlat)
{
  const int nxp = nx + 1, nyp = ny + 1, w = nx + 2;
  int i, j;
  for (j = 0; j < nyp; j++) for (i = 0; i < nxp; i++) {
    const int c = j*nxp + i;
    double p0[2], p1[2], v0[3], v1[3];
    p0[0] = xc[c]; p0[1] = yc[c];
    orc_ll2xyz(p0[0], p0[1], v0);
    if (i < nx) {
      p1[0] = xc[c+1]; p1[1] = yc[c+1];
      dx[j*nx+i] = orc_gc_distance(p0, p1);
      orc_ll2xyz(p1[0], p1[1], v1);
      orc_cross(v0, v1, en_n + 3*(j*nx+i));
      orc_unit(en_n + 3*(j*nx+i));
    }
    if (j < ny) {
      p1[0] = xc[c+nxp]; p1[1] = yc[c+nxp];
      dy[c] = orc_gc_distance(p0, p1);
      orc_ll2xyz(p1[0], p1[1], v1);
      orc_cross(v1, v0, en_e + 3*c);
      orc_unit(en_e + 3*c);
    }
    if (i < nx && j < ny) {
      const int m = j*nx + i;
      double ul[2], lr[2], ur[2];
      double lon = xt[(j+1)*w + i+1], lat = yt[(j+1)*w + i+1];
      ul[0] = xc[c+nxp]; ul[1] = yc[c+nxp]; lr[0] = xc[c+1]; lr[1] = yc[c+1]; ur[0] = xc[c+nxp+1]; ur[1] = yc[c+nxp+1];
      area[m] = orc_excess_area(p0, ul, lr, ur);
      vlon[3*m] = -sin(lon); vlon[3*m+1] = cos(lon); vlon[3*m+2] = 0.;
      vlat[3*m] = -sin(lat)*cos(lon); vlat[3*m+1] = -sin(lat)*sin(lon); vlat[3*m+2] = cos(lat);
    }
  }
  for (j = 0; j < nyp; j++) { edge_w[j] = 0.5; edge_e[j] = 0.5; }
  for (i = 0; i < nxp; i++) { edge_s[i] = 0.5; edge_n[i] = 0.5; }
  for (j = 1; j < ny; j++) {
    double m0[2], m1[2], c[2], d1, d2;
    orc_mid_point(xt[j*w], yt[j*w], xt[j*w+1], yt[j*w+1], m0);
    orc_mid_point(xt[(j+1)*w], yt[(j+1)*w], xt[(j+1)*w+1], yt[(j+1)*w+1], m1);
    c[0] = xc[j*nxp]; c[1] = yc[j*nxp];
    d1 = orc_gc_distance(m0, c); d2 = orc_gc_distance(m1, c);
    edge_w[j] = d2/(d1+d2);
    orc_mid_point(xt[j*w+nx], yt[j*w+nx], xt[j*w+nxp], yt[j*w+nxp], m0);
    orc_mid_point(xt[(j+1)*w+nx], yt[(j+1)*w+nx], xt[(j+1)*w+nxp], yt[(j+1)*w+nxp], m1);
    c[0] = xc[j*nxp+nx]; c[1] = yc[j*nxp+nx];
    d1 = orc_gc_distance(m0, c); d2 = orc_gc_distance(m1, c);
    edge_e[j] = d2/(d1+d2);
  }
  for (i = 1; i < nx; i++) {
    double m0[2], m1[2], c[2], d1, d2;
    orc_mid_point(xt[i], yt[i], xt[w+i], yt[w+i], m0);
    orc_mid_point(xt[i+1], yt[i+1], xt[w+i+1], yt[w+i+1], m1);
    c[0] = xc[i]; c[1] = yc[i];
    d1 = orc_gc_distance(m0, c); d2 = orc_gc_distance(m1, c);
    edge_s[i] = d2/(d1+d2);
    orc_mid_point(xt[ny*w+i], yt[ny*w+i], xt[nyp*w+i], yt[nyp*w+i], m0);
    orc_mid_point(xt[ny*w+i+1], yt[ny*w+i+1], xt[nyp*w+i+1], yt[nyp*w+i+1], m1);
    c[0] = xc[ny*nxp+i]; c[1] = yc[ny*nxp+i];
    d1 = orc_gc_distance(m0, c); d2 = orc_gc_distance(m1, c);
    edge_n[i] = d2/(d1+d2);
  }
}
