/* TEST INFRASTRUCTURE — CPU oracle (see xgrid_oracle.h for the contract).
 *
 * Restates, in plain C, the reference's legacy ("2dx2d") exchange-grid path and the
 * conservative apply.  Arithmetic follows the reference operation-for-operation (same
 * association order, no FMA: build with -ffp-contract=off) so that accept/reject decisions
 * and areas are reproducible; only the candidate search is restructured (row / block pruning
 * that provably skips pairs the reference's own bounding-box tests would reject), which
 * leaves the emitted list and its order unchanged.
 *
 * Reference citations are relative to /root/reference.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "xgrid_oracle.h"

#define ORC_PI      3.14159265358979323846
#define ORC_TPI     (2.0*ORC_PI)
#define ORC_HPI     (0.5*ORC_PI)
#define ORC_RADIUS  6371000.0                 /* constant.h:23 */
#define ORC_SMALL   1.e-10                    /* mosaic_util.h:34 SMALL_VALUE */
#define ORC_POLE_TOL 1.e-6                    /* mosaic_util.c:35 TOLORENCE */
#define ORC_MAXV    8                         /* create_xgrid.c:627 MAX_V */
#define ORC_MV      50                        /* create_xgrid.h:31 MV */
#define ORC_AREA_RATIO_THRESH 1.e-6           /* create_xgrid.c:27 */
#define ORC_MASK_THRESH 0.5                   /* create_xgrid.c:28 */

/* Optional event counters (build with -DORC_COUNT_OPS -> liboracle_xgrid_ops.so, see oracle/count_ops.py):
 * they count how often each primitive step runs so that DESIGN.md can state the algorithmic FP64
 * operations per exchange cell from the source-level operation count of each step. */
enum { CNT_CLIP_CALLS, CNT_INSIDE_EDGE, CNT_INTERSECT, CNT_AREA_CALLS, CNT_AREA_EDGE_GEN, CNT_AREA_EDGE_FLAT,
       CNT_AREA_EDGE_POLE, CNT_RATIO_TESTS, CNT_ACCEPTED, CNT_CTRLON_EDGE, CNT_CTRLAT_EDGE_GEN, CNT_CTRLAT_EDGE_FLAT,
       CNT_CLIP_NONEMPTY, CNT_N };
#ifdef ORC_COUNT_OPS
static long orc_cnt[CNT_N];
#define ORC_COUNT(k) (orc_cnt[k]++)
void orc_counts_get(long *out) { int k; for (k = 0; k < CNT_N; k++) out[k] = orc_cnt[k]; }
void orc_counts_reset(void) { int k; for (k = 0; k < CNT_N; k++) orc_cnt[k] = 0; }
#else
#define ORC_COUNT(k) ((void)0)
#endif

static void orc_die(const char *msg)
{
  fprintf(stderr, "xgrid_oracle FATAL: %s\n", msg);
  exit(1);
}

/* ------------------------------------------------------------------------------------------
 * fix_lon — mosaic_util.c:667-738.  Pairs pole vertices, inserts twin pole vertices for an
 * edge crossing a pole, unwraps longitudes to be contiguous and within pi of tlon on average.
 * ---------------------------------------------------------------------------------------- */
static int orc_vtx_remove(double x[], double y[], int n, int at)      /* mosaic_util.c:636 */
{
  int k;
  for (k = at; k < n-1; k++) { x[k] = x[k+1]; y[k] = y[k+1]; }
  return n-1;
}

static int orc_vtx_insert(double x[], double y[], int n, int at, double lon, double lat) /* :646 */
{
  int k;
  for (k = n-1; k >= at; k--) { x[k+1] = x[k]; y[k+1] = y[k]; }
  x[at] = lon; y[at] = lat;
  return n+1;
}

int orc_fix_lon(double x[], double y[], int n, double tlon)
{
  const double near_pole = ORC_HPI - ORC_POLE_TOL;
  double sum, shift;
  int nn = n, i;

  /* every pole vertex must appear exactly twice in a row (mosaic_util.c:679-692) */
  for (i = 0; i < nn; i++) {
    if (fabs(y[i]) >= near_pole) {
      int prev = (i+nn-1)%nn, next = (i+1)%nn;
      if (y[prev] == y[i] && y[next] == y[i]) { nn = orc_vtx_remove(x, y, nn, i); i--; }
      else if (y[prev] != y[i] && y[next] != y[i]) { nn = orc_vtx_insert(x, y, nn, i, x[i], y[i]); i++; }
    }
  }
  /* pole pair takes the longitudes of its non-pole neighbours (mosaic_util.c:693-700) */
  for (i = 0; i < nn; i++) {
    if (fabs(y[i]) >= near_pole) {
      int prev = (i+nn-1)%nn, next = (i+1)%nn;
      if (y[prev] != y[i]) x[i] = x[prev];
      if (y[next] != y[i]) x[i] = x[next];
    }
  }
  /* an edge whose longitudes differ by pi runs through a pole: add twin pole vertices (:702-717) */
  for (i = 0; i < nn; i++) {
    int prev = (i+nn-1)%nn;
    double d = x[i] - x[prev];
    if (fabs(d + ORC_PI) < ORC_SMALL || fabs(d - ORC_PI) < ORC_SMALL) {
      double xa = x[prev], xb = x[i];
      double yp = (y[i] < 0.0) ? -ORC_HPI : ORC_HPI;
      nn = orc_vtx_insert(x, y, nn, i, xb, yp);
      nn = orc_vtx_insert(x, y, nn, i, xa, yp);
      break;
    }
  }
  if (nn == 0) return 0;
  /* unwrap so that consecutive vertices differ by at most pi (:718-725) */
  sum = x[0];
  for (i = 1; i < nn; i++) {
    double d = x[i] - x[i-1];
    if (d < -ORC_PI) d = d + ORC_TPI;
    else if (d > ORC_PI) d = d - ORC_TPI;
    x[i] = x[i-1] + d;
    sum += x[i];
  }
  /* bring the mean within pi of tlon (:727-729) */
  shift = (sum/nn) - tlon;
  if (shift < -ORC_PI)      for (i = 0; i < nn; i++) x[i] += ORC_TPI;
  else if (shift > ORC_PI)  for (i = 0; i < nn; i++) x[i] -= ORC_TPI;
  return nn;
}

/* ------------------------------------------------------------------------------------------
 * poly_area — mosaic_util.c:417-459 (poly_area_main; the pole-rotation branch of :474-515 is
 * off unless --rotate_poly, fregrid never sets it).  area = | -sum dlon * sin(mid lat) * sinc |
 * ---------------------------------------------------------------------------------------- */
double orc_poly_area(const double x[], const double y[], int n)
{
  double acc = 0.0;
  int i;
  ORC_COUNT(CNT_AREA_CALLS);
  for (i = 0; i < n; i++) {
    int ip = (i+1)%n;
    double dx = x[ip] - x[i];
    double lat1 = y[ip], lat2 = y[i];
    if (dx > ORC_PI)  dx = dx - 2.0*ORC_PI;
    if (dx < -ORC_PI) dx = dx + 2.0*ORC_PI;
    if (fabs(dx + ORC_PI) < ORC_SMALL || fabs(dx - ORC_PI) < ORC_SMALL) {
      acc += ORC_PI;                       /* side through a pole (:434-437) */
      ORC_COUNT(CNT_AREA_EDGE_POLE);
      continue;
    }
    if (fabs(lat1 - lat2) < ORC_SMALL) {
      ORC_COUNT(CNT_AREA_EDGE_FLAT);
      acc -= dx * sin(0.5*(lat1 + lat2));
    } else {
      ORC_COUNT(CNT_AREA_EDGE_GEN);
      double dy = 0.5*(lat1 - lat2);
      double dat = sin(dy)/dy;
      acc -= dx * sin(0.5*(lat1 + lat2)) * dat;
    }
  }
  if (acc < 0) return -acc*ORC_RADIUS*ORC_RADIUS;
  return acc*ORC_RADIUS*ORC_RADIUS;
}

/* poly_ctrlat — create_xgrid.c:2096-2121 */
double orc_poly_ctrlat(const double x[], const double y[], int n)
{
  double acc = 0.0;
  int i;
  for (i = 0; i < n; i++) {
    int ip = (i+1)%n;
    double dx = x[ip] - x[i];
    double lat1 = y[ip], lat2 = y[i];
    double dy = lat2 - lat1;
    double hdy = dy*0.5;
    double avg_y = (lat1 + lat2)*0.5;
    if (dx == 0.0) continue;
    if (dx > ORC_PI)   dx = dx - 2.0*ORC_PI;
    if (dx <= -ORC_PI) dx = dx + 2.0*ORC_PI;
    if (fabs(hdy) < ORC_SMALL) {
      ORC_COUNT(CNT_CTRLAT_EDGE_FLAT);
      acc -= dx*(2*cos(avg_y) + lat2*sin(avg_y) - cos(lat1));
    } else {
      ORC_COUNT(CNT_CTRLAT_EDGE_GEN);
      acc -= dx*((sin(hdy)/hdy)*(2*cos(avg_y) + lat2*sin(avg_y)) - cos(lat1));
    }
  }
  return acc*ORC_RADIUS*ORC_RADIUS;
}

/* poly_ctrlon — create_xgrid.c:2170-2217 */
double orc_poly_ctrlon(const double x[], const double y[], int n, double clon)
{
  double acc = 0.0;
  int i;
  for (i = 0; i < n; i++) {
    int ip = (i+1)%n;
    double phi1 = x[ip], phi2 = x[i], lat1 = y[ip], lat2 = y[i];
    double dphi = phi1 - phi2, dphi1, dphi2, f1, f2;
    if (dphi == 0.0) continue;
    ORC_COUNT(CNT_CTRLON_EDGE);
    f1 = 0.5*(cos(lat1)*sin(lat1) + lat1);
    f2 = 0.5*(cos(lat2)*sin(lat2) + lat2);
    if (dphi > ORC_PI)  dphi = dphi - 2.0*ORC_PI;
    if (dphi < -ORC_PI) dphi = dphi + 2.0*ORC_PI;
    dphi1 = phi1 - clon;
    if (dphi1 > ORC_PI)  dphi1 -= 2.0*ORC_PI;
    if (dphi1 < -ORC_PI) dphi1 += 2.0*ORC_PI;
    dphi2 = phi2 - clon;
    if (dphi2 > ORC_PI)  dphi2 -= 2.0*ORC_PI;
    if (dphi2 < -ORC_PI) dphi2 += 2.0*ORC_PI;
    if (fabs(dphi2 - dphi1) < ORC_PI) {
      acc -= dphi * (dphi1*f1 + dphi2*f2)/2.0;
    } else {
      double fac = (dphi1 > 0.0) ? ORC_PI : -ORC_PI;
      double fint = f1 + (f2 - f1)*(fac - dphi1)/fabs(dphi);
      acc -= 0.5*dphi1*(dphi1 - fac)*f1 - 0.5*dphi2*(dphi2 + fac)*f2 + 0.5*fac*(dphi1 + dphi2)*fint;
    }
  }
  return acc*ORC_RADIUS*ORC_RADIUS;
}

/* inside_edge — create_xgrid.c:2342-2350: point on or left-of-normal of the directed edge */
static int orc_inside_edge(double x0, double y0, double x1, double y1, double x, double y)
{
  double product = (x - x0)*(y1 - y0) + (x0 - x1)*(y - y0);
  ORC_COUNT(CNT_INSIDE_EDGE);
  return (product <= 1.e-12) ? 1 : 0;
}

/* clip_2dx2d — create_xgrid.c:1266-1341.  Polygon 1 is cut by every edge of polygon 2. */
int orc_clip_2dx2d(const double lon1[], const double lat1[], int n1,
                   const double lon2[], const double lat2[], int n2,
                   double lon_out[], double lat_out[])
{
  double px[ORC_MV], py[ORC_MV], qx[ORC_MV], qy[ORC_MV];
  int np = n1, k, e, wrap = 0;
  ORC_COUNT(CNT_CLIP_CALLS);

  for (k = 0; k < n1; k++) {
    px[k] = lon1[k]; py[k] = lat1[k];
    if (px[k] > ORC_TPI || px[k] < 0.0) wrap = 1;
  }
  for (k = 0; k < n2; k++) { qx[k] = lon2[k]; qy[k] = lat2[k]; }
  if (wrap) {                               /* pimod heuristic, :1282-1290 and :1343-1349 */
    for (k = 0; k < n1; k++) { if (px[k] < -ORC_PI) px[k] += ORC_TPI; else if (px[k] > ORC_PI) px[k] -= ORC_TPI; }
    for (k = 0; k < n2; k++) { if (qx[k] < -ORC_PI) qx[k] += ORC_TPI; else if (qx[k] > ORC_PI) qx[k] -= ORC_TPI; }
  }

  {
    double ex0 = qx[n2-1], ey0 = qy[n2-1];
    for (e = 0; e < n2; e++) {
      double ex1 = qx[e], ey1 = qy[e];
      double ax = px[np-1], ay = py[np-1];
      int was_in = orc_inside_edge(ex0, ey0, ex1, ey1, ax, ay);
      int no = 0;
      for (k = 0; k < np; k++) {
        double bx = px[k], by = py[k];
        int is_in = orc_inside_edge(ex0, ey0, ex1, ey1, bx, by);
        if (is_in != was_in) {
          double dy1 = by - ay, dy2 = ey1 - ey0, dx1 = bx - ax, dx2 = ex1 - ex0;
          double ds1 = ay*bx - by*ax, ds2 = ey0*ex1 - ey1*ex0;
          double determ = dy2*dx1 - dy1*dx2;
          ORC_COUNT(CNT_INTERSECT);
          if (fabs(determ) < 1.0e-30) orc_die("clip_2dx2d: parallel edges");
          lon_out[no]   = (dx2*ds1 - dx1*ds2)/determ;
          lat_out[no++] = (dy2*ds1 - dy1*ds2)/determ;
        }
        if (is_in) { lon_out[no] = bx; lat_out[no++] = by; }
        ax = bx; ay = by; was_in = is_in;
      }
      np = no;
      if (np == 0) return 0;
      for (k = 0; k < np; k++) { px[k] = lon_out[k]; py[k] = lat_out[k]; }
      ex0 = ex1; ey0 = ey1;
    }
  }
  return np;
}

/* one grid cell -> fix_lon'd polygon (create_xgrid.c:75-85, :757-768) */
static int orc_cell_poly(const double *lon, const double *lat, int nxp, int i, int j, double x[], double y[],
                         double *ymin, double *ymax)
{
  int n0 = j*nxp + i, n1 = n0 + 1, n2 = (j+1)*nxp + i + 1, n3 = (j+1)*nxp + i, k;
  x[0] = lon[n0]; y[0] = lat[n0];
  x[1] = lon[n1]; y[1] = lat[n1];
  x[2] = lon[n2]; y[2] = lat[n2];
  x[3] = lon[n3]; y[3] = lat[n3];
  if (ymin) {
    double lo = y[0], hi = y[0];
    for (k = 1; k < 4; k++) { if (y[k] < lo) lo = y[k]; if (y[k] > hi) hi = y[k]; }
    *ymin = lo; *ymax = hi;
  }
  return orc_fix_lon(x, y, 4, ORC_PI);
}

/* get_grid_area — create_xgrid.c:66-88 */
void orc_get_grid_area(int nlon, int nlat, const double *lon, const double *lat, double *area)
{
  int i, j, nxp = nlon + 1;
  double x[20], y[20];
  for (j = 0; j < nlat; j++) for (i = 0; i < nlon; i++) {
    int n = orc_cell_poly(lon, lat, nxp, i, j, x, y, NULL, NULL);
    area[(size_t)j*nlon + i] = orc_poly_area(x, y, n);
  }
}

static void orc_minmaxavg(const double *v, int n, double *lo, double *hi, double *avg)
{
  double a = v[0], b = v[0], s = 0; int k;
  for (k = 1; k < n; k++) { if (v[k] < a) a = v[k]; if (v[k] > b) b = v[k]; }
  for (k = 0; k < n; k++) s += v[k];        /* avgval_double sums from 0 (mosaic_util.c:199-201) */
  *lo = a; *hi = b; *avg = s/n;
}

/* ------------------------------------------------------------------------------------------
 * create_xgrid_2dx2d_order1 / _order2 — create_xgrid.c:621-871 / :893-1152.
 * Emission order: source cells row-major (j1,i1), then destination index ij ascending.
 * ---------------------------------------------------------------------------------------- */
#define ORC_BLK 64
long orc_create_xgrid_2dx2d(int order, int nlon_in, int nlat_in, int nlon_out, int nlat_out,
                            const double *lon_in, const double *lat_in,
                            const double *lon_out, const double *lat_out, const double *mask_in,
                            long cap, int *i_in, int *j_in, int *i_out, int *j_out,
                            double *xarea, double *xclon, double *xclat)
{
  const int nx1 = nlon_in, ny1 = nlat_in, nx2 = nlon_out, ny2 = nlat_out;
  const int nx1p = nx1 + 1, nx2p = nx2 + 1;
  const size_t n2cells = (size_t)nx2*ny2;
  const int nblk = (nx2 + ORC_BLK - 1)/ORC_BLK;
  double *area_in = (double *)malloc((size_t)nx1*ny1*sizeof(double));
  double *area_out = (double *)malloc(n2cells*sizeof(double));
  double *o_ymin = (double *)malloc(n2cells*sizeof(double)), *o_ymax = (double *)malloc(n2cells*sizeof(double));
  double *o_xmin = (double *)malloc(n2cells*sizeof(double)), *o_xmax = (double *)malloc(n2cells*sizeof(double));
  double *o_xavg = (double *)malloc(n2cells*sizeof(double));
  int    *o_n    = (int *)malloc(n2cells*sizeof(int));
  double *o_x    = (double *)malloc(n2cells*ORC_MAXV*sizeof(double));
  double *o_y    = (double *)malloc(n2cells*ORC_MAXV*sizeof(double));
  double *row_lo = (double *)malloc((size_t)ny2*sizeof(double)), *row_hi = (double *)malloc((size_t)ny2*sizeof(double));
  double *b_ylo = (double *)malloc((size_t)ny2*nblk*sizeof(double)), *b_yhi = (double *)malloc((size_t)ny2*nblk*sizeof(double));
  double *b_xlo = (double *)malloc((size_t)ny2*nblk*sizeof(double)), *b_xhi = (double *)malloc((size_t)ny2*nblk*sizeof(double));
  long nxgrid = 0;
  int i1, j1, i2, j2, k;

  orc_get_grid_area(nx1, ny1, lon_in, lat_in, area_in);
  orc_get_grid_area(nx2, ny2, lon_out, lat_out, area_out);

  /* destination cell precompute, create_xgrid.c:714-739 */
  for (j2 = 0; j2 < ny2; j2++) {
    row_lo[j2] = 1e300; row_hi[j2] = -1e300;
    for (k = 0; k < nblk; k++) {
      b_ylo[(size_t)j2*nblk+k] = 1e300; b_yhi[(size_t)j2*nblk+k] = -1e300;
      b_xlo[(size_t)j2*nblk+k] = 1e300; b_xhi[(size_t)j2*nblk+k] = -1e300;
    }
    for (i2 = 0; i2 < nx2; i2++) {
      size_t n = (size_t)j2*nx2 + i2, b = (size_t)j2*nblk + i2/ORC_BLK;
      double x[ORC_MV], y[ORC_MV];
      int nv = orc_cell_poly(lon_out, lat_out, nx2p, i2, j2, x, y, &o_ymin[n], &o_ymax[n]);
      if (nv > ORC_MAXV) orc_die("create_xgrid: n2_in is greater than MAX_V");
      orc_minmaxavg(x, nv, &o_xmin[n], &o_xmax[n], &o_xavg[n]);
      o_n[n] = nv;
      for (k = 0; k < nv; k++) { o_x[n*ORC_MAXV+k] = x[k]; o_y[n*ORC_MAXV+k] = y[k]; }
      if (o_ymin[n] < row_lo[j2]) row_lo[j2] = o_ymin[n];
      if (o_ymax[n] > row_hi[j2]) row_hi[j2] = o_ymax[n];
      if (o_ymin[n] < b_ylo[b]) b_ylo[b] = o_ymin[n];
      if (o_ymax[n] > b_yhi[b]) b_yhi[b] = o_ymax[n];
      if (o_xmin[n] < b_xlo[b]) b_xlo[b] = o_xmin[n];
      if (o_xmax[n] > b_xhi[b]) b_xhi[b] = o_xmax[n];
    }
  }

  for (j1 = 0; j1 < ny1; j1++) for (i1 = 0; i1 < nx1; i1++) {
    double x1[ORC_MV], y1[ORC_MV], s_ymin, s_ymax, s_xmin, s_xmax, s_xavg;
    int n1;
    if (!(mask_in[(size_t)j1*nx1 + i1] > ORC_MASK_THRESH)) continue;
    n1 = orc_cell_poly(lon_in, lat_in, nx1p, i1, j1, x1, y1, &s_ymin, &s_ymax);
    orc_minmaxavg(x1, n1, &s_xmin, &s_xmax, &s_xavg);

    for (j2 = 0; j2 < ny2; j2++) {
      int b;
      /* whole-row prune: every cell of the row would fail the test at create_xgrid.c:777 */
      if (row_lo[j2] >= s_ymax || row_hi[j2] <= s_ymin) continue;
      for (b = 0; b < nblk; b++) {
        size_t bb = (size_t)j2*nblk + b;
        int ilo = b*ORC_BLK, ihi = ilo + ORC_BLK;
        int hit = 0;
        if (ihi > nx2) ihi = nx2;
        if (b_ylo[bb] >= s_ymax || b_yhi[bb] <= s_ymin) continue;
        /* block-level longitude prune, conservative over the three possible 2*pi shifts (:786-801) */
        for (k = -1; k <= 1 && !hit; k++) {
          double lo = b_xlo[bb], hi = b_xhi[bb];
          if (k < 0) { lo -= ORC_TPI; hi -= ORC_TPI; }
          if (k > 0) { lo += ORC_TPI; hi += ORC_TPI; }
          if (!(lo >= s_xmax || hi <= s_xmin)) hit = 1;
        }
        if (!hit) continue;
        for (i2 = ilo; i2 < ihi; i2++) {
          size_t ij = (size_t)j2*nx2 + i2;
          double x2[ORC_MAXV], y2[ORC_MAXV], xo[ORC_MV], yo[ORC_MV];
          double xmin2, xmax2, dx;
          int n2, no;
          if (o_ymin[ij] >= s_ymax || o_ymax[ij] <= s_ymin) continue;      /* :777 */
          n2 = o_n[ij];
          for (k = 0; k < n2; k++) { x2[k] = o_x[ij*ORC_MAXV+k]; y2[k] = o_y[ij*ORC_MAXV+k]; }
          xmin2 = o_xmin[ij]; xmax2 = o_xmax[ij];
          dx = o_xavg[ij] - s_xavg;                                          /* :786-796 */
          if (dx < -ORC_PI) {
            xmin2 += ORC_TPI; xmax2 += ORC_TPI;
            for (k = 0; k < n2; k++) x2[k] += ORC_TPI;
          } else if (dx > ORC_PI) {
            xmin2 -= ORC_TPI; xmax2 -= ORC_TPI;
            for (k = 0; k < n2; k++) x2[k] -= ORC_TPI;
          }
          if (xmin2 >= s_xmax || xmax2 <= s_xmin) continue;                  /* :801 */
          no = orc_clip_2dx2d(x1, y1, n1, x2, y2, n2, xo, yo);               /* :802 */
          if (no > 0) {
            ORC_COUNT(CNT_CLIP_NONEMPTY); ORC_COUNT(CNT_RATIO_TESTS);
            double xa = orc_poly_area(xo, yo, no) * mask_in[(size_t)j1*nx1 + i1];
            double a1 = area_in[(size_t)j1*nx1 + i1], a2 = area_out[ij];
            double min_area = (a1 < a2) ? a1 : a2;
            if (xa/min_area > ORC_AREA_RATIO_THRESH) {                       /* :807 */
              if (nxgrid >= cap) { nxgrid = -1; goto done; }
              ORC_COUNT(CNT_ACCEPTED);
              xarea[nxgrid] = xa;
              if (order == 2) {                                              /* :1091-1092 */
                xclon[nxgrid] = orc_poly_ctrlon(xo, yo, no, s_xavg);
                xclat[nxgrid] = orc_poly_ctrlat(xo, yo, no);
              }
              i_in[nxgrid] = i1; j_in[nxgrid] = j1; i_out[nxgrid] = i2; j_out[nxgrid] = j2;
              nxgrid++;
            }
          }
        }
      }
    }
  }
done:
  free(area_in); free(area_out); free(o_ymin); free(o_ymax); free(o_xmin); free(o_xmax); free(o_xavg);
  free(o_n); free(o_x); free(o_y); free(row_lo); free(row_hi); free(b_ylo); free(b_yhi); free(b_xlo); free(b_xhi);
  return nxgrid;
}

/* ------------------------------------------------------------------------------------------
 * setup_conserve_interp — conserve_interp.c:127-367 (generate branch), single destination
 * tile, serial (npes == 1) so the mpp gathers at :204-215 are identities.
 * ---------------------------------------------------------------------------------------- */
long orc_setup_conserve_interp_ex(int ntiles_in, const int *nx_in, const int *ny_in,
                                  const double *lonc_in, const double *latc_in,
                                  int nx_out, int ny_out, const double *lonc_out, const double *latc_out,
                                  unsigned int opcode, long cap,
                                  int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                                  double *area, double *di, double *dj, double *xclon_out, double *xclat_out)
{
  const int order = (opcode & ORC_CONSERVE_ORDER2) ? 2 : 1;
  long total = 0;
  size_t off = 0, offc = 0, ncell_tot = 0;
  size_t *cell_off = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  double *c_area = NULL, *c_clon = NULL, *c_clat = NULL, *xclon = NULL, *xclat = NULL;
  double y_min, y_max;
  int m;
  size_t q;

  for (m = 0; m < ntiles_in; m++) { cell_off[m] = ncell_tot; ncell_tot += (size_t)nx_in[m]*ny_in[m]; }
  if (order == 2) {
    c_area = (double *)calloc(ncell_tot, sizeof(double));
    c_clon = (double *)calloc(ncell_tot, sizeof(double));
    c_clat = (double *)calloc(ncell_tot, sizeof(double));
    xclon = (double *)malloc((size_t)cap*sizeof(double));
    xclat = (double *)malloc((size_t)cap*sizeof(double));
  }
  y_min = y_max = latc_out[0];
  for (q = 1; q < (size_t)(nx_out+1)*(ny_out+1); q++) {                       /* :169-170 */
    if (latc_out[q] < y_min) y_min = latc_out[q];
    if (latc_out[q] > y_max) y_max = latc_out[q];
  }

  for (m = 0; m < ntiles_in; m++) {
    const int nx = nx_in[m], ny = ny_in[m];
    const double *lon = lonc_in + off, *lat = latc_in + off;
    double *mask = (double *)malloc((size_t)nx*ny*sizeof(double));
    long n, k;
    for (q = 0; q < (size_t)nx*ny; q++) mask[q] = 1.0;                        /* :160-161 */

    if (opcode & ORC_GREAT_CIRCLE) {                                          /* :163-167 */
      n = orc_create_xgrid_great_circle(nx, ny, nx_out, ny_out, lon, lat, lonc_out, latc_out, mask,
                                        cap - total, i_in+total, j_in+total, i_out+total, j_out+total,
                                        area+total, xclon ? xclon+total : NULL, xclat ? xclat+total : NULL);
    } else {
      int jstart = ny, jend = -1, i, j, ny_now;                               /* :171-184 */
      for (j = 0; j <= ny; j++) for (i = 0; i <= nx; i++) {
        double yy = lat[(size_t)j*(nx+1) + i];
        if (yy > y_min) { if (j < jstart) jstart = j; }
        if (yy < y_max) { if (j > jend) jend = j; }
      }
      jstart = (jstart-1 > 0) ? jstart-1 : 0;
      jend = (jend+1 < ny-1) ? jend+1 : ny-1;
      ny_now = jend - jstart + 1;
      n = orc_create_xgrid_2dx2d(order, nx, ny_now, nx_out, ny_out,
                                 lon + (size_t)jstart*(nx+1), lat + (size_t)jstart*(nx+1),
                                 lonc_out, latc_out, mask, cap - total,
                                 i_in+total, j_in+total, i_out+total, j_out+total, area+total,
                                 xclon ? xclon+total : NULL, xclat ? xclat+total : NULL);
      for (k = 0; k < n; k++) j_in[total+k] += jstart;                        /* :190,:200 */
    }
    free(mask);
    if (n < 0) { total = -1; goto done; }
    for (k = 0; k < n; k++) t_in[total+k] = m;
    if (order == 2 && !(opcode & ORC_GREAT_CIRCLE)) {
      for (k = 0; k < n; k++) {                                               /* :216-221 */
        size_t ii = cell_off[m] + (size_t)j_in[total+k]*nx + i_in[total+k];
        c_area[ii] += area[total+k];
        c_clon[ii] += xclon[total+k];
        c_clat[ii] += xclat[total+k];
      }
    }
    if (order == 2) {
      for (k = 0; k < n; k++) {                                               /* :256-257, :303-304 */
        di[total+k] = xclon[total+k]/area[total+k];
        dj[total+k] = xclat[total+k]/area[total+k];
      }
    }
    total += n;
    off += (size_t)(nx+1)*(ny+1);
  }

  if (order == 2) {                                                           /* :319-358 */
    off = 0; offc = 0;
    for (m = 0; m < ntiles_in; m++) {
      const int nx = nx_in[m], ny = ny_in[m];
      const double *lon = lonc_in + off, *lat = latc_in + off;
      double *cell_area = (double *)malloc((size_t)nx*ny*sizeof(double));
      int i, j;
      orc_get_grid_area(nx, ny, lon, lat, cell_area);     /* grid_in[n].cell_area, fregrid_util.c:388 */
      for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
        size_t ii = (size_t)j*nx + i, g = offc + ii;
        if (c_area[g] > 0) {
          if (fabs(c_area[g] - cell_area[ii])/cell_area[ii] < 1.e-3) {        /* AREA_RATIO :35,:330 */
            c_clon[g] /= c_area[g];
            c_clat[g] /= c_area[g];
          } else {
            double x[ORC_MV], y[ORC_MV], lo, hi, avg;
            int nv = orc_cell_poly(lon, lat, nx+1, i, j, x, y, NULL, NULL);
            orc_minmaxavg(x, nv, &lo, &hi, &avg);
            c_clon[g] = orc_poly_ctrlon(x, y, nv, avg)/cell_area[ii];
            c_clat[g] = orc_poly_ctrlat(x, y, nv)/cell_area[ii];
          }
        }
      }
      free(cell_area);
      off += (size_t)(nx+1)*(ny+1);
      offc += (size_t)nx*ny;
    }
    {
      long k;
      for (k = 0; k < total; k++) {                                           /* :351-358 */
        size_t ii = cell_off[t_in[k]] + (size_t)j_in[k]*nx_in[t_in[k]] + i_in[k];
        di[k] -= c_clon[ii];
        dj[k] -= c_clat[ii];
      }
    }
  }
done:
  if (order == 2 && total > 0) {                      /* the generators' raw xgrid_clon / xgrid_clat, for the checkers */
    if (xclon_out) memcpy(xclon_out, xclon, (size_t)total*sizeof(double));
    if (xclat_out) memcpy(xclat_out, xclat, (size_t)total*sizeof(double));
  }
  free(cell_off); free(c_area); free(c_clon); free(c_clat); free(xclon); free(xclat);
  return total;
}

long orc_setup_conserve_interp(int ntiles_in, const int *nx_in, const int *ny_in,
                               const double *lonc_in, const double *latc_in,
                               int nx_out, int ny_out, const double *lonc_out, const double *latc_out,
                               unsigned int opcode, long cap,
                               int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                               double *area, double *di, double *dj)
{
  return orc_setup_conserve_interp_ex(ntiles_in, nx_in, ny_in, lonc_in, latc_in, nx_out, ny_out, lonc_out, latc_out, opcode, cap,
                                      t_in, i_in, j_in, i_out, j_out, area, di, dj, NULL, NULL);
}

/* The order-2 centroid correction alone (conserve_interp.c:204-221, :256-257, :319-358) for a finished list of ONE output
 * tile: per source cell the sums of area / xgrid_clon / xgrid_clat in list order, the AREA_RATIO test against the cell's
 * own area, the analytic centroid otherwise, then di = xgrid_clon/area - centroid.  Lets a full-size list produced on the
 * GPU (whose areas and raw centroids are checked against the reference on sampled destination bands) be checked for
 * tile1_distance without running the whole reference generator. */
void orc_order2_distance(int ntiles_in, const int *nx_in, const int *ny_in, const double *lonc_in, const double *latc_in,
                         long n, const int *t_in, const int *i_in, const int *j_in,
                         const double *area, const double *xclon, const double *xclat, double *di, double *dj)
{
  size_t ncell_tot = 0, off = 0, offc = 0;
  size_t *cell_off = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  double *c_area, *c_clon, *c_clat;
  int m;
  long k;
  for (m = 0; m < ntiles_in; m++) { cell_off[m] = ncell_tot; ncell_tot += (size_t)nx_in[m]*ny_in[m]; }
  c_area = (double *)calloc(ncell_tot, sizeof(double));
  c_clon = (double *)calloc(ncell_tot, sizeof(double));
  c_clat = (double *)calloc(ncell_tot, sizeof(double));
  for (k = 0; k < n; k++) {                                                   /* :216-221 */
    size_t ii = cell_off[t_in[k]] + (size_t)j_in[k]*nx_in[t_in[k]] + i_in[k];
    c_area[ii] += area[k]; c_clon[ii] += xclon[k]; c_clat[ii] += xclat[k];
    di[k] = xclon[k]/area[k];                                                 /* :256-257 */
    dj[k] = xclat[k]/area[k];
  }
  for (m = 0; m < ntiles_in; m++) {                                           /* :319-349 */
    const int nx = nx_in[m], ny = ny_in[m];
    const double *lon = lonc_in + off, *lat = latc_in + off;
    double *cell_area = (double *)malloc((size_t)nx*ny*sizeof(double));
    int i, j;
    orc_get_grid_area(nx, ny, lon, lat, cell_area);
    for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
      size_t ii = (size_t)j*nx + i, g = offc + ii;
      if (c_area[g] > 0) {
        if (fabs(c_area[g] - cell_area[ii])/cell_area[ii] < 1.e-3) {
          c_clon[g] /= c_area[g];
          c_clat[g] /= c_area[g];
        } else {
          double x[ORC_MV], y[ORC_MV], lo, hi, avg;
          int nv = orc_cell_poly(lon, lat, nx+1, i, j, x, y, NULL, NULL);
          orc_minmaxavg(x, nv, &lo, &hi, &avg);
          c_clon[g] = orc_poly_ctrlon(x, y, nv, avg)/cell_area[ii];
          c_clat[g] = orc_poly_ctrlat(x, y, nv)/cell_area[ii];
        }
      }
    }
    free(cell_area);
    off += (size_t)(nx+1)*(ny+1);
    offc += (size_t)nx*ny;
  }
  for (k = 0; k < n; k++) {                                                   /* :351-358 */
    size_t ii = cell_off[t_in[k]] + (size_t)j_in[k]*nx_in[t_in[k]] + i_in[k];
    di[k] -= c_clon[ii];
    dj[k] -= c_clat[ii];
  }
  free(cell_off); free(c_area); free(c_clon); free(c_clat);
}

/* ------------------------------------------------------------------------------------------
 * do_scalar_conserve_interp — conserve_interp.c:507-910, for cell_methods mean, no weights,
 * no cell_measures, no target-grid rescale (the configuration fregrid uses by default).
 * Order 2 reads data_in with a 1-cell halo: (nx+2)*(ny+2) per level per tile.
 * ---------------------------------------------------------------------------------------- */
void orc_conserve_apply(int order, long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                        const int *i_out, const int *j_out, const double *area,
                        const double *di, const double *dj,
                        int ntiles_in, const int *nx_in, const int *ny_in,
                        const double *data_in, const double *grad_x, const double *grad_y,
                        const int *grad_mask, int has_missing, double missing_in, int monotonic,
                        int nx_out, int ny_out, int nz, double *data_out)
{
  const int halo = (order == 2) ? 1 : 0;
  const size_t nout = (size_t)nx_out*ny_out;
  double missing = has_missing ? missing_in : -1.e20;                         /* :541-542 */
  double *out_area = (double *)calloc(nout*nz, sizeof(double));
  int *out_miss = (int *)calloc(nout*nz, sizeof(int));
  size_t *doff = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  size_t *goff = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  size_t *moff = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  size_t a = 0, b = 0, c = 0, q;
  long n;
  int m, k;

  for (m = 0; m < ntiles_in; m++) {
    doff[m] = a; goff[m] = b; moff[m] = c;
    a += (size_t)(nx_in[m]+2*halo)*(ny_in[m]+2*halo)*nz;
    b += (size_t)nx_in[m]*ny_in[m]*nz;
    c += (size_t)nx_in[m]*ny_in[m];
  }
  for (q = 0; q < nout*nz; q++) data_out[q] = 0.0;                            /* :556-560 */

  if (order == 1) {
    for (n = 0; n < nxgrid; n++) {                                            /* :563-614 */
      int t = t_in[n], nx1 = nx_in[t], ny1 = ny_in[t];
      for (k = 0; k < nz; k++) {
        size_t n1 = doff[t] + (size_t)k*nx1*ny1 + (size_t)j_in[n]*nx1 + i_in[n];
        size_t n0 = (size_t)k*nout + (size_t)j_out[n]*nx_out + i_out[n];
        if (has_missing && data_in[n1] == missing) continue;
        data_out[n0] += data_in[n1]*area[n];
        out_area[n0] += area[n];
        out_miss[n0] = 1;
      }
    }
  } else if (monotonic) {                                                     /* :617-742, nz == 1 */
    size_t ncell = c;
    double *f_bar_max = (double *)malloc(ncell*sizeof(double)), *f_bar_min = (double *)malloc(ncell*sizeof(double));
    double *f_max = (double *)malloc(ncell*sizeof(double)), *f_min = (double *)malloc(ncell*sizeof(double));
    double *xdata = (double *)malloc((size_t)(nxgrid > 0 ? nxgrid : 1)*sizeof(double));
    for (m = 0; m < ntiles_in; m++) {
      int nx1 = nx_in[m], ny1 = ny_in[m], i, j, ii, jj;
      for (j = 0; j < ny1; j++) for (i = 0; i < nx1; i++) {
        size_t g = moff[m] + (size_t)j*nx1 + i;
        f_bar_max[g] = -1.e20; f_bar_min[g] = 1.e20; f_max[g] = -1.e20; f_min[g] = 1.e20;
        for (jj = j-1; jj <= j+1; jj++) for (ii = i-1; ii <= i+1; ii++) {
          double v = data_in[doff[m] + (size_t)(jj+1)*(nx1+2) + ii + 1];
          if (v != missing) {
            if (v > f_bar_max[g]) f_bar_max[g] = v;
            if (v < f_bar_min[g]) f_bar_min[g] = v;
          }
        }
      }
    }
    for (n = 0; n < nxgrid; n++) {                                            /* :647-669 */
      int t = t_in[n], nx1 = nx_in[t];
      size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
      size_t n2 = doff[t] + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1;
      if (data_in[n2] != missing) {
        if (grad_mask[g]) xdata[n] = data_in[n2];
        else xdata[n] = data_in[n2] + grad_x[goff[t] + (g - moff[t])]*di[n] + grad_y[goff[t] + (g - moff[t])]*dj[n];
        if (xdata[n] > f_max[g]) f_max[g] = xdata[n];
        if (xdata[n] < f_min[g]) f_min[g] = xdata[n];
      } else xdata[n] = missing;
    }
    for (n = 0; n < nxgrid; n++) {                                            /* :680-714 */
      int t = t_in[n], nx1 = nx_in[t];
      size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
      double f_bar = data_in[doff[t] + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1];
      if (xdata[n] == missing) continue;
      if (f_max[g] > f_bar_max[g]) {
        xdata[n] = f_bar + ((xdata[n]-f_bar)/(f_max[g]-f_bar)) * (f_bar_max[g]-f_bar);
        if (xdata[n] > f_bar_max[g]) {
          if (xdata[n] - f_bar_max[g] < 1.e-10) xdata[n] = f_bar_max[g];
          if (xdata[n] > f_bar_max[g]) orc_die(" xdata is greater than f_bar_max ");
        }
      } else if (f_min[g] < f_bar_min[g]) {
        xdata[n] = f_bar + ((xdata[n]-f_bar)/(f_min[g]-f_bar)) * (f_bar_min[g]-f_bar);
        if (xdata[n] < f_bar_min[g]) {
          if (f_bar_min[g] - xdata[n] < 1.e-10) xdata[n] = f_bar_min[g];
          if (xdata[n] < f_bar_min[g]) orc_die(" xdata is less than f_bar_min ");
        }
      }
    }
    for (n = 0; n < nxgrid; n++) {                                            /* :723-740 */
      size_t n0 = (size_t)j_out[n]*nx_out + i_out[n];
      if (xdata[n] == missing) continue;
      data_out[n0] += xdata[n]*area[n];
      out_area[n0] += area[n];
    }
    free(f_bar_max); free(f_bar_min); free(f_max); free(f_min); free(xdata);
  } else {
    for (n = 0; n < nxgrid; n++) {                                            /* :745-811 */
      int t = t_in[n], nx1 = nx_in[t], ny1 = ny_in[t];
      for (k = 0; k < nz; k++) {
        size_t n0 = (size_t)k*nout + (size_t)j_out[n]*nx_out + i_out[n];
        size_t n1 = (size_t)k*nx1*ny1 + (size_t)j_in[n]*nx1 + i_in[n];
        size_t n2 = doff[t] + (size_t)k*(nx1+2)*(ny1+2) + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1;
        if (has_missing) {
          if (data_in[n2] == missing) continue;
          if (grad_mask[moff[t] + n1])
            data_out[n0] += data_in[n2]*area[n];
          else
            data_out[n0] += (data_in[n2] + grad_x[goff[t]+n1]*di[n] + grad_y[goff[t]+n1]*dj[n])*area[n];
        } else {
          data_out[n0] += (data_in[n2] + grad_x[goff[t]+n1]*di[n] + grad_y[goff[t]+n1]*dj[n])*area[n];
        }
        out_area[n0] += area[n];
        out_miss[n0] = 1;
      }
    }
  }
  for (q = 0; q < nout*nz; q++) {                                             /* :832-839 */
    if (out_area[q] > 0) data_out[q] /= out_area[q];
    else if (out_miss[q] == 1) data_out[q] = 0.0;
    else data_out[q] = missing;
  }
  free(out_area); free(out_miss); free(doff); free(goff); free(moff);
}

/* ------------------------------------------------------------------------------------------
 * do_scalar_conserve_interp with the per-source-cell factors of conserve_interp.c:572-585 / :595-603 / :731-737 /
 * :767-777 / :797-802 (weight field, cell_methods "sum", cell_measures) and the --target_grid rescale (:841-865), for one
 * field-level (nz == 1).  weight / cell_area / farea: concatenated over tiles (nx*ny each), NULL when unused.
 * ---------------------------------------------------------------------------------------- */
static double orc_entry_area(double area, size_t g, const double *weight, int cell_methods, const double *cell_area,
                             const double *farea)
{
  if (weight) area *= weight[g];
  if (cell_methods == 1) area /= cell_area[g];
  else if (farea) area *= (farea[g]/cell_area[g]);
  return area;
}

void orc_conserve_apply_ex(int order, long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                           const int *i_out, const int *j_out, const double *area,
                           const double *di, const double *dj,
                           int ntiles_in, const int *nx_in, const int *ny_in,
                           const double *data_in, const double *grad_x, const double *grad_y,
                           const int *grad_mask, int has_missing, double missing_in, int monotonic,
                           int cell_methods, const double *weight, const double *cell_area, const double *farea,
                           int target_grid, const double *dst_cell_area,
                           int nx_out, int ny_out, double *data_out)
{
  const int halo = (order == 2) ? 1 : 0;
  const size_t nout = (size_t)nx_out*ny_out;
  double missing = has_missing ? missing_in : -1.e20;
  double *out_area = (double *)calloc(nout, sizeof(double));
  int *out_miss = (int *)calloc(nout, sizeof(int));
  size_t *doff = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  size_t *moff = (size_t *)malloc((size_t)ntiles_in*sizeof(size_t));
  double *xdata = NULL, *f_bar_max = NULL, *f_bar_min = NULL, *f_max = NULL, *f_min = NULL;
  size_t a = 0, c = 0, q;
  long n;
  int m;
  for (m = 0; m < ntiles_in; m++) {
    doff[m] = a; moff[m] = c;
    a += (size_t)(nx_in[m]+2*halo)*(ny_in[m]+2*halo);
    c += (size_t)nx_in[m]*ny_in[m];
  }
  for (q = 0; q < nout; q++) data_out[q] = 0.0;
  if (order == 2 && monotonic) {
    f_bar_max = (double *)malloc(c*sizeof(double)); f_bar_min = (double *)malloc(c*sizeof(double));
    f_max = (double *)malloc(c*sizeof(double)); f_min = (double *)malloc(c*sizeof(double));
    xdata = (double *)malloc((size_t)(nxgrid > 0 ? nxgrid : 1)*sizeof(double));
    for (m = 0; m < ntiles_in; m++) {
      int nx1 = nx_in[m], ny1 = ny_in[m], i, j, ii, jj;
      for (j = 0; j < ny1; j++) for (i = 0; i < nx1; i++) {
        size_t g = moff[m] + (size_t)j*nx1 + i;
        f_bar_max[g] = -1.e20; f_bar_min[g] = 1.e20; f_max[g] = -1.e20; f_min[g] = 1.e20;
        for (jj = j-1; jj <= j+1; jj++) for (ii = i-1; ii <= i+1; ii++) {
          double v = data_in[doff[m] + (size_t)(jj+1)*(nx1+2) + ii + 1];
          if (v != missing) { if (v > f_bar_max[g]) f_bar_max[g] = v; if (v < f_bar_min[g]) f_bar_min[g] = v; }
        }
      }
    }
    for (n = 0; n < nxgrid; n++) {
      int t = t_in[n], nx1 = nx_in[t];
      size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
      double f = data_in[doff[t] + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1];
      if (f != missing) {
        xdata[n] = grad_mask[g] ? f : f + grad_x[g]*di[n] + grad_y[g]*dj[n];
        if (xdata[n] > f_max[g]) f_max[g] = xdata[n];
        if (xdata[n] < f_min[g]) f_min[g] = xdata[n];
      } else xdata[n] = missing;
    }
    for (n = 0; n < nxgrid; n++) {
      int t = t_in[n], nx1 = nx_in[t];
      size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
      double f_bar = data_in[doff[t] + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1];
      if (xdata[n] == missing) continue;
      if (f_max[g] > f_bar_max[g]) {
        xdata[n] = f_bar + ((xdata[n]-f_bar)/(f_max[g]-f_bar)) * (f_bar_max[g]-f_bar);
        if (xdata[n] > f_bar_max[g] && xdata[n] - f_bar_max[g] < 1.e-10) xdata[n] = f_bar_max[g];
      } else if (f_min[g] < f_bar_min[g]) {
        xdata[n] = f_bar + ((xdata[n]-f_bar)/(f_min[g]-f_bar)) * (f_bar_min[g]-f_bar);
        if (xdata[n] < f_bar_min[g] && f_bar_min[g] - xdata[n] < 1.e-10) xdata[n] = f_bar_min[g];
      }
    }
  }
  for (n = 0; n < nxgrid; n++) {
    int t = t_in[n], nx1 = nx_in[t];
    size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
    size_t n0 = (size_t)j_out[n]*nx_out + i_out[n];
    size_t nd = (order == 2) ? doff[t] + (size_t)(j_in[n]+1)*(nx1+2) + i_in[n] + 1 : g;
    double ar = area[n];
    if (order == 2 && monotonic) {
      if (xdata[n] == missing) continue;
      ar = orc_entry_area(ar, g, weight, cell_methods, cell_area, farea);
      data_out[n0] += xdata[n]*ar;
      out_area[n0] += ar;
      continue;
    }
    if (weight) ar *= weight[g];                                              /* before the missing test, :572, :766 */
    if (has_missing && data_in[nd] == missing) continue;
    if (cell_methods == 1) ar /= cell_area[g];
    else if (farea) ar *= (farea[g]/cell_area[g]);
    if (order == 1) data_out[n0] += data_in[nd]*ar;
    else if (has_missing && grad_mask[g]) data_out[n0] += data_in[nd]*ar;
    else data_out[n0] += (data_in[nd] + grad_x[g]*di[n] + grad_y[g]*dj[n])*ar;
    out_area[n0] += ar;
    out_miss[n0] = 1;
  }
  if (cell_methods == 1) {                                                    /* :821-830 */
    for (q = 0; q < nout; q++) if (out_area[q] == 0) data_out[q] = (out_miss[q] == 0) ? missing : 0.0;
  } else {
    for (q = 0; q < nout; q++) {                                              /* :832-839 */
      if (out_area[q] > 0) data_out[q] /= out_area[q];
      else if (out_miss[q] == 1) data_out[q] = 0.0;
      else data_out[q] = missing;
    }
    if (target_grid) {                                                        /* :841-865 */
      for (q = 0; q < nout; q++) out_area[q] = 0.0;
      for (n = 0; n < nxgrid; n++) {
        int t = t_in[n], nx1 = nx_in[t];
        size_t g = moff[t] + (size_t)j_in[n]*nx1 + i_in[n];
        size_t n0 = (size_t)j_out[n]*nx_out + i_out[n];
        if (farea) out_area[n0] += (area[n]*farea[g]/cell_area[g]);
        else out_area[n0] += area[n];
      }
      for (q = 0; q < nout; q++) if (data_out[q] != missing) data_out[q] *= (out_area[q]/dst_cell_area[q]);
    }
  }
  free(out_area); free(out_miss); free(doff); free(moff);
  free(xdata); free(f_bar_max); free(f_bar_min); free(f_max); free(f_min);
}

/* the libm this oracle (and oracle/_ref) is linked against, exposed element-wise so tests can pin
 * the product's ref_sin/ref_cos/ref_sincos (csrc/ref_trig.cuh) to it bit for bit */
void sincos(double, double *, double *);
void orc_libm_trig(long n, const double *x, double *s, double *c, double *ss, double *sc)
{
  long i;
  for (i = 0; i < n; i++) {
    s[i] = sin(x[i]);
    c[i] = cos(x[i]);
    sincos(x[i], &ss[i], &sc[i]);
  }
}
