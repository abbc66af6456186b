/* Host-side synthesis of the grids fregrid reads, so benchmarks and tests need no netCDF files.
 *
 * xgb_cubed_sphere_grid : the six tiles of make_hgrid's "gnomonic_ed" cubed sphere
 *                         (tools/make_hgrid/create_gnomonic_cubic_grid.c:101-741 with no stretch,
 *                         no nests, shift_fac = 18), subsampled and converted to radians the
 *                         way fregrid's get_input_grid does (tools/fregrid/fregrid_util.c:227-241).
 * xgb_latlon_grid       : fregrid's --nlon/--nlat regular output grid
 *                         (tools/fregrid/fregrid_util.c:588-603).
 *
 * The operation order follows the reference so that, with the same libm, the vertex
 * coordinates come out bit-identical (tests/test_grid_synth.py checks this against the
 * reference generator when it is available).  Plain C, no CUDA: this is input preparation,
 * not part of the timed path.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define GS_R2D (180/M_PI)          /* constant.h:40-41 */
#define GS_D2R (M_PI/180)
#define GS_EPS10 1.e-10
#define GS_EPS4  1.e-4
#define GS_RADIUS 6371000.

static void gs_ll2xyz(double lon, double lat, double *x, double *y, double *z)   /* mosaic_util.c:212-222 */
{
  *x = cos(lat)*cos(lon);
  *y = cos(lat)*sin(lon);
  *z = sin(lat);
}

static void gs_xyz2ll(double x, double y, double z, double *lon, double *lat)     /* mosaic_util.c:228-253 */
{
  double dist = sqrt(x*x + y*y + z*z);
  x /= dist; y /= dist; z /= dist;
  if (fabs(x) + fabs(y) < GS_EPS10) *lon = 0;
  else *lon = atan2(y, x);
  *lat = asin(z);
  if (*lon < 0.) *lon = 2.*M_PI + *lon;
}

/* mirror image of (lon0,lat0) in the plane through (lon1,lat1), (lon2,lat2) and the centre
 * (create_gnomonic_cubic_grid.c:1571-1590) */
static void gs_mirror(double lon1, double lat1, double lon2, double lat2, double lon0, double lat0,
                      double *lon, double *lat)
{
  double p0[3], p1[3], p2[3], nb[3], pp[3], pdot;
  int k;
  gs_ll2xyz(lon0, lat0, &p0[0], &p0[1], &p0[2]);
  gs_ll2xyz(lon1, lat1, &p1[0], &p1[1], &p1[2]);
  gs_ll2xyz(lon2, lat2, &p2[0], &p2[1], &p2[2]);
  nb[0] = p1[1]*p2[2] - p1[2]*p2[1];
  nb[1] = p1[2]*p2[0] - p1[0]*p2[2];
  nb[2] = p1[0]*p2[1] - p1[1]*p2[0];
  pdot = sqrt(nb[0]*nb[0] + nb[1]*nb[1] + nb[2]*nb[2]);
  for (k = 0; k < 3; k++) nb[k] = nb[k]/pdot;
  pdot = p0[0]*nb[0] + p0[1]*nb[1] + p0[2]*nb[2];
  for (k = 0; k < 3; k++) pp[k] = p0[k] - 2*pdot*nb[k];
  gs_xyz2ll(pp[0], pp[1], pp[2], lon, lat);
}

/* equal-distance gnomonic face (create_gnomonic_cubic_grid.c:1465-1538) */
static void gs_gnomonic_ed(int ni, double *lam, double *the)
{
  const int nip = ni + 1;
  const double rsq3 = 1./sqrt(3.);
  const double alpha = asin(rsq3);
  const double dely = 2.*alpha/ni;
  double *x = (double *)malloc((size_t)nip*nip*sizeof(double));
  double *y = (double *)malloc((size_t)nip*nip*sizeof(double));
  double *z = (double *)malloc((size_t)nip*nip*sizeof(double));
  int i, j;

  for (j = 0; j < nip; j++) {
    lam[j*nip]      = 0.75*M_PI;
    lam[j*nip + ni] = 1.25*M_PI;
    the[j*nip]      = -alpha + dely*j;
    the[j*nip + ni] = the[j*nip];
  }
  for (i = 1; i < ni; i++) {
    gs_mirror(lam[0], the[0], lam[ni*nip + ni], the[ni*nip + ni], lam[i*nip], the[i*nip], &lam[i], &the[i]);
    lam[ni*nip + i] = lam[i];
    the[ni*nip + i] = -the[i];
  }
  gs_ll2xyz(lam[0], the[0], &x[0], &y[0], &z[0]);
  gs_ll2xyz(lam[ni], the[ni], &x[ni], &y[ni], &z[ni]);
  gs_ll2xyz(lam[ni*nip], the[ni*nip], &x[ni*nip], &y[ni*nip], &z[ni*nip]);
  gs_ll2xyz(lam[ni*nip + ni], the[ni*nip + ni], &x[ni*nip + ni], &y[ni*nip + ni], &z[ni*nip + ni]);
  for (j = 1; j < ni; j++) {
    int n = j*nip;
    gs_ll2xyz(lam[n], the[n], &x[n], &y[n], &z[n]);
    y[n] = -y[n]*rsq3/x[n];
    z[n] = -z[n]*rsq3/x[n];
  }
  for (i = 1; i < ni; i++) {
    gs_ll2xyz(lam[i], the[i], &x[i], &y[i], &z[i]);
    y[i] = -y[i]*rsq3/x[i];
    z[i] = -z[i]*rsq3/x[i];
  }
  for (j = 0; j < nip; j++) for (i = 0; i < nip; i++) x[j*nip + i] = -rsq3;
  for (j = 1; j < nip; j++) for (i = 1; i < nip; i++) {
    y[j*nip + i] = y[i];
    z[j*nip + i] = z[j*nip];
  }
  for (j = 0; j < nip*nip; j++) gs_xyz2ll(x[j], y[j], z[j], &lam[j], &the[j]);
  free(x); free(y); free(z);
}

/* symmetrise about the face centre lines (create_gnomonic_cubic_grid.c:1596-1631) */
static void gs_symm_ed(int ni, double *lam, double *the)
{
  const int nip = ni + 1;
  int i, j;
  for (j = 1; j < nip; j++) for (i = 1; i < ni; i++) lam[j*nip + i] = lam[i];
  for (j = 0; j < nip; j++) for (i = 0; i < ni/2; i++) {
    int ip = ni - i;
    double avg = 0.5*(lam[j*nip + i] - lam[j*nip + ip]);
    lam[j*nip + i] = avg + M_PI;
    lam[j*nip + ip] = M_PI - avg;
    avg = 0.5*(the[j*nip + i] + the[j*nip + ip]);
    the[j*nip + i] = avg;
    the[j*nip + ip] = avg;
  }
  for (j = 0; j < ni/2; j++) {
    int jp = ni - j;
    for (i = 1; i < ni; i++) {
      double avg = 0.5*(lam[j*nip + i] + lam[jp*nip + i]);
      lam[j*nip + i] = avg;
      lam[jp*nip + i] = avg;
      avg = 0.5*(the[j*nip + i] - the[jp*nip + i]);
      the[j*nip + i] = avg;
      the[jp*nip + i] = -avg;
    }
  }
}

/* rotation of a spherical point about a coordinate axis, angle in degrees
 * (rot_3d with degrees=1, convert=1: create_gnomonic_cubic_grid.c:1759-1835) */
static void gs_rot(int axis, double lon, double lat, double r, double angle_deg,
                   double *lon2, double *lat2, double *r2)
{
  double x1 = r*cos(lon)*cos(lat), y1 = r*sin(lon)*cos(lat), z1 = -r*sin(lat);
  double ang = angle_deg*GS_D2R, c = cos(ang), s = sin(ang), x2, y2, z2;
  if (axis == 1)      { x2 = x1;           y2 = c*y1 + s*z1;  z2 = -s*y1 + c*z1; }
  else if (axis == 2) { x2 = c*x1 - s*z1;  y2 = y1;           z2 = s*x1 + c*z1; }
  else                { x2 = c*x1 + s*y1;  y2 = -s*x1 + c*y1; z2 = z1; }
  *r2 = sqrt(x2*x2 + y2*y2 + z2*z2);
  if ((fabs(x2) + fabs(y2)) < GS_EPS10) *lon2 = 0.;
  else *lon2 = atan2(y2, x2);
  *lat2 = acos(z2/(*r2)) - M_PI/2.;
}

/* four-fold symmetrisation of tile 1 and generation of tiles 2..6
 * (create_gnomonic_cubic_grid.c:1637-1751) */
static void gs_mirror_grid(int ni, double *x, double *y)
{
  const int nip = ni + 1;
  const int half = (int)ceil(nip/2.);
  int i, j, nt;
  for (j = 0; j < half; j++) {
    int jp = ni - j;
    for (i = 0; i < half; i++) {
      int ip = ni - i;
      double x1 = 0.25*(fabs(x[j*nip + i]) + fabs(x[j*nip + ip]) + fabs(x[jp*nip + i]) + fabs(x[jp*nip + ip]));
      double y1;
      x[j*nip + i]   = x1*(x[j*nip + i]   >= 0 ? 1 : -1);
      x[j*nip + ip]  = x1*(x[j*nip + ip]  >= 0 ? 1 : -1);
      x[jp*nip + i]  = x1*(x[jp*nip + i]  >= 0 ? 1 : -1);
      x[jp*nip + ip] = x1*(x[jp*nip + ip] >= 0 ? 1 : -1);
      y1 = 0.25*(fabs(y[j*nip + i]) + fabs(y[j*nip + ip]) + fabs(y[jp*nip + i]) + fabs(y[jp*nip + ip]));
      y[j*nip + i]   = y1*(y[j*nip + i]   >= 0 ? 1 : -1);
      y[j*nip + ip]  = y1*(y[j*nip + ip]  >= 0 ? 1 : -1);
      y[jp*nip + i]  = y1*(y[jp*nip + i]  >= 0 ? 1 : -1);
      y[jp*nip + ip] = y1*(y[jp*nip + ip] >= 0 ? 1 : -1);
      if (nip%2) {
        if (i == (nip - 1)/2) { x[j*nip + i] = 0.0; x[jp*nip + i] = 0.0; }
      }
    }
  }
  for (nt = 1; nt < 6; nt++) {
    for (j = 0; j < nip; j++) for (i = 0; i < nip; i++) {
      double x1 = x[j*nip + i], y1 = y[j*nip + i], z1 = GS_RADIUS, x2, y2, z2;
      const int mid = (nip - 1)/2;
      switch (nt) {
      case 1:
        gs_rot(3, x1, y1, z1, -90., &x2, &y2, &z2);
        break;
      case 2:
        gs_rot(3, x1, y1, z1, -90., &x2, &y2, &z2);
        gs_rot(1, x2, y2, z2, 90., &x1, &y1, &z1);
        x2 = x1; y2 = y1; z2 = z1;
        if (nip%2) {
          if (i == mid && i == j) { x2 = 0; y2 = M_PI*0.5; }
          if (j == mid && i < mid) x2 = 0;
          if (j == mid && i > mid) x2 = M_PI;
        }
        break;
      case 3:
        gs_rot(3, x1, y1, z1, -180., &x2, &y2, &z2);
        gs_rot(1, x2, y2, z2, 90., &x1, &y1, &z1);
        x2 = x1; y2 = y1; z2 = z1;
        if (nip%2) { if (j == mid) x2 = M_PI; }
        break;
      case 4:
        gs_rot(3, x1, y1, z1, 90., &x2, &y2, &z2);
        gs_rot(2, x2, y2, z2, 90., &x1, &y1, &z1);
        x2 = x1; y2 = y1; z2 = z1;
        break;
      default:
        gs_rot(2, x1, y1, z1, 90., &x2, &y2, &z2);
        gs_rot(3, x2, y2, z2, 0., &x1, &y1, &z1);
        x2 = x1; y2 = y1; z2 = z1;
        if (nip%2) {
          if (i == mid && i == j) { x2 = 0; y2 = -M_PI*0.5; }
          if (i == mid && j > mid) x2 = 0;
          if (i == mid && j < mid) x2 = M_PI;
        }
        break;
      }
      x[(size_t)nt*nip*nip + j*nip + i] = x2;
      y[(size_t)nt*nip*nip + j*nip + i] = y2;
    }
  }
}

/* cell centres: normalised sum of the four corner vectors (create_gnomonic_cubic_grid.c:2008-2048) */
static void gs_cell_center(int ni, const double *lonc, const double *latc, double *lont, double *latt)
{
  const int nip = ni + 1;
  double *xc = (double *)malloc((size_t)nip*nip*sizeof(double));
  double *yc = (double *)malloc((size_t)nip*nip*sizeof(double));
  double *zc = (double *)malloc((size_t)nip*nip*sizeof(double));
  int i, j;
  for (j = 0; j < nip*nip; j++) gs_ll2xyz(lonc[j], latc[j], &xc[j], &yc[j], &zc[j]);
  for (j = 0; j < ni; j++) for (i = 0; i < ni; i++) {
    int p1 = j*nip + i, p2 = p1 + 1, p3 = (j + 1)*nip + i + 1, p4 = (j + 1)*nip + i;
    double xt = xc[p1] + xc[p2] + xc[p3] + xc[p4];
    double yt = yc[p1] + yc[p2] + yc[p3] + yc[p4];
    double zt = zc[p1] + zc[p2] + zc[p3] + zc[p4];
    double dd = sqrt(pow(xt, 2) + pow(yt, 2) + pow(zt, 2));
    xt /= dd; yt /= dd; zt /= dd;
    gs_xyz2ll(xt, yt, zt, &lont[j*ni + i], &latt[j*ni + i]);
  }
  free(xc); free(yc); free(zc);
}

/* lonc/latc: 6*(ni+1)*(ni+1) corners in radians; lont/latt: 6*ni*ni centres (may be NULL).
 * Returns 0 on success. */
int xgb_cubed_sphere_grid(int ni, double *lonc, double *latc, double *lont, double *latt)
{
  const int nip = ni + 1;
  const size_t np = (size_t)nip*nip;
  double *lam, *the, *xc, *yc;
  size_t n;
  int i, j, t;
  if (ni <= 0 || ni%1) return 1;
  lam = (double *)malloc(np*sizeof(double));
  the = (double *)malloc(np*sizeof(double));
  xc = (double *)malloc(6*np*sizeof(double));
  yc = (double *)malloc(6*np*sizeof(double));
  if (!lam || !the || !xc || !yc) return 1;

  gs_gnomonic_ed(ni, lam, the);
  gs_symm_ed(ni, lam, the);
  for (n = 0; n < np; n++) { xc[n] = lam[n] - M_PI; yc[n] = the[n]; }     /* :331-336 */
  gs_mirror_grid(ni, xc, yc);
  for (n = 0; n < 6*np; n++) {                                             /* :343-349, shift_fac = 18 */
    xc[n] -= M_PI/18.;
    if (xc[n] < 0.) xc[n] += 2.*M_PI;
    if (fabs(xc[n]) < GS_EPS10) xc[n] = 0;
    if (fabs(yc[n]) < GS_EPS10) yc[n] = 0;
  }
  /* shared edges take one owner's values (:352-385) */
  for (j = 0; j < nip; j++) {
    xc[np + j*nip]   = xc[j*nip + ni];            yc[np + j*nip]   = yc[j*nip + ni];
    xc[2*np + j*nip] = xc[ni*nip + ni - j];       yc[2*np + j*nip] = yc[ni*nip + ni - j];
  }
  for (i = 0; i < nip; i++) {
    xc[4*np + ni*nip + i] = xc[(ni - i)*nip];             yc[4*np + ni*nip + i] = yc[(ni - i)*nip];
    xc[5*np + ni*nip + i] = xc[i];                        yc[5*np + ni*nip + i] = yc[i];
    xc[2*np + i]          = xc[np + ni*nip + i];          yc[2*np + i]          = yc[np + ni*nip + i];
    xc[3*np + i]          = xc[np + (ni - i)*nip + ni];   yc[3*np + i]          = yc[np + (ni - i)*nip + ni];
  }
  for (j = 0; j < nip; j++) {
    xc[5*np + j*nip + ni] = xc[np + ni - j];              yc[5*np + j*nip + ni] = yc[np + ni - j];
    xc[3*np + j*nip]      = xc[2*np + j*nip + ni];        yc[3*np + j*nip]      = yc[2*np + j*nip + ni];
    xc[4*np + j*nip]      = xc[2*np + ni*nip + ni - j];   yc[4*np + j*nip]      = yc[2*np + ni*nip + ni - j];
  }
  for (i = 0; i < nip; i++) {
    xc[4*np + i] = xc[3*np + ni*nip + i];                 yc[4*np + i] = yc[3*np + ni*nip + i];
    xc[5*np + i] = xc[3*np + (ni - i)*nip + ni];          yc[5*np + i] = yc[3*np + (ni - i)*nip + ni];
  }
  for (j = 0; j < nip; j++) {
    xc[5*np + j*nip] = xc[4*np + j*nip + ni];             yc[5*np + j*nip] = yc[4*np + j*nip + ni];
  }

  if (lont && latt) {
    double *ct = (double *)malloc((size_t)ni*ni*sizeof(double)), *cy = (double *)malloc((size_t)ni*ni*sizeof(double));
    for (t = 0; t < 6; t++) {
      gs_cell_center(ni, xc + t*np, yc + t*np, ct, cy);
      for (n = 0; n < (size_t)ni*ni; n++) {                /* degrees on disk, radians in fregrid */
        lont[(size_t)t*ni*ni + n] = (ct[n]*GS_R2D)*GS_D2R;
        latt[(size_t)t*ni*ni + n] = (cy[n]*GS_R2D)*GS_D2R;
      }
    }
    free(ct); free(cy);
  }
  for (n = 0; n < 6*np; n++) {                             /* :717-720 then fregrid_util.c:230-231 */
    lonc[n] = (xc[n]*GS_R2D)*GS_D2R;
    latc[n] = (yc[n]*GS_R2D)*GS_D2R;
  }
  free(lam); free(the); free(xc); free(yc);
  return 0;
}

/* regular lat-lon destination grid of fregrid --nlon/--nlat with the default bounds
 * lonbegin=0, lonend=360, latbegin=-90, latend=90 (fregrid_util.c:588-603, fregrid.c:296-299).
 * lonc/latc: (nlon+1)*(nlat+1) vertices in radians. */
int xgb_latlon_grid(int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend,
                    double *lonc, double *latc)
{
  int i, j;
  double dlon, dlat;
  if (nlon <= 0 || nlat <= 0) return 1;
  dlon = (lonend - lonbegin)/nlon;
  dlat = (latend - latbegin)/nlat;
  for (j = 0; j <= nlat; j++) for (i = 0; i <= nlon; i++) {
    lonc[(size_t)j*(nlon + 1) + i] = (lonbegin + i*dlon)*GS_D2R;
    latc[(size_t)j*(nlon + 1) + i] = (latbegin + j*dlat)*GS_D2R;
  }
  return 0;
}
