/* fregrid remap files in classic netCDF: the writer of setup_conserve_interp (reference tools/fregrid/
 * conserve_interp.c:382-438) and the readers read_mosaic_xgrid_size / _order1 / _order2 (tools/libfrencutils/
 * read_mosaic.c:330-560) on top of nc3.c.
 *
 * Layout written (and expected), in this order:
 *   dimensions  string = 255, ncells = nxgrid, two = 2
 *   int    tile1(ncells)              standard_name = "tile_number_in_mosaic1"              1-based tile
 *   int    tile1_cell(ncells, two)    standard_name = "parent_cell_indices_in_mosaic1"      (i, j) 1-based
 *   int    tile2_cell(ncells, two)    standard_name = "parent_cell_indices_in_mosaic2"      (i + isc, j + jsc) 1-based
 *   double xgrid_area(ncells)         standard_name = "exchange_grid_area", units = "m2"
 *   double tile1_distance(ncells,two) standard_name = "distance_from_parent1_cell_centroid"  order 2 only
 * no global attributes.  The in-memory lists are 0-based (conserve_interp.c:405-423).
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/xgrid_b200.h"
#include "nc3.h"

void xgb_set_error_str(const char *s);       /* xgrid_capi.cu */
static void xgb_set_error(const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  xgb_set_error_str(buf);
}

#define REMAP_STRING 255            /* constant.h:25 */
#define REMAP_RADIUS 6371000.0      /* constant.h:23 */

static int g_nc_format = 2;         /* 64-bit offset; the reference's own default (netCDF-4 classic model) is HDF5 */

int xgb_set_nc_format(const char *name)
{
  /* set_in_format, mpp_io.c:1526-1540 */
  if (!name) return 0;
  if (!strcmp(name, "classic")) g_nc_format = 1;
  else if (!strcmp(name, "64bit_offset")) g_nc_format = 2;
  else if (!strcmp(name, "cdf5") || !strcmp(name, "64bit_data")) g_nc_format = 5;
  else if (!strcmp(name, "netcdf4") || !strcmp(name, "netcdf4_classic")) {
    xgb_set_error("format = %s needs HDF5, which this library does not write; use classic, 64bit_offset or cdf5", name);
    return 1;
  } else {
    xgb_set_error("mpp_io(mpp_open): format = %s is not a valid option", name);
    return 1;
  }
  return 0;
}

int xgb_get_nc_format(void) { return g_nc_format; }

static int *interleave_int(long long n, const int *a, int abase, const int *b, int bbase)
{
  int *out = (int *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(int));
  long long i;
  if (!out) return NULL;
  for (i = 0; i < n; ++i) { out[2 * i] = a[i] + abase; out[2 * i + 1] = b[i] + bbase; }
  return out;
}

int xgb_remap_write(const char *path, int order, long long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                    const int *i_out, const int *j_out, int isc, int jsc, const double *area, const double *di, const double *dj)
{
  char err[320];
  nc3_file *f;
  int d_str, d_n, d_two, dims[2], v_t1, v_c1, v_c2, v_area, v_dist = -1, rc = 1;
  int *ibuf = NULL;
  double *dbuf = NULL;
  long long i;
  if (!path || nxgrid <= 0 || !t_in || !i_in || !j_in || !i_out || !j_out || !area || (order == 2 && (!di || !dj))) {
    xgb_set_error("xgb_remap_write: bad arguments (the reference writes no file for an empty exchange grid)");
    return 1;
  }
  f = nc3_create(path, g_nc_format, err, sizeof err);
  if (!f) { xgb_set_error("%s", err); return 1; }
  d_str = nc3_def_dim(f, "string", REMAP_STRING);
  d_n = nc3_def_dim(f, "ncells", nxgrid);
  d_two = nc3_def_dim(f, "two", 2);
  (void)d_str;
  dims[0] = d_n; dims[1] = d_two;
  v_t1 = nc3_def_var(f, "tile1", NC3_INT, 1, &d_n);
  nc3_put_att_text(f, v_t1, "standard_name", "tile_number_in_mosaic1");
  v_c1 = nc3_def_var(f, "tile1_cell", NC3_INT, 2, dims);
  nc3_put_att_text(f, v_c1, "standard_name", "parent_cell_indices_in_mosaic1");
  v_c2 = nc3_def_var(f, "tile2_cell", NC3_INT, 2, dims);
  nc3_put_att_text(f, v_c2, "standard_name", "parent_cell_indices_in_mosaic2");
  v_area = nc3_def_var(f, "xgrid_area", NC3_DOUBLE, 1, &d_n);
  nc3_put_att_text(f, v_area, "standard_name", "exchange_grid_area");
  nc3_put_att_text(f, v_area, "units", "m2");
  if (order == 2) {
    v_dist = nc3_def_var(f, "tile1_distance", NC3_DOUBLE, 2, dims);
    nc3_put_att_text(f, v_dist, "standard_name", "distance_from_parent1_cell_centroid");
  }
  if (v_t1 < 0 || v_c1 < 0 || v_c2 < 0 || v_area < 0 || (order == 2 && v_dist < 0) || nc3_enddef(f)) goto io_error;

  ibuf = (int *)malloc((size_t)nxgrid * sizeof(int));
  if (!ibuf) { xgb_set_error("xgb_remap_write: out of memory"); goto done; }
  for (i = 0; i < nxgrid; ++i) ibuf[i] = t_in[i] + 1;
  if (nc3_put_var_int(f, v_t1, ibuf)) goto io_error;
  free(ibuf);
  ibuf = interleave_int(nxgrid, i_in, 1, j_in, 1);
  if (!ibuf) { xgb_set_error("xgb_remap_write: out of memory"); goto done; }
  if (nc3_put_var_int(f, v_c1, ibuf)) goto io_error;
  free(ibuf);
  ibuf = interleave_int(nxgrid, i_out, isc + 1, j_out, jsc + 1);
  if (!ibuf) { xgb_set_error("xgb_remap_write: out of memory"); goto done; }
  if (nc3_put_var_int(f, v_c2, ibuf)) goto io_error;
  if (nc3_put_var_double(f, v_area, area)) goto io_error;
  if (order == 2) {
    dbuf = (double *)malloc((size_t)nxgrid * 2 * sizeof(double));
    if (!dbuf) { xgb_set_error("xgb_remap_write: out of memory"); goto done; }
    for (i = 0; i < nxgrid; ++i) { dbuf[2 * i] = di[i]; dbuf[2 * i + 1] = dj[i]; }
    if (nc3_put_var_double(f, v_dist, dbuf)) goto io_error;
  }
  rc = 0;
  goto done;
io_error:
  xgb_set_error("xgb_remap_write(%s): %s", path, nc3_strerror(f));
done:
  free(ibuf);
  free(dbuf);
  if (nc3_close(f) && rc == 0) { xgb_set_error("xgb_remap_write(%s): close failed", path); rc = 1; }
  return rc;
}

long long xgb_remap_size(const char *path)
{
  /* read_mosaic_xgrid_size, read_mosaic.c:330-337: the length of dimension "ncells" */
  char err[320];
  nc3_file *f = nc3_open(path, err, sizeof err);
  long long n;
  int d;
  if (!f) { xgb_set_error("%s", err); return -1; }
  d = nc3_dim_id(f, "ncells");
  n = (d < 0) ? -1 : nc3_dim_len(f, d);
  if (d < 0) xgb_set_error("%s has no dimension ncells", path);
  nc3_close(f);
  return n;
}

int xgb_remap_read(const char *path, int order, long long cap, int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                   double *area, double *di, double *dj)
{
  char err[320];
  nc3_file *f = nc3_open(path, err, sizeof err);
  long long n, i;
  int d, v, rc = 1;
  int *ibuf = NULL;
  double *dbuf = NULL;
  const double garea = 4 * M_PI * REMAP_RADIUS * REMAP_RADIUS;
  if (!f) { xgb_set_error("%s", err); return 1; }
  d = nc3_dim_id(f, "ncells");
  if (d < 0) { xgb_set_error("%s has no dimension ncells", path); goto done; }
  n = nc3_dim_len(f, d);
  if (n > cap) { xgb_set_error("xgb_remap_read: %s holds %lld cells, the buffers %lld", path, n, cap); goto done; }
  ibuf = (int *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(int));
  if (!ibuf) { xgb_set_error("xgb_remap_read: out of memory"); goto done; }
  /* every variable is checked for the shape the buffers were sized for — (ncells) or (ncells, two = 2) — before it is read:
     a file with another layout is refused instead of overrunning them */
#define NEED_VAR(name, pairs) do { \
    v = nc3_var_id(f, name); \
    if (v < 0) { xgb_set_error("%s has no variable %s", path, name); goto done; } \
    { const int nd_ = nc3_var_ndims(f, v); const int *di_ = nc3_var_dimids(f, v); \
      if (nd_ != ((pairs) ? 2 : 1) || di_[0] != d || ((pairs) && nc3_dim_len(f, di_[1]) != 2) || nc3_var_type(f, v) == NC3_CHAR) { \
        xgb_set_error("%s: variable %s is not a numeric (%s) array", path, name, (pairs) ? "ncells, 2" : "ncells"); goto done; } } \
  } while (0)
  NEED_VAR("tile1_cell", 1);
  if (nc3_get_var_int(f, v, ibuf)) goto io_error;
  for (i = 0; i < n; ++i) { i_in[i] = ibuf[2 * i] - 1; j_in[i] = ibuf[2 * i + 1] - 1; }
  NEED_VAR("tile2_cell", 1);
  if (nc3_get_var_int(f, v, ibuf)) goto io_error;
  for (i = 0; i < n; ++i) { i_out[i] = ibuf[2 * i] - 1; j_out[i] = ibuf[2 * i + 1] - 1; }
  NEED_VAR("xgrid_area", 0);
  if (nc3_get_var_double(f, v, area)) goto io_error;
  /* the reference scales to unit-sphere area on read (read_mosaic.c:437) and back in setup_conserve_interp
     (conserve_interp.c:86); the two roundings are kept so that READ reproduces its numbers */
  for (i = 0; i < n; ++i) { area[i] /= garea; area[i] *= garea; }
  if (order == 2) {
    NEED_VAR("tile1_distance", 1);
    dbuf = (double *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(double));
    if (!dbuf) { xgb_set_error("xgb_remap_read: out of memory"); goto done; }
    if (nc3_get_var_double(f, v, dbuf)) goto io_error;
    for (i = 0; i < n; ++i) { di[i] = dbuf[2 * i]; dj[i] = dbuf[2 * i + 1]; }
  }
  NEED_VAR("tile1", 0);
  if (nc3_get_var_int(f, v, t_in)) goto io_error;
  for (i = 0; i < n; ++i) t_in[i] -= 1;                   /* conserve_interp.c:110 */
  rc = 0;
  goto done;
io_error:
  xgb_set_error("xgb_remap_read(%s): %s", path, nc3_strerror(f));
done:
  free(ibuf);
  free(dbuf);
  nc3_close(f);
  return rc;
}
