/* TEST INFRASTRUCTURE — the subset of the netCDF-C API the reference's I/O layer and tool mains call (oracle/shim/netcdf.h),
 * implemented over the classic-format reader / writer fre-nctools_b200/csrc/nc3.c, so that the UNMODIFIED reference
 * fregrid / make_coupler_mosaic compile and run in this image (no libnetcdf, no HDF5) on classic netCDF files.
 * Numeric variables are moved as doubles / ints and converted; text as bytes.  Not a general netCDF library. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "netcdf.h"
#include "nc3.h"

#define MAXF 256
static nc3_file *g_f[MAXF];
static char g_msg[512] = "netCDF shim error";

const char *nc_strerror(int status) { (void)status; return g_msg; }

static nc3_file *F(int ncid) { return (ncid >= 0 && ncid < MAXF) ? g_f[ncid] : NULL; }
static int fail(const char *what, nc3_file *f, int code)
{
  snprintf(g_msg, sizeof g_msg, "%s: %s", what, f ? nc3_strerror(f) : "bad ncid");
  return code;
}
static int slot(void) { int i; for (i = 0; i < MAXF; i++) if (!g_f[i]) return i; return -1; }

int nc_open(const char *path, int mode, int *ncidp)
{
  char err[400];
  int s = slot();
  (void)mode;
  if (s < 0) { snprintf(g_msg, sizeof g_msg, "too many open files"); return NC_EINVAL; }
  g_f[s] = nc3_open(path, err, sizeof err);
  if (!g_f[s]) { snprintf(g_msg, sizeof g_msg, "%s", err); return NC_EIO; }
  *ncidp = s;
  return NC_NOERR;
}

int nc__create(const char *path, int cmode, size_t initialsz, size_t *chunksizehintp, int *ncidp)
{
  char err[400];
  int s = slot();
  /* no HDF5 here: NC_NETCDF4 (with or without NC_CLASSIC_MODEL) becomes 64-bit offset; NC_CLASSIC_MODEL alone is CDF-1 */
  const int fmt = (cmode & (NC_64BIT_OFFSET | NC_NETCDF4)) ? 2 : 1;
  (void)initialsz; (void)chunksizehintp;
  if (s < 0) { snprintf(g_msg, sizeof g_msg, "too many open files"); return NC_EINVAL; }
  g_f[s] = nc3_create(path, fmt, err, sizeof err);
  if (!g_f[s]) { snprintf(g_msg, sizeof g_msg, "%s", err); return NC_EIO; }
  *ncidp = s;
  return NC_NOERR;
}
int nc_create(const char *path, int cmode, int *ncidp) { return nc__create(path, cmode, 0, NULL, ncidp); }

int nc_close(int ncid)
{
  nc3_file *f = F(ncid);
  if (!f) return fail("nc_close", NULL, NC_EBADID);
  g_f[ncid] = NULL;
  return nc3_close(f) ? NC_EIO : NC_NOERR;
}
int nc_sync(int ncid) { return F(ncid) ? NC_NOERR : NC_EBADID; }
int nc_redef(int ncid) { (void)ncid; snprintf(g_msg, sizeof g_msg, "nc_redef is not supported by the classic-format shim"); return NC_EINVAL; }
int nc_enddef(int ncid) { nc3_file *f = F(ncid); if (!f) return NC_EBADID; return nc3_enddef(f) ? fail("nc_enddef", f, NC_EIO) : NC_NOERR; }
int nc__enddef(int ncid, size_t a, size_t b, size_t c, size_t d) { (void)a; (void)b; (void)c; (void)d; return nc_enddef(ncid); }

int nc_inq_format(int ncid, int *formatp)
{
  nc3_file *f = F(ncid);
  if (!f) return NC_EBADID;
  { const int k = nc3_format(f); *formatp = (k == 1) ? NC_FORMAT_CLASSIC : (k == 2) ? NC_FORMAT_64BIT : NC_FORMAT_64BIT_DATA; }
  return NC_NOERR;
}
int nc_inq_nvars(int ncid, int *n) { nc3_file *f = F(ncid); if (!f) return NC_EBADID; *n = nc3_nvars(f); return NC_NOERR; }
int nc_inq_unlimdim(int ncid, int *id) { nc3_file *f = F(ncid); if (!f) return NC_EBADID; *id = nc3_unlimdim(f); return NC_NOERR; }
int nc_inq_dimid(int ncid, const char *name, int *idp)
{
  nc3_file *f = F(ncid); int d;
  if (!f) return NC_EBADID;
  d = nc3_dim_id(f, name);
  if (d < 0) { snprintf(g_msg, sizeof g_msg, "dimension %s not found", name); return NC_EBADDIM; }
  *idp = d;
  return NC_NOERR;
}
int nc_inq_dimlen(int ncid, int dimid, size_t *lenp)
{
  nc3_file *f = F(ncid);
  if (!f || dimid < 0 || dimid >= nc3_ndims(f)) return NC_EBADDIM;
  *lenp = (size_t)nc3_dim_len(f, dimid);
  return NC_NOERR;
}
int nc_inq_dimname(int ncid, int dimid, char *name)
{
  nc3_file *f = F(ncid);
  if (!f || dimid < 0 || dimid >= nc3_ndims(f)) return NC_EBADDIM;
  strcpy(name, nc3_dim_name(f, dimid));
  return NC_NOERR;
}
int nc_inq_varid(int ncid, const char *name, int *varidp)
{
  nc3_file *f = F(ncid); int v;
  if (!f) return NC_EBADID;
  v = nc3_var_id(f, name);
  if (v < 0) { snprintf(g_msg, sizeof g_msg, "variable %s not found", name); return NC_ENOTVAR; }
  *varidp = v;
  return NC_NOERR;
}
static int vok(nc3_file *f, int v) { return f && v >= 0 && v < nc3_nvars(f); }
int nc_inq_varname(int ncid, int v, char *name) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; strcpy(name, nc3_var_name(f, v)); return NC_NOERR; }
int nc_inq_vartype(int ncid, int v, nc_type *t) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; *t = nc3_var_type(f, v); return NC_NOERR; }
int nc_inq_varndims(int ncid, int v, int *n) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; *n = nc3_var_ndims(f, v); return NC_NOERR; }
int nc_inq_vardimid(int ncid, int v, int *ids)
{
  nc3_file *f = F(ncid); int k;
  if (!vok(f, v)) return NC_ENOTVAR;
  for (k = 0; k < nc3_var_ndims(f, v); k++) ids[k] = nc3_var_dimids(f, v)[k];
  return NC_NOERR;
}
int nc_inq_varnatts(int ncid, int v, int *n)
{
  nc3_file *f = F(ncid);
  if (!f || (v != NC_GLOBAL && !vok(f, v))) return NC_ENOTVAR;
  *n = nc3_var_natts(f, v == NC_GLOBAL ? NC3_GLOBAL : v);
  return NC_NOERR;
}
static int VA(int v) { return v == NC_GLOBAL ? NC3_GLOBAL : v; }
int nc_inq_att(int ncid, int v, const char *name, nc_type *tp, size_t *lenp)
{
  nc3_file *f = F(ncid); int t; long long n;
  if (!f) return NC_EBADID;
  if (nc3_att_inq(f, VA(v), name, &t, &n) < 0 || n < 0) { snprintf(g_msg, sizeof g_msg, "attribute %s not found", name); return NC_ENOTATT; }
  if (tp) *tp = t;
  if (lenp) *lenp = (size_t)n;
  return NC_NOERR;
}
int nc_inq_atttype(int ncid, int v, const char *name, nc_type *tp) { return nc_inq_att(ncid, v, name, tp, NULL); }
int nc_inq_attlen(int ncid, int v, const char *name, size_t *lenp) { return nc_inq_att(ncid, v, name, NULL, lenp); }
int nc_inq_attname(int ncid, int v, int attnum, char *name)
{
  nc3_file *f = F(ncid); const char *s;
  if (!f) return NC_EBADID;
  s = nc3_att_name(f, VA(v), attnum);
  if (!s) return NC_ENOTATT;
  strcpy(name, s);
  return NC_NOERR;
}
int nc_get_att_text(int ncid, int v, const char *name, char *ip)
{
  nc3_file *f = F(ncid); size_t n; nc_type t; char *tmp; int st;
  if ((st = nc_inq_att(ncid, v, name, &t, &n)) != NC_NOERR) return st;
  tmp = (char *)malloc(n + 2);
  if (nc3_get_att_text(f, VA(v), name, tmp, n + 1)) { free(tmp); return fail("nc_get_att_text", f, NC_ENOTATT); }
  memcpy(ip, tmp, n);                                  /* netCDF does not terminate attribute text */
  free(tmp);
  return NC_NOERR;
}
int nc_get_att_double(int ncid, int v, const char *name, double *ip)
{
  nc3_file *f = F(ncid); size_t n; nc_type t; int st;
  if ((st = nc_inq_att(ncid, v, name, &t, &n)) != NC_NOERR) return st;
  return nc3_get_att_double(f, VA(v), name, ip, (int)n) < 0 ? fail("nc_get_att_double", f, NC_ENOTATT) : NC_NOERR;
}
int nc_get_att_int(int ncid, int v, const char *name, int *ip)
{
  size_t n, k; nc_type t; double *d; int st;
  if ((st = nc_inq_att(ncid, v, name, &t, &n)) != NC_NOERR) return st;
  d = (double *)malloc((n + 1) * sizeof(double));
  st = nc_get_att_double(ncid, v, name, d);
  for (k = 0; st == NC_NOERR && k < n; k++) ip[k] = (int)d[k];
  free(d);
  return st;
}
int nc_get_att_short(int ncid, int v, const char *name, short *ip)
{
  size_t n, k; nc_type t; double *d; int st;
  if ((st = nc_inq_att(ncid, v, name, &t, &n)) != NC_NOERR) return st;
  d = (double *)malloc((n + 1) * sizeof(double));
  st = nc_get_att_double(ncid, v, name, d);
  for (k = 0; st == NC_NOERR && k < n; k++) ip[k] = (short)d[k];
  free(d);
  return st;
}
int nc_put_att_text(int ncid, int v, const char *name, size_t len, const char *op)
{
  nc3_file *f = F(ncid); char *tmp; int st;
  if (!f) return NC_EBADID;
  tmp = (char *)malloc(len + 1);
  memcpy(tmp, op, len); tmp[len] = 0;
  while (len > 0 && tmp[len - 1] == 0) --len;          /* nc3 stores strlen bytes */
  st = nc3_put_att_text(f, VA(v), name, tmp);
  free(tmp);
  return st ? fail("nc_put_att_text", f, NC_EINVAL) : NC_NOERR;
}
int nc_put_att_double(int ncid, int v, const char *name, nc_type xtype, size_t len, const double *op)
{
  nc3_file *f = F(ncid);
  if (!f) return NC_EBADID;
  return nc3_put_att_double(f, VA(v), name, xtype, (int)len, op) ? fail("nc_put_att_double", f, NC_EINVAL) : NC_NOERR;
}
int nc_copy_att(int ncid_in, int v_in, const char *name, int ncid_out, int v_out)
{
  nc3_file *a = F(ncid_in), *b = F(ncid_out);
  if (!a || !b) return NC_EBADID;
  return nc3_copy_att(a, VA(v_in), name, b, VA(v_out)) ? fail("nc_copy_att", b, NC_ENOTATT) : NC_NOERR;
}
int nc_def_dim(int ncid, const char *name, size_t len, int *idp)
{
  nc3_file *f = F(ncid); int d;
  if (!f) return NC_EBADID;
  d = nc3_def_dim(f, name, (long long)len);
  if (d < 0) return fail("nc_def_dim", f, NC_EINVAL);
  if (idp) *idp = d;
  return NC_NOERR;
}
int nc_def_var(int ncid, const char *name, nc_type xtype, int ndims, const int *dimidsp, int *varidp)
{
  nc3_file *f = F(ncid); int v;
  if (!f) return NC_EBADID;
  v = nc3_def_var(f, name, xtype, ndims, dimidsp);
  if (v < 0) return fail("nc_def_var", f, NC_EINVAL);
  if (varidp) *varidp = v;
  return NC_NOERR;
}
int nc_def_var_deflate(int ncid, int v, int s, int d, int l) { (void)ncid; (void)v; (void)s; (void)d; (void)l; return NC_NOERR; }
int nc_inq_var_deflate(int ncid, int v, int *s, int *d, int *l) { (void)ncid; (void)v; if (s) *s = 0; if (d) *d = 0; if (l) *l = 0; return NC_NOERR; }

/* element count of a hyperslab, or of the whole variable when count == NULL */
static size_t nelem(nc3_file *f, int v, const size_t *count)
{
  size_t n = 1; int k;
  for (k = 0; k < nc3_var_ndims(f, v); k++)
    n *= count ? count[k] : (size_t)nc3_dim_len(f, nc3_var_dimids(f, v)[k]);
  return n;
}
static void whole(nc3_file *f, int v, size_t *start, size_t *count)
{
  int k;
  for (k = 0; k < nc3_var_ndims(f, v); k++) { start[k] = 0; count[k] = (size_t)nc3_dim_len(f, nc3_var_dimids(f, v)[k]); }
}

int nc_get_vara_text(int ncid, int v, const size_t *s, const size_t *c, char *ip) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_get_vara_text(f, v, s, c, ip) ? fail("nc_get_vara_text", f, NC_EIO) : NC_NOERR; }
int nc_get_vara_int(int ncid, int v, const size_t *s, const size_t *c, int *ip) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_get_vara_int(f, v, s, c, ip) ? fail("nc_get_vara_int", f, NC_EIO) : NC_NOERR; }
int nc_get_vara_double(int ncid, int v, const size_t *s, const size_t *c, double *ip) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_get_vara_double(f, v, s, c, ip) ? fail("nc_get_vara_double", f, NC_EIO) : NC_NOERR; }
int nc_get_vara_short(int ncid, int v, const size_t *s, const size_t *c, short *ip)
{
  nc3_file *f = F(ncid); size_t n, k; int *t; int st;
  if (!vok(f, v)) return NC_ENOTVAR;
  n = nelem(f, v, c); t = (int *)malloc((n + 1) * sizeof(int));
  st = nc3_get_vara_int(f, v, s, c, t);
  for (k = 0; !st && k < n; k++) ip[k] = (short)t[k];
  free(t);
  return st ? fail("nc_get_vara_short", f, NC_EIO) : NC_NOERR;
}
int nc_get_vara_float(int ncid, int v, const size_t *s, const size_t *c, float *ip)
{
  nc3_file *f = F(ncid); size_t n, k; double *t; int st;
  if (!vok(f, v)) return NC_ENOTVAR;
  n = nelem(f, v, c); t = (double *)malloc((n + 1) * sizeof(double));
  st = nc3_get_vara_double(f, v, s, c, t);
  for (k = 0; !st && k < n; k++) ip[k] = (float)t[k];   /* float data come back exactly: float -> double -> float */
  free(t);
  return st ? fail("nc_get_vara_float", f, NC_EIO) : NC_NOERR;
}
#define WHOLE_GET(NAME, T, VARA)                                                    \
  int NAME(int ncid, int v, T *ip)                                                  \
  {                                                                                 \
    nc3_file *f = F(ncid); size_t s[NC3_MAX_DIMS], c[NC3_MAX_DIMS];                 \
    if (!vok(f, v)) return NC_ENOTVAR;                                              \
    whole(f, v, s, c);                                                              \
    return VARA(ncid, v, s, c, ip);                                                 \
  }
WHOLE_GET(nc_get_var_text, char, nc_get_vara_text)
WHOLE_GET(nc_get_var_int, int, nc_get_vara_int)
WHOLE_GET(nc_get_var_double, double, nc_get_vara_double)
WHOLE_GET(nc_get_var_short, short, nc_get_vara_short)
WHOLE_GET(nc_get_var_float, float, nc_get_vara_float)

int nc_put_vara_text(int ncid, int v, const size_t *s, const size_t *c, const char *op) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_put_vara_text(f, v, s, c, op) ? fail("nc_put_vara_text", f, NC_EIO) : NC_NOERR; }
int nc_put_vara_int(int ncid, int v, const size_t *s, const size_t *c, const int *op) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_put_vara_int(f, v, s, c, op) ? fail("nc_put_vara_int", f, NC_EIO) : NC_NOERR; }
int nc_put_vara_double(int ncid, int v, const size_t *s, const size_t *c, const double *op) { nc3_file *f = F(ncid); if (!vok(f, v)) return NC_ENOTVAR; return nc3_put_vara_double(f, v, s, c, op) ? fail("nc_put_vara_double", f, NC_EIO) : NC_NOERR; }
int nc_put_vara_short(int ncid, int v, const size_t *s, const size_t *c, const short *op)
{
  nc3_file *f = F(ncid); size_t n, k; int *t; int st;
  if (!vok(f, v)) return NC_ENOTVAR;
  n = nelem(f, v, c); t = (int *)malloc((n + 1) * sizeof(int));
  for (k = 0; k < n; k++) t[k] = op[k];
  st = nc3_put_vara_int(f, v, s, c, t);
  free(t);
  return st ? fail("nc_put_vara_short", f, NC_EIO) : NC_NOERR;
}
#define WHOLE_PUT(NAME, T, VARA)                                                    \
  int NAME(int ncid, int v, const T *op)                                            \
  {                                                                                 \
    nc3_file *f = F(ncid); size_t s[NC3_MAX_DIMS], c[NC3_MAX_DIMS];                 \
    if (!vok(f, v)) return NC_ENOTVAR;                                              \
    whole(f, v, s, c);                                                              \
    return VARA(ncid, v, s, c, op);                                                 \
  }
WHOLE_PUT(nc_put_var_text, char, nc_put_vara_text)
WHOLE_PUT(nc_put_var_int, int, nc_put_vara_int)
WHOLE_PUT(nc_put_var_double, double, nc_put_vara_double)
WHOLE_PUT(nc_put_var_short, short, nc_put_vara_short)
