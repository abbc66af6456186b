/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Link-time stand-ins for the reference's libnetcdf-backed I/O layer (mpp_io.c, read_mosaic.c),
 * which cannot be built here (no libnetcdf).  The remap-file calls of setup_conserve_interp are backed by an in-memory
 * store (below) so that the reference's WRITE / READ branches can be run and compared with the product's files; every
 * other entry point is never reached by the path under test and aborts loudly. */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <stdarg.h>
#include "netcdf.h"

#define STUB_DIE(name) do { fprintf(stderr, "oracle/_ref: I/O stub %s called\n", name); abort(); } while (0)

/* ---- an in-memory stand-in for the netCDF files the reference's conservative setup writes and reads ----------------
 * setup_conserve_interp's WRITE branch (conserve_interp.c:368-443) defines dimensions and variables and puts whole variables
 * and (ncells, 1) column blocks; its READ branch (:62-125) goes through read_mosaic_xgrid_size / _order1 / _order2 and one
 * mpp_get_var_value("tile1").  The store below records exactly what the reference hands the I/O layer, so the tests can
 * (a) compare it, name for name and value for value, with the file the product writes, and (b) feed the reference's READ
 * branch the contents of a file the product wrote.  Nothing here touches a disk. */
#include <string.h>
#define MPP_WRITE 100                /* mpp_io.h:32-34 */
#define MPP_READ  200
#define S_MAXF 8
#define S_MAXD 8
#define S_MAXV 16
#define S_MAXA 4
typedef struct { char name[256]; long size; } SDim;
typedef struct {
  char name[256]; int type, ndim, dims[4], natts;
  char attn[S_MAXA][64], attv[S_MAXA][256];
  void *data; size_t nelem;
} SVar;
typedef struct { char name[512]; int ndims, nvars; SDim dims[S_MAXD]; SVar vars[S_MAXV]; } SFile;
static SFile sfiles[S_MAXF];
static int nsfiles = 0;

static size_t s_tsize(int type) { return type == NC_DOUBLE ? 8 : (type == NC_INT || type == NC_FLOAT) ? 4 : type == NC_SHORT ? 2 : 1; }
static SFile *s_file(int fid) { if (fid < 0 || fid >= nsfiles) STUB_DIE("bad file id"); return &sfiles[fid]; }
static void s_reset(SFile *f) { int v; for (v = 0; v < f->nvars; v++) free(f->vars[v].data); memset(f, 0, sizeof *f); }

int stub_find(const char *name) { int i; for (i = 0; i < nsfiles; i++) if (!strcmp(sfiles[i].name, name)) return i; return -1; }
int stub_new_file(const char *name)
{
  int i = stub_find(name);
  if (i < 0) { if (nsfiles >= S_MAXF) STUB_DIE("too many files"); i = nsfiles++; }
  s_reset(&sfiles[i]);
  strncpy(sfiles[i].name, name, sizeof sfiles[i].name - 1);
  return i;
}
void stub_clear(void) { int i; for (i = 0; i < nsfiles; i++) s_reset(&sfiles[i]); nsfiles = 0; }
int stub_ndims(int f) { return s_file(f)->ndims; }
const char *stub_dim_name(int f, int d) { return s_file(f)->dims[d].name; }
long stub_dim_size(int f, int d) { return s_file(f)->dims[d].size; }
int stub_nvars(int f) { return s_file(f)->nvars; }
const char *stub_var_name(int f, int v) { return s_file(f)->vars[v].name; }
int stub_var_type(int f, int v) { return s_file(f)->vars[v].type; }
int stub_var_ndim(int f, int v) { return s_file(f)->vars[v].ndim; }
int stub_var_dim(int f, int v, int k) { return s_file(f)->vars[v].dims[k]; }
int stub_var_natts(int f, int v) { return s_file(f)->vars[v].natts; }
const char *stub_var_att_name(int f, int v, int a) { return s_file(f)->vars[v].attn[a]; }
const char *stub_var_att_value(int f, int v, int a) { return s_file(f)->vars[v].attv[a]; }
const void *stub_var_data(int f, int v) { return s_file(f)->vars[v].data; }
long stub_var_nelem(int f, int v) { return (long)s_file(f)->vars[v].nelem; }
void stub_set_var_data(int f, int v, const void *src) { SVar *x = &s_file(f)->vars[v]; memcpy(x->data, src, x->nelem * s_tsize(x->type)); }

int  mpp_def_dim(int fid, const char *name, int size);
int  mpp_def_var(int fid, const char *name, nc_type type, int ndim, const int *dims, int natts, ...);
int stub_add_dim(int f, const char *name, long size) { return mpp_def_dim(f, name, (int)size); }
int stub_add_var(int f, const char *name, int type, int ndim, const int *dims) { return mpp_def_var(f, name, type, ndim, dims, 0); }

int  mpp_open(const char *file, int action)
{
  int i;
  if (action == MPP_WRITE) return stub_new_file(file);
  i = stub_find(file);
  if (i < 0) { fprintf(stderr, "oracle/_ref: mpp_open(%s): not in the in-memory store\n", file); abort(); }
  return i;
}
void mpp_close(int fid) { (void)s_file(fid); }
int  mpp_def_dim(int fid, const char *name, int size)
{
  SFile *f = s_file(fid);
  if (f->ndims >= S_MAXD) STUB_DIE("too many dimensions");
  strncpy(f->dims[f->ndims].name, name, 255); f->dims[f->ndims].size = size;
  return f->ndims++;
}
int  mpp_def_var(int fid, const char *name, nc_type type, int ndim, const int *dims, int natts, ...)
{
  SFile *f = s_file(fid);
  SVar *v;
  va_list ap;
  int k;
  if (f->nvars >= S_MAXV || ndim > 4 || natts > S_MAXA) STUB_DIE("variable table full");
  v = &f->vars[f->nvars];
  memset(v, 0, sizeof *v);
  strncpy(v->name, name, 255); v->type = type; v->ndim = ndim; v->natts = natts; v->nelem = 1;
  for (k = 0; k < ndim; k++) { v->dims[k] = dims[k]; v->nelem *= (size_t)f->dims[dims[k]].size; }
  va_start(ap, natts);
  for (k = 0; k < natts; k++) {              /* (name, value) string pairs, mpp_io.c mpp_def_var */
    strncpy(v->attn[k], va_arg(ap, const char *), 63);
    strncpy(v->attv[k], va_arg(ap, const char *), 255);
  }
  va_end(ap);
  v->data = calloc(v->nelem ? v->nelem : 1, s_tsize(type));
  return f->nvars++;
}
void mpp_def_global_att(int fid, const char *name, const char *val) { (void)fid; (void)name; (void)val; STUB_DIE("mpp_def_global_att"); }
void mpp_end_def(int fid) { (void)s_file(fid); }
int  mpp_field_exist(const char *file, const char *field) { (void)file; (void)field; STUB_DIE("mpp_field_exist"); return 0; }
void mpp_get_global_att(int fid, const char *name, void *val) { (void)fid; (void)name; (void)val; STUB_DIE("mpp_get_global_att"); }
void mpp_get_var_value(int fid, int vid, void *data)
{ SVar *v = &s_file(fid)->vars[vid]; memcpy(data, v->data, v->nelem * s_tsize(v->type)); }
int  mpp_get_varid(int fid, const char *name)
{
  SFile *f = s_file(fid);
  int v;
  for (v = 0; v < f->nvars; v++) if (!strcmp(f->vars[v].name, name)) return v;
  fprintf(stderr, "oracle/_ref: mpp_get_varid: no variable %s\n", name); abort();
  return -1;
}
void mpp_put_var_value(int fid, int vid, const void *data)
{ SVar *v = &s_file(fid)->vars[vid]; memcpy(v->data, data, v->nelem * s_tsize(v->type)); }
void mpp_put_var_value_block(int fid, int vid, const size_t *start, const size_t *n, const void *data)
{
  /* hyperslab of a variable of rank <= 2 (all the conservative setup writes): rows start[0]..+n[0], columns start[1]..+n[1] */
  SFile *f = s_file(fid);
  SVar *v = &f->vars[vid];
  const size_t ts = s_tsize(v->type);
  const size_t ncol = (v->ndim == 2) ? (size_t)f->dims[v->dims[1]].size : 1;
  const size_t c0 = (v->ndim == 2) ? start[1] : 0, nc = (v->ndim == 2) ? n[1] : 1;
  size_t r, c, k = 0;
  if (v->ndim < 1 || v->ndim > 2) STUB_DIE("mpp_put_var_value_block: rank");
  for (r = 0; r < n[0]; r++) for (c = 0; c < nc; c++, k++)
    memcpy((char *)v->data + ((start[0] + r) * ncol + c0 + c) * ts, (const char *)data + k * ts, ts);
}
void read_mosaic_contact(const char *f, int *a, int *b, int *c, int *d, int *e, int *g, int *h, int *i, int *j, int *k)
{ (void)f; (void)a; (void)b; (void)c; (void)d; (void)e; (void)g; (void)h; (void)i; (void)j; (void)k; STUB_DIE("read_mosaic_contact"); }
int  read_mosaic_ncontacts(const char *f) { (void)f; STUB_DIE("read_mosaic_ncontacts"); return 0; }
int  read_mosaic_ntiles(const char *f) { (void)f; STUB_DIE("read_mosaic_ntiles"); return 0; }
/* read_mosaic.c:330-337, :407-447, :518-558 restated over the store (the real ones call libnetcdf): same index shifts, same
 * division of the area by 4*pi*R^2 */
#include <math.h>
static const SVar *s_var(const char *file, const char *name)
{
  int f = stub_find(file), v;
  if (f < 0) { fprintf(stderr, "oracle/_ref: %s: not in the in-memory store\n", file); abort(); }
  for (v = 0; v < sfiles[f].nvars; v++) if (!strcmp(sfiles[f].vars[v].name, name)) return &sfiles[f].vars[v];
  fprintf(stderr, "oracle/_ref: %s has no variable %s\n", file, name); abort();
  return NULL;
}
int  read_mosaic_xgrid_size(const char *file)
{
  int f = stub_find(file), d;
  if (f < 0) { fprintf(stderr, "oracle/_ref: %s: not in the in-memory store\n", file); abort(); }
  for (d = 0; d < sfiles[f].ndims; d++) if (!strcmp(sfiles[f].dims[d].name, "ncells")) return (int)sfiles[f].dims[d].size;
  STUB_DIE("read_mosaic_xgrid_size: no ncells");
  return 0;
}
void read_mosaic_xgrid_order2(const char *file, int *i1, int *j1, int *i2, int *j2, double *a, double *di, double *dj)
{
  const int ncells = read_mosaic_xgrid_size(file);
  const int *c1 = (const int *)s_var(file, "tile1_cell")->data, *c2 = (const int *)s_var(file, "tile2_cell")->data;
  const double *ar = (const double *)s_var(file, "xgrid_area")->data;
  const double *dist = di ? (const double *)s_var(file, "tile1_distance")->data : NULL;
  const double garea = 4 * M_PI * 6371000.0 * 6371000.0;      /* RADIUS, constant.h:23 */
  int n;
  for (n = 0; n < ncells; n++) {
    i1[n] = c1[n * 2] - 1; j1[n] = c1[n * 2 + 1] - 1;
    i2[n] = c2[n * 2] - 1; j2[n] = c2[n * 2 + 1] - 1;
    if (di) { di[n] = dist[n * 2]; dj[n] = dist[n * 2 + 1]; }
    a[n] = ar[n] / garea;
  }
}
void read_mosaic_xgrid_order1(const char *file, int *i1, int *j1, int *i2, int *j2, double *a)
{ read_mosaic_xgrid_order2(file, i1, j1, i2, j2, a, NULL, NULL); }
