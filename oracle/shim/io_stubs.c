/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Link-time stand-ins for the reference's libnetcdf-backed I/O layer (mpp_io.c, read_mosaic.c),
 * which cannot be built here (no libnetcdf).  The math path under test never reaches them;
 * any call is a bug in the harness, so each one aborts loudly. */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <stdarg.h>
#include "netcdf.h"

#define STUB_DIE(name) do { fprintf(stderr, "oracle/_ref: I/O stub %s called\n", name); abort(); } while (0)

int  mpp_open(const char *file, int action) { (void)file; (void)action; STUB_DIE("mpp_open"); return -1; }
void mpp_close(int fid) { (void)fid; STUB_DIE("mpp_close"); }
int  mpp_def_dim(int fid, const char *name, int size) { (void)fid; (void)name; (void)size; STUB_DIE("mpp_def_dim"); return -1; }
int  mpp_def_var(int fid, const char *name, nc_type type, int ndim, const int *dims, int natts, ...)
{ (void)fid; (void)name; (void)type; (void)ndim; (void)dims; (void)natts; STUB_DIE("mpp_def_var"); return -1; }
void mpp_def_global_att(int fid, const char *name, const char *val) { (void)fid; (void)name; (void)val; STUB_DIE("mpp_def_global_att"); }
void mpp_end_def(int fid) { (void)fid; STUB_DIE("mpp_end_def"); }
int  mpp_field_exist(const char *file, const char *field) { (void)file; (void)field; STUB_DIE("mpp_field_exist"); return 0; }
void mpp_get_global_att(int fid, const char *name, void *val) { (void)fid; (void)name; (void)val; STUB_DIE("mpp_get_global_att"); }
void mpp_get_var_value(int fid, int vid, void *data) { (void)fid; (void)vid; (void)data; STUB_DIE("mpp_get_var_value"); }
int  mpp_get_varid(int fid, const char *name) { (void)fid; (void)name; STUB_DIE("mpp_get_varid"); return -1; }
void mpp_put_var_value(int fid, int vid, const void *data) { (void)fid; (void)vid; (void)data; STUB_DIE("mpp_put_var_value"); }
void mpp_put_var_value_block(int fid, int vid, const size_t *start, const size_t *n, const void *data)
{ (void)fid; (void)vid; (void)start; (void)n; (void)data; STUB_DIE("mpp_put_var_value_block"); }
void read_mosaic_contact(const char *f, int *a, int *b, int *c, int *d, int *e, int *g, int *h, int *i, int *j, int *k)
{ (void)f; (void)a; (void)b; (void)c; (void)d; (void)e; (void)g; (void)h; (void)i; (void)j; (void)k; STUB_DIE("read_mosaic_contact"); }
int  read_mosaic_ncontacts(const char *f) { (void)f; STUB_DIE("read_mosaic_ncontacts"); return 0; }
int  read_mosaic_ntiles(const char *f) { (void)f; STUB_DIE("read_mosaic_ntiles"); return 0; }
int  read_mosaic_xgrid_size(const char *f) { (void)f; STUB_DIE("read_mosaic_xgrid_size"); return 0; }
void read_mosaic_xgrid_order1(const char *f, int *i1, int *j1, int *i2, int *j2, double *a)
{ (void)f; (void)i1; (void)j1; (void)i2; (void)j2; (void)a; STUB_DIE("read_mosaic_xgrid_order1"); }
void read_mosaic_xgrid_order2(const char *f, int *i1, int *j1, int *i2, int *j2, double *a, double *di, double *dj)
{ (void)f; (void)i1; (void)j1; (void)i2; (void)j2; (void)a; (void)di; (void)dj; STUB_DIE("read_mosaic_xgrid_order2"); }
