/* Classic netCDF (CDF-1 "classic", CDF-2 "64-bit offset", CDF-5 "64-bit data") reader and writer in plain C.
 *
 * fregrid reads its mosaic / grid / field files and writes remap and output files through libnetcdf
 * (reference tools/libfrencutils/mpp_io.c); `--format classic` and `--format 64bit_offset` select these on-disk formats
 * (mpp_io.c:163-175, 1526-1540).  This file implements exactly that on-disk format (the netCDF classic format
 * specification: big-endian header of dimension, attribute and variable lists, then the fixed-size variables in
 * definition order, then the records), so files written here are readable by libnetcdf and vice versa.  netCDF-4
 * (HDF5) files — the reference's default (mpp_io.c:52) — are recognised by their magic number and READ through h5r.c behind
 * these same calls (nc3_format() == 4; dimensions from the dimension scales, variables in creation order, int64 and wide
 * unsigned data as NC3_DOUBLE); they are never written.
 *
 * Not thread-safe per file; all functions return 0 on success and a negative value on error with the message
 * available from nc3_strerror(f) (or the err buffer of nc3_open / nc3_create).
 */
#ifndef XGB_NC3_H
#define XGB_NC3_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NC3_BYTE = 1, NC3_CHAR = 2, NC3_SHORT = 3, NC3_INT = 4, NC3_FLOAT = 5, NC3_DOUBLE = 6 };
#define NC3_GLOBAL (-1)
#define NC3_MAX_DIMS 8

typedef struct nc3_file nc3_file;

/* ---- reading ---- */
nc3_file *nc3_open(const char *path, char *err, size_t errlen);
int nc3_format(const nc3_file *f);                        /* 1, 2 or 5; 4 for a netCDF-4 file */
int nc3_ndims(const nc3_file *f);
int nc3_nvars(const nc3_file *f);
int nc3_dim_id(const nc3_file *f, const char *name);      /* -1 if absent */
long long nc3_dim_len(const nc3_file *f, int dimid);      /* current length (number of records for the unlimited one) */
const char *nc3_dim_name(const nc3_file *f, int dimid);
int nc3_unlimdim(const nc3_file *f);                      /* -1 if none */
int nc3_var_id(const nc3_file *f, const char *name);      /* -1 if absent */
const char *nc3_var_name(const nc3_file *f, int varid);
int nc3_var_type(const nc3_file *f, int varid);
int nc3_var_ndims(const nc3_file *f, int varid);
const int *nc3_var_dimids(const nc3_file *f, int varid);
int nc3_var_natts(const nc3_file *f, int varid);          /* varid = NC3_GLOBAL for global attributes */
const char *nc3_att_name(const nc3_file *f, int varid, int attnum);
/* attribute lookup: type and length (-1 if absent) */
int nc3_att_inq(const nc3_file *f, int varid, const char *name, int *type, long long *len);
int nc3_get_att_text(const nc3_file *f, int varid, const char *name, char *out, size_t outlen);   /* NUL-terminated */
int nc3_get_att_double(const nc3_file *f, int varid, const char *name, double *out, int maxn);     /* any numeric type */
/* hyperslab reads; numeric types are converted to the requested type */
int nc3_get_vara_double(nc3_file *f, int varid, const size_t *start, const size_t *count, double *out);
int nc3_get_vara_int(nc3_file *f, int varid, const size_t *start, const size_t *count, int *out);
int nc3_get_vara_text(nc3_file *f, int varid, const size_t *start, const size_t *count, char *out);
int nc3_get_var_double(nc3_file *f, int varid, double *out);  /* whole variable */
int nc3_get_var_int(nc3_file *f, int varid, int *out);

/* ---- writing ---- */
nc3_file *nc3_create(const char *path, int format, char *err, size_t errlen);       /* format 1, 2 or 5 */
int nc3_def_dim(nc3_file *f, const char *name, long long len);                        /* len 0 = unlimited; returns dimid */
int nc3_def_var(nc3_file *f, const char *name, int type, int ndims, const int *dimids); /* returns varid */
int nc3_put_att_text(nc3_file *f, int varid, const char *name, const char *text);
int nc3_put_att_double(nc3_file *f, int varid, const char *name, int type, int n, const double *vals);
/* copy one attribute (any type) / every attribute of a variable from an open file into a file in define mode */
int nc3_copy_att(const nc3_file *fin, int varid_in, const char *name, nc3_file *fout, int varid_out);
int nc3_copy_atts(const nc3_file *fin, int varid_in, nc3_file *fout, int varid_out);
int nc3_enddef(nc3_file *f);
int nc3_put_vara_double(nc3_file *f, int varid, const size_t *start, const size_t *count, const double *in);
int nc3_put_vara_int(nc3_file *f, int varid, const size_t *start, const size_t *count, const int *in);
int nc3_put_vara_text(nc3_file *f, int varid, const size_t *start, const size_t *count, const char *in);
int nc3_put_var_double(nc3_file *f, int varid, const double *in);
int nc3_put_var_int(nc3_file *f, int varid, const int *in);

int nc3_close(nc3_file *f);                               /* flushes numrecs; frees f */
const char *nc3_strerror(const nc3_file *f);

#ifdef __cplusplus
}
#endif
#endif
