/* Stand-in for <netcdf.h> (TEST INFRASTRUCTURE, oracle build only).
 *
 * There is no libnetcdf / HDF5 in this image.  The reference's math files only need nc_type and a few constants; its I/O
 * layer (tools/libfrencutils/mpp_io.c, read_mosaic.c) and the tools' main programs (fregrid.c, fregrid_util.c,
 * make_coupler_mosaic.c) call about fifty netCDF-C functions.  This header declares exactly those, with the public netCDF-C
 * signatures and constant values; oracle/shim/nc_shim.c implements them over the classic-format reader / writer
 * csrc/nc3.c, so the UNMODIFIED reference tools compile and run here on classic (CDF-1/2/5) files.  netCDF-4/HDF5 does not
 * exist in this shim: a create request for it yields a 64-bit-offset file. */
#ifndef ORACLE_SHIM_NETCDF_H
#define ORACLE_SHIM_NETCDF_H
#include <stddef.h>
typedef int nc_type;
#define NC_NAT    0
#define NC_BYTE   1
#define NC_CHAR   2
#define NC_SHORT  3
#define NC_INT    4
#define NC_FLOAT  5
#define NC_DOUBLE 6
#define NC_FILL_INT    (-2147483647)
#define NC_FILL_DOUBLE (9.9692099683868690e+36)
#define NC_MAX_NAME 256
#define NC_MAX_VAR_DIMS 1024
#define NC_NOERR 0
#define NC_GLOBAL (-1)
#define NC_UNLIMITED 0L
#define NC_NOWRITE 0x0000
#define NC_WRITE 0x0001
#define NC_CLOBBER 0x0000
#define NC_NOCLOBBER 0x0004
#define NC_64BIT_OFFSET 0x0200
#define NC_CLASSIC_MODEL 0x0100
#define NC_NETCDF4 0x1000
#define NC_FORMAT_CLASSIC 1
#define NC_FORMAT_64BIT 2
#define NC_FORMAT_64BIT_OFFSET 2
#define NC_FORMAT_NETCDF4 3
#define NC_FORMAT_NETCDF4_CLASSIC 4
#define NC_FORMAT_64BIT_DATA 5
#define NC_EBADID (-33)
#define NC_ENOTVAR (-49)
#define NC_ENOTATT (-43)
#define NC_EBADDIM (-46)
#define NC_EINVAL (-36)
#define NC_EIO (-68)

const char *nc_strerror(int status);
int nc_open(const char *path, int mode, int *ncidp);
int nc_create(const char *path, int cmode, int *ncidp);
int nc__create(const char *path, int cmode, size_t initialsz, size_t *chunksizehintp, int *ncidp);
int nc_close(int ncid);
int nc_sync(int ncid);
int nc_redef(int ncid);
int nc_enddef(int ncid);
int nc__enddef(int ncid, size_t h_minfree, size_t v_align, size_t v_minfree, size_t r_align);
int nc_inq_format(int ncid, int *formatp);
int nc_inq_nvars(int ncid, int *nvarsp);
int nc_inq_unlimdim(int ncid, int *unlimdimidp);
int nc_inq_dimid(int ncid, const char *name, int *idp);
int nc_inq_dimlen(int ncid, int dimid, size_t *lenp);
int nc_inq_dimname(int ncid, int dimid, char *name);
int nc_inq_varid(int ncid, const char *name, int *varidp);
int nc_inq_varname(int ncid, int varid, char *name);
int nc_inq_vartype(int ncid, int varid, nc_type *xtypep);
int nc_inq_varndims(int ncid, int varid, int *ndimsp);
int nc_inq_vardimid(int ncid, int varid, int *dimidsp);
int nc_inq_varnatts(int ncid, int varid, int *nattsp);
int nc_inq_att(int ncid, int varid, const char *name, nc_type *xtypep, size_t *lenp);
int nc_inq_atttype(int ncid, int varid, const char *name, nc_type *xtypep);
int nc_inq_attlen(int ncid, int varid, const char *name, size_t *lenp);
int nc_inq_attname(int ncid, int varid, int attnum, char *name);
int nc_get_att_text(int ncid, int varid, const char *name, char *ip);
int nc_get_att_double(int ncid, int varid, const char *name, double *ip);
int nc_get_att_int(int ncid, int varid, const char *name, int *ip);
int nc_get_att_short(int ncid, int varid, const char *name, short *ip);
int nc_put_att_text(int ncid, int varid, const char *name, size_t len, const char *op);
int nc_put_att_double(int ncid, int varid, const char *name, nc_type xtype, size_t len, const double *op);
int nc_copy_att(int ncid_in, int varid_in, const char *name, int ncid_out, int varid_out);
int nc_def_dim(int ncid, const char *name, size_t len, int *idp);
int nc_def_var(int ncid, const char *name, nc_type xtype, int ndims, const int *dimidsp, int *varidp);
int nc_def_var_deflate(int ncid, int varid, int shuffle, int deflate, int deflate_level);
int nc_inq_var_deflate(int ncid, int varid, int *shufflep, int *deflatep, int *deflate_levelp);
int nc_get_var_text(int ncid, int varid, char *ip);
int nc_get_var_int(int ncid, int varid, int *ip);
int nc_get_var_double(int ncid, int varid, double *ip);
int nc_get_var_short(int ncid, int varid, short *ip);
int nc_get_var_float(int ncid, int varid, float *ip);
int nc_get_vara_text(int ncid, int varid, const size_t *startp, const size_t *countp, char *ip);
int nc_get_vara_int(int ncid, int varid, const size_t *startp, const size_t *countp, int *ip);
int nc_get_vara_double(int ncid, int varid, const size_t *startp, const size_t *countp, double *ip);
int nc_get_vara_short(int ncid, int varid, const size_t *startp, const size_t *countp, short *ip);
int nc_get_vara_float(int ncid, int varid, const size_t *startp, const size_t *countp, float *ip);
int nc_put_var_text(int ncid, int varid, const char *op);
int nc_put_var_int(int ncid, int varid, const int *op);
int nc_put_var_double(int ncid, int varid, const double *op);
int nc_put_var_short(int ncid, int varid, const short *op);
int nc_put_vara_text(int ncid, int varid, const size_t *startp, const size_t *countp, const char *op);
int nc_put_vara_int(int ncid, int varid, const size_t *startp, const size_t *countp, const int *op);
int nc_put_vara_double(int ncid, int varid, const size_t *startp, const size_t *countp, const double *op);
int nc_put_vara_short(int ncid, int varid, const size_t *startp, const size_t *countp, const short *op);
#endif
