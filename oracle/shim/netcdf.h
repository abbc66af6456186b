/* Types-only stand-in for <netcdf.h> (TEST INFRASTRUCTURE, oracle build only).
 * The reference's globals.h / mpp_io.h mention nc_type and a few NC_* constants
 * but the math path never calls libnetcdf.  Values follow the public netCDF-C ABI. */
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
#define NC_NOERR 0
#endif
