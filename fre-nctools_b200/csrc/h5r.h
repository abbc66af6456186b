/* Read-only subset of the HDF5 file format, enough for the netCDF-4 files fregrid is handed.
 *
 * The reference opens every input through libnetcdf (tools/libfrencutils/mpp_io.c:109-140) and its DEFAULT output format
 * is NC_FORMAT_NETCDF4_CLASSIC (mpp_io.c:52, :163-169), so the mosaic, supergrid, field and remap files of a stock FRE
 * workflow are HDF5 files.  This reader follows the published "HDF5 File Format Specification Version 3.0" and covers what
 * libnetcdf (HDF5 1.8 - 1.14, default library-version bounds) writes for the classic data model:
 *   superblock versions 0 - 3 (any user-block offset); object headers version 1 and 2 with continuation blocks;
 *   groups as symbol tables (B-tree v1 + local heap) and as link messages, compact or dense (fractal heap);
 *   dataspace v1 / v2, datatypes fixed-point / floating-point / string / variable-length string, fill value;
 *   layouts compact, contiguous and chunked (B-tree v1 chunk index) with the deflate, shuffle and fletcher32 filters;
 *   attributes v1 - v3, compact or dense, variable-length strings through the global heap; dimension scales
 *   (CLASS / NAME / DIMENSION_LIST / _Netcdf4Dimid) for the netCDF dimension names.
 * Refused by name: layout version 4 chunk indices (written only with libver bounds >= 1.10), szip and other filters,
 * compound / enum / array / opaque data, shared (committed) datatypes, external links, nested groups (ignored).
 * nc3.c presents such a file through the same nc3_* calls as a classic file (nc3_open dispatches on the magic number).
 */
#ifndef XGB_H5R_H
#define XGB_H5R_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H5R_MAX_DIMS 8
enum { H5R_INT = 0, H5R_FLOAT = 1, H5R_STRING = 3, H5R_VLEN_STRING = 9, H5R_OTHER = -1 };

typedef struct {
  int cls;                 /* H5R_INT, H5R_FLOAT, H5R_STRING (fixed length `size`), H5R_VLEN_STRING, H5R_OTHER */
  int size;                /* bytes per element in the file */
  int is_signed;           /* H5R_INT */
  int big_endian;
} h5r_type;

typedef struct {
  char *name;
  h5r_type type;
  int rank;                             /* 0: scalar */
  long long dims[H5R_MAX_DIMS];
  long long nelem;                      /* 0 for a null dataspace */
  unsigned char *data;                  /* nelem * type.size bytes as in the file; for H5R_VLEN_STRING: nelem NUL-terminated
                                           strings back to back (data_len bytes) */
  size_t data_len;
  /* DIMENSION_LIST: object-header addresses of the dimension scales attached to each dimension (first scale only) */
  int ndimrefs;
  unsigned long long dimrefs[H5R_MAX_DIMS];
} h5r_att;

typedef struct h5r_dset {
  char *name;
  unsigned long long addr;              /* object header address: the target of DIMENSION_LIST references */
  long long order;                      /* link creation order (-1 when the file does not track it) */
  h5r_type type;
  int rank;
  long long dims[H5R_MAX_DIMS], maxdims[H5R_MAX_DIMS];   /* maxdims -1: unlimited */
  int natts;
  h5r_att *atts;
  int supported;                        /* 0: carries something this subset does not read; `why` says what */
  char why[96];
  void *priv;
} h5r_dset;

typedef struct h5r_file h5r_file;

h5r_file *h5r_open(const char *path, char *err, size_t errlen);
void h5r_close(h5r_file *f);
int h5r_ndsets(const h5r_file *f);                         /* datasets linked from the root group */
const h5r_dset *h5r_dset_at(const h5r_file *f, int i);
int h5r_ngatts(const h5r_file *f);                         /* attributes of the root group */
const h5r_att *h5r_gatt_at(const h5r_file *f, int i);
/* hyperslab of dataset i: `out` receives count[0]*...*count[rank-1] elements of type.size bytes each, in HOST byte order
 * (strings: bytes as stored).  Unwritten chunks read as the fill value.  0 on success, else -1 and h5r_strerror. */
int h5r_read(h5r_file *f, int i, const size_t *start, const size_t *count, void *out);
const char *h5r_strerror(const h5r_file *f);

#ifdef __cplusplus
}
#endif
#endif
