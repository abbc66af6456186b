// Binary layout of the reference's fregrid data contract, so that libxgrid_b200 can export setup_conserve_interp and
// do_scalar_conserve_interp with the reference's own signatures (tools/fregrid/conserve_interp.h) and an unchanged
// fregrid can call them.  Field order and types mirror tools/libfrencutils/globals.h:66-222 and
// tools/libfrencutils/mpp_domain.h:40-49 (nc_type is an int, STRING is 255, constant.h:25); the names carry an xgb_
// prefix because this header is never seen by the caller, only the layout matters.  tests/test_capi_cpu.py compares
// sizeof/offsetof of every member used here with the compiled reference's (oracle/ref_driver.c: ref_abi_layout).
#pragma once
#include <stddef.h>

#define XGB_STRING 255

typedef struct {                      /* Var_config, globals.h:66-98 */
  char name[XGB_STRING];
  int vid;
  int type;                           /* nc_type */
  int ndim;
  int index[4];
  int nz, nn, kstart, kend, lstart, lend;
  int has_naxis, has_zaxis, has_taxis;
  double missing, scale, offset;
  int has_missing;
  int interp_method;
  int cell_measures, cell_methods, use_volume;
  int area_vid, area_fid, area_has_taxis, area_has_naxis, area_has_zaxis;
  double area_missing;
  char area_name[XGB_STRING];
  int do_regrid, is_axis_data;
  int dimsize[5];
} xgb_Var_config;

typedef struct {                      /* Field_config, globals.h:100-110 */
  char* file;
  int* fid;
  int nvar;
  double* data;
  double* area;
  double* grad_x;
  double* grad_y;
  int* grad_mask;
  xgb_Var_config* var;
} xgb_Field_config;

typedef struct {                      /* Interp_config, globals.h:149-163 */
  size_t nxgrid;
  int *i_in, *j_in, *i_out, *j_out, *t_in;
  double *di_in, *dj_in, *area, *weight;
  int* index;
  char remap_file[XGB_STRING];
  int file_exist;
} xgb_Interp_config;

typedef struct {                      /* domain2D, mpp_domain.h:40-49 */
  int isc, iec, jsc, jec, isd, ied, jsd, jed, nxc, nyc, nxd, nyd, nxg, nyg;
  int *isclist, *ieclist, *jsclist, *jeclist;
  int xhalo, yhalo;
} xgb_domain2D;

typedef struct {                      /* Grid_config, globals.h:176-222 */
  int is_cyclic, is_tripolar, halo, nx, ny, nx_fine, ny_fine, isc, iec, jsc, jec, nxc, nyc;
  double *lonc, *latc, *lont, *latt, *xt, *yt, *xc, *yc, *zt, *dx, *dy, *area;
  double *lonc1D, *latc1D, *lont1D, *latt1D, *latt1D_fine;
  double *en_e, *en_n, *edge_w, *edge_e, *edge_s, *edge_n, *vlon_t, *vlat_t, *cosrot, *sinrot, *weight, *cell_area;
  int weight_exist, rotate;
  xgb_domain2D domain;
} xgb_Grid_config;
