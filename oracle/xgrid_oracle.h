/* TEST INFRASTRUCTURE — CPU oracle for the fregrid conservative-regridding hot path.
 *
 * This is a plain-C restatement of the reference's algorithm (mlee03/FRE-NCtools,
 * tools/libfrencutils/create_xgrid.c, mosaic_util.c, gradient_c2l.c and
 * tools/fregrid/conserve_interp.c).  It exists only so tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg can check the CUDA product; nothing under
 * fre-nctools_b200/ may include, link or call it.
 *
 * Parity pin: every function here is checked against the UNMODIFIED reference compiled from
 * /root/reference (oracle/_ref/libfrenc_ref.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py,
 * and against golden vectors generated from that reference (tests/golden/, tests/golden/make_golden.py).
 */
#ifndef XGRID_ORACLE_H
#define XGRID_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CONSERVE_ORDER1 1u
#define ORC_CONSERVE_ORDER2 2u
#define ORC_GREAT_CIRCLE    4096u
#define ORC_MONOTONIC       16384u

/* polygon primitives */
int    orc_fix_lon(double lon[], double lat[], int n, double tlon);
double orc_poly_area(const double lon[], const double lat[], int n);
double orc_poly_ctrlon(const double lon[], const double lat[], int n, double clon);
double orc_poly_ctrlat(const double lon[], const double lat[], int n);
int    orc_clip_2dx2d(const double lon1[], const double lat1[], int n1,
                      const double lon2[], const double lat2[], int n2,
                      double lon_out[], double lat_out[]);
void   orc_get_grid_area(int nlon, int nlat, const double *lon, const double *lat, double *area);

/* exchange-grid generators; return the count, or -1 if `cap` slots were not enough */
long orc_create_xgrid_2dx2d(int order, int nlon_in, int nlat_in, int nlon_out, int nlat_out,
                            const double *lon_in, const double *lat_in,
                            const double *lon_out, const double *lat_out, const double *mask_in,
                            long cap, int *i_in, int *j_in, int *i_out, int *j_out,
                            double *xarea, double *xclon, double *xclat);

/* great-circle path */
int    orc_clip_2dx2d_great_circle(const double x1[], const double y1[], const double z1[], int n1,
                                   const double x2[], const double y2[], const double z2[], int n2,
                                   double xo[], double yo[], double zo[]);
double orc_great_circle_area(int n, const double *x, const double *y, const double *z);
void   orc_get_grid_great_circle_area(int nlon, int nlat, const double *lon, const double *lat, double *area);
long   orc_create_xgrid_great_circle(int nlon_in, int nlat_in, int nlon_out, int nlat_out,
                                     const double *lon_in, const double *lat_in,
                                     const double *lon_out, const double *lat_out, const double *mask_in,
                                     long cap, int *i_in, int *j_in, int *i_out, int *j_out,
                                     double *xarea, double *xclon, double *xclat);

/* whole-mosaic setup (setup_conserve_interp): tiles concatenated; one destination tile.
 * di/dj may be NULL for order 1.  Returns nxgrid or -1 on overflow of `cap`. */
long orc_setup_conserve_interp(int ntiles_in, const int *nx_in, const int *ny_in,
                               const double *lonc_in, const double *latc_in,
                               int nx_out, int ny_out, const double *lonc_out, const double *latc_out,
                               unsigned int opcode, long cap,
                               int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                               double *area, double *di, double *dj);

/* the same, also returning the generators' raw xgrid_clon / xgrid_clat (order 2; may be NULL) */
long orc_setup_conserve_interp_ex(int ntiles_in, const int *nx_in, const int *ny_in,
                                  const double *lonc_in, const double *latc_in,
                                  int nx_out, int ny_out, const double *lonc_out, const double *latc_out,
                                  unsigned int opcode, long cap,
                                  int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                                  double *area, double *di, double *dj, double *xclon_out, double *xclat_out);
/* the order-2 centroid correction alone (conserve_interp.c:204-221, :319-358) for a finished list of one output tile */
void orc_order2_distance(int ntiles_in, const int *nx_in, const int *ny_in, const double *lonc_in, const double *latc_in,
                         long n, const int *t_in, const int *i_in, const int *j_in,
                         const double *area, const double *xclon, const double *xclat, double *di, double *dj);

/* apply (do_scalar_conserve_interp, mean cell_methods, no cell_measures/weights/target) */
void orc_conserve_apply(int order, long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                        const int *i_out, const int *j_out, const double *area,
                        const double *di, const double *dj,
                        int ntiles_in, const int *nx_in, const int *ny_in,
                        const double *data_in, const double *grad_x, const double *grad_y,
                        const int *grad_mask, int has_missing, double missing, int monotonic,
                        int nx_out, int ny_out, int nz, double *data_out);

/* the same with weight field / cell_methods sum / cell_measures / target-grid rescale, one field-level */
void orc_conserve_apply_ex(int order, long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                           const int *i_out, const int *j_out, const double *area,
                           const double *di, const double *dj,
                           int ntiles_in, const int *nx_in, const int *ny_in,
                           const double *data_in, const double *grad_x, const double *grad_y,
                           const int *grad_mask, int has_missing, double missing, int monotonic,
                           int cell_methods, const double *weight, const double *cell_area, const double *farea,
                           int target_grid, const double *dst_cell_area,
                           int nx_out, int ny_out, double *data_out);

/* order-2 gradient terms */
void orc_calc_c2l_grid_info(int nx, int ny, const double *xt, const double *yt,
                            const double *xc, const double *yc,
                            double *dx, double *dy, double *area,
                            double *edge_w, double *edge_e, double *edge_s, double *edge_n,
                            double *en_n, double *en_e, double *vlon, double *vlat);
void orc_grad_c2l(int nx, int ny, const double *pin, const double *dx, const double *dy,
                  const double *area, const double *edge_w, const double *edge_e,
                  const double *edge_s, const double *edge_n, const double *en_n, const double *en_e,
                  const double *vlon, const double *vlat, double *grad_x, double *grad_y);

void orc_grad_mask(int nx, int ny, const double *pin, double missing, int *mask);
double orc_spherical_angle(const double v1[3], const double v2[3], const double v3[3]);

/* host libm, element-wise (pins csrc/ref_trig.cuh) */
void orc_libm_trig(long n, const double *x, double *s, double *c, double *ss, double *sc);

#ifdef __cplusplus
}
#endif
#endif
