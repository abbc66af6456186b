/* libxgrid_b200 — C ABI of the B200-native conservative-regridding engine.
 *
 * Drop-in boundary: the C function seam between fregrid (tools/fregrid/) and libfrencutils
 * (tools/libfrencutils/) of mlee03/FRE-NCtools.  Part 1 re-exports the reference's own entry
 * points with identical names, argument order, units (radians, 0-based indices, m^2) and
 * error behaviour ("FATAL Error: ..." on stderr, exit(1) — mosaic_util.c:57-65); a fregrid
 * built against this library instead of create_xgrid.o gets the CUDA path with no source
 * change.  Part 2 is the batched, device-resident interface the per-tile pointer API cannot
 * express (all tiles at once, 64-bit counts, results left in HBM, multi-GPU windows).
 *
 * There is no CPU fallback: every entry point needs a CUDA device and fails loudly without one.
 * Plain pointers and sizes only; no C++/torch types.
 */
#ifndef XGRID_B200_H
#define XGRID_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Part 1 — reference-signature entry points (host pointers, caller-allocated outputs).
 * ---------------------------------------------------------------------------------------- */

/* create_xgrid.h:39  (create_xgrid.c:45) — capacity the callers allocate for the outputs below */
int get_maxxgrid(void);

/* create_xgrid.h:40  (create_xgrid.c:66-88) */
void get_grid_area(const int *nlon, const int *nlat, const double *lon, const double *lat, double *area);

/* create_xgrid.h:66-69  (create_xgrid.c:621-871).  Returns the number of exchange cells. */
int create_xgrid_2dx2d_order1(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out, double *xgrid_area);

/* create_xgrid.h:70-73  (create_xgrid.c:893-1152) */
int create_xgrid_2dx2d_order2(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out,
                              double *xgrid_area, double *xgrid_clon, double *xgrid_clat);

/* Fortran-callable twins (create_xgrid.c:608, :881) */
int create_xgrid_2dx2d_order1_(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                               const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                               const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out, double *xgrid_area);
int create_xgrid_2dx2d_order2_(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                               const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                               const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out,
                               double *xgrid_area, double *xgrid_clon, double *xgrid_clat);

/* ------------------------------------------------------------------------------------------
 * Part 2 — batched / device-resident interface.
 *
 * A plan is bound to one CUDA device and one stream.  Typical use (what fregrid's
 * setup_conserve_interp, conserve_interp.c:42-500, does per output tile):
 *     p = xgb_plan_create(dev);
 *     xgb_plan_set_dst(p, nx2, ny2, lon_out, lat_out, 0);
 *     xgb_plan_set_src(p, ntiles, nx1[], ny1[], lon_in, lat_in, NULL, 0);
 *     n = xgb_plan_generate(p, XGB_CONSERVE_ORDER2);
 *     xgb_plan_result_host(p, t_in, i_in, j_in, i_out, j_out, area, di, dj);
 * All functions returning int return 0 on success, non-zero on error (xgb_last_error()).
 * ---------------------------------------------------------------------------------------- */

/* opcode bits, numerically equal to the reference's (globals.h:46-61) */
#define XGB_CONSERVE_ORDER1 1u
#define XGB_CONSERVE_ORDER2 2u
#define XGB_GREAT_CIRCLE    4096u
#define XGB_MONOTONIC       16384u

typedef struct xgb_plan xgb_plan;

int          xgb_device_count(void);
const char  *xgb_last_error(void);

xgb_plan    *xgb_plan_create(int device);
void         xgb_plan_destroy(xgb_plan *p);
/* the plan's cudaStream_t, for callers that enqueue their own work around it */
void        *xgb_plan_stream(xgb_plan *p);
int          xgb_plan_sync(xgb_plan *p);

/* Destination tile: (nx+1)*(ny+1) vertex longitudes/latitudes, radians, row-major
 * (Grid_config.lonc/latc, globals.h:188-189).  on_device != 0: pointers are device pointers. */
int xgb_plan_set_dst(xgb_plan *p, int nx, int ny, const double *lon, const double *lat, int on_device);

/* Source mosaic: ntiles tiles concatenated, tile n has (nx[n]+1)*(ny[n]+1) vertices.
 * mask: per source cell (concatenated), NULL = all 1.0 (conserve_interp.c:160-161). */
int xgb_plan_set_src(xgb_plan *p, int ntiles, const int *nx, const int *ny,
                     const double *lon, const double *lat, const double *mask, int on_device);

/* Restrict generation to source cells [begin, end) of the concatenated (tile-major, row-major)
 * cell index space — the unit of multi-GPU sharding.  Default: all cells. */
int xgb_plan_set_src_window(xgb_plan *p, long long begin, long long end);

/* Split the source cells into nparts contiguous windows of (nearly) equal candidate-pair count.
 * bounds receives nparts+1 cell indices (bounds[0] = 0, bounds[nparts] = ncells). */
int xgb_plan_partition(xgb_plan *p, int nparts, long long *bounds);

/* Generate the exchange grid of the current window.  opcode: XGB_CONSERVE_ORDER1 or
 * XGB_CONSERVE_ORDER2, optionally | XGB_GREAT_CIRCLE.  Returns nxgrid (>= 0) or -1.
 * Results stay in HBM, in the reference's emission order (tile, j_in, i_in, then j_out*nx+i_out). */
long long xgb_plan_generate(xgb_plan *p, unsigned int opcode);

/* candidate pairs examined by the last generate (clip-kernel work items) */
long long xgb_plan_last_npairs(xgb_plan *p);

/* Device time of the phases of generate, measured with CUDA events on the plan's stream:
 * [0] candidate count + scan, [1] candidate fill, [2] clip kernel, [3] scan of accepted counts,
 * [4] scatter + order-2 finalize.  last5: the last generate (ms); sum5/generates: accumulated
 * since xgb_plan_reset_phase_ms.  Any pointer may be NULL. */
int  xgb_plan_phase_ms(xgb_plan *p, float *last5, double *sum5, long long *generates);
void xgb_plan_reset_phase_ms(xgb_plan *p);

/* kernels launched by this library since it was loaded (all plans) */
long long xgb_kernel_launches(void);

/* DFMA throughput of `device` measured now (TFLOP/s, 2 flops per DFMA): the FP64 roofline denominator */
int xgb_fp64_peak_tflops(int device, double *tflops);

/* Device pointers to the last result (valid until the next generate/destroy).  Layout =
 * Interp_config (globals.h:149-163): t_in,i_in,j_in,i_out,j_out int32[nxgrid]; area f64 (m^2);
 * di,dj f64 (order 2 only: tile1_distance, conserve_interp.c:351-358; NULL for order 1). */
typedef struct {
  long long nxgrid;
  int *t_in, *i_in, *j_in, *i_out, *j_out;
  double *area, *di, *dj;
  double *xgrid_clon, *xgrid_clat;   /* order 2: raw centroid integrals (create_xgrid.c:1091-1092) */
} xgb_xgrid_view;
int xgb_plan_result_device(xgb_plan *p, xgb_xgrid_view *view);

/* Copy the last result to caller-allocated host arrays of length nxgrid (any may be NULL). */
int xgb_plan_result_host(xgb_plan *p, int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                         double *area, double *di, double *dj);
int xgb_plan_result_centroids_host(xgb_plan *p, double *xgrid_clon, double *xgrid_clat);

/* Cell areas computed on the device (get_grid_area): source cells concatenated / destination cells. */
int xgb_plan_src_area_host(xgb_plan *p, double *area);
int xgb_plan_dst_area_host(xgb_plan *p, double *area);

/* ------------------------------------------------------------------------------------------
 * Part 3 — input synthesis and self-checks (host side; not on the timed path).
 * ---------------------------------------------------------------------------------------- */

/* make_hgrid "gnomonic_ed" cubed sphere as fregrid reads it (make_hgrid/create_gnomonic_cubic_grid.c:101,
 * fregrid_util.c:227-241): lonc/latc 6*(ni+1)^2 vertices, lont/latt 6*ni^2 centres (may be NULL), radians. */
int xgb_cubed_sphere_grid(int ni, double *lonc, double *latc, double *lont, double *latt);
/* fregrid --nlon/--nlat regular output grid (fregrid_util.c:588-603); degrees in, radians out. */
int xgb_latlon_grid(int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend,
                    double *lonc, double *latc);
/* sin/cos/sincos of csrc/ref_trig.cuh evaluated by the host build and by a device kernel (host pointers). */
void xgb_ref_trig_host(long long n, const double *x, double *s, double *c, double *ss, double *sc);
int  xgb_ref_trig_device(long long n, const double *x, double *s, double *c, double *ss, double *sc);

#ifdef __cplusplus
}
#endif
#endif /* XGRID_B200_H */
