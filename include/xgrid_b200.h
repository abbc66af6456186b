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

/* create_xgrid.h:41  (create_xgrid.c:98-137) — spherical-excess cell areas of the great-circle path */
void get_grid_great_circle_area(const int *nlon, const int *nlat, const double *lon, const double *lat, double *area);

/* create_xgrid.h:78-80  (create_xgrid.c:1366-1466) — great-circle clipping; xgrid_clon/clat are zero-filled as in the
 * reference (:1444-1445) */
int create_xgrid_great_circle(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out,
                              double *xgrid_area, double *xgrid_clon, double *xgrid_clat);
int create_xgrid_great_circle_(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
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

/* create_xgrid.h:47-64 — exchange grids between a regular grid given by its 1-D cell bounds and a 2-D grid (runoff_regrid.c:472,
 * interp.c:277,329): lon_in/lat_in (1dx2d) resp. lon_out/lat_out (2dx1d) are nlon+1 / nlat+1 bounds, the other grid is 2-D vertex
 * arrays; mask_in lives on the input grid.  Emission order of the reference (1-D cells row-major, then 2-D cells row-major). */
int create_xgrid_1dx2d_order1(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out, double *xgrid_area);
int create_xgrid_1dx2d_order2(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out,
                              double *xgrid_area, double *xgrid_clon, double *xgrid_clat);
int create_xgrid_2dx1d_order1(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out, double *xgrid_area);
int create_xgrid_2dx1d_order2(const int *nlon_in, const int *nlat_in, const int *nlon_out, const int *nlat_out,
                              const double *lon_in, const double *lat_in, const double *lon_out, const double *lat_out,
                              const double *mask_in, int *i_in, int *j_in, int *i_out, int *j_out,
                              double *xgrid_area, double *xgrid_clon, double *xgrid_clat);
/* Fortran twins (create_xgrid.c:196, :296, :404, :504) */
int create_xgrid_1dx2d_order1_(const int *, const int *, const int *, const int *, const double *, const double *, const double *,
                               const double *, const double *, int *, int *, int *, int *, double *);
int create_xgrid_1dx2d_order2_(const int *, const int *, const int *, const int *, const double *, const double *, const double *,
                               const double *, const double *, int *, int *, int *, int *, double *, double *, double *);
int create_xgrid_2dx1d_order1_(const int *, const int *, const int *, const int *, const double *, const double *, const double *,
                               const double *, const double *, int *, int *, int *, int *, double *);
int create_xgrid_2dx1d_order2_(const int *, const int *, const int *, const int *, const double *, const double *, const double *,
                               const double *, const double *, int *, int *, int *, int *, double *, double *, double *);

/* conserve_interp.h:25-32 — fregrid's own L3 entry points over its Grid_config / Interp_config / Field_config structs
 * (globals.h:66-222; declared here with opaque struct tags: callers include the reference's headers, the library mirrors
 * their layout in csrc/fregrid_abi.h and a test pins sizeof/offsetof against the compiled reference).
 *   setup_conserve_interp (conserve_interp.c:42): the generate branch for all (output, input) tile pairs; interp[n].* are
 *     malloc'ed like the reference does (:238-258).  Order 2 with several output tiles sums the exchange cells of every
 *     output tile per source cell before the centroid correction, like :204-221 / :319-358.  READ (:62-125) and WRITE
 *     (:368-443) go through the classic-netCDF remap-file reader / writer of Part 4 (CDF-1/2/5 read and written; netCDF-4/HDF5
 *     files are read through csrc/h5r.c, not written).
 *   do_scalar_conserve_interp (conserve_interp.c:507): order 1 / order 2 per variable (interp_method), missing values,
 *     MONOTONIC, any nz; weight fields, cell_methods sum, cell_measures and TARGET as in :535-539, :572-585, :841-865. */
void setup_conserve_interp(int ntiles_in, const void *grid_in /* const Grid_config* */, int ntiles_out,
                           void *grid_out /* Grid_config* */, void *interp /* Interp_config* */, unsigned int opcode);
void do_scalar_conserve_interp(void *interp /* Interp_config* */, int varid, int ntiles_in, const void *grid_in,
                               int ntiles_out, const void *grid_out, const void *field_in /* const Field_config* */,
                               void *field_out /* Field_config* */, unsigned int opcode, int nz);
/* test hook: the shared-reciprocal division of the apply kernel (csrc/shared_div.cuh) against `/` on the device */
int xgb_shared_div_check(long long n, const double *a, const double *b, long long *nbad);
/* sizeof/offsetof of the mirrored structs (test hook) */
int xgb_abi_layout(size_t *out, int cap);

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
/* The same for fregrid's --nlon/--nlat/--lonBegin... regular output grid (get_output_grid_by_size, fregrid_util.c:588-603):
 * the vertices are computed on the device with the reference's arithmetic (bit-identical to xgb_latlon_grid), nothing is
 * uploaded.  Degrees. */
int xgb_plan_set_dst_latlon(xgb_plan *p, int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend);

/* Source mosaic: ntiles tiles concatenated, tile n has (nx[n]+1)*(ny[n]+1) vertices.
 * mask: per source cell (concatenated), NULL = all 1.0 (conserve_interp.c:160-161). */
int xgb_plan_set_src(xgb_plan *p, int ntiles, const int *nx, const int *ny,
                     const double *lon, const double *lat, const double *mask, int on_device);

/* Restrict generation to source cells [begin, end) of the concatenated (tile-major, row-major)
 * cell index space — the unit of multi-GPU sharding.  Default: all cells. */
int xgb_plan_set_src_window(xgb_plan *p, long long begin, long long end);

/* The source mosaic for a process that only ever generates the given windows (one GPU of a multi-GPU run): same cell index
 * space, window bounds and emission order as xgb_plan_set_src + xgb_plan_set_src_windows, but only the vertex rows the windows
 * touch are copied to the device and only their cells are precomputed (host pointers; pinned memory gives asynchronous
 * copies).  No host synchronisation; kernel-detected errors are reported by the next generate. */
int xgb_plan_set_src_sharded(xgb_plan *p, int ntiles, const int *nx, const int *ny, const double *lon, const double *lat,
                             const double *mask, int nwin, const long long *begin, const long long *end);

/* Several windows at once (at most 64), visited in the order given: one GPU's interleaved share of the mosaic.  The
 * result lists the windows' exchange cells one window after the other; xgb_plan_window_counts returns how many each
 * window produced (for global offsets when the pieces of several GPUs are put back into the serial order). */
int xgb_plan_set_src_windows(xgb_plan *p, int nwin, const long long *begin, const long long *end);
int xgb_plan_window_counts(xgb_plan *p, long long *counts);

/* Split the source cells into nparts contiguous windows of (nearly) equal candidate-pair count.
 * bounds receives nparts+1 cell indices (bounds[0] = 0, bounds[nparts] = ncells). */
int xgb_plan_partition(xgb_plan *p, int nparts, long long *bounds);
/* The same with unequal parts: window k receives share[k] / sum(share) of the candidate pairs (share == NULL: equal parts).
 * For cost-balanced sharding: a pair in a pole cap costs more than a mid-latitude pair (one warp enumerates the cap's rows,
 * its exchange cells are summed sequentially), so the rank that owns a pole is given fewer pairs
 * (fre-nctools_b200/distributed.py rebalance_shares; replaces the equal-rows decomposition of fregrid_util.c:489-492). */
int xgb_plan_partition_shares(xgb_plan *p, int nparts, const double *share, long long *bounds);

/* Generate the exchange grid of the current window.  opcode: XGB_CONSERVE_ORDER1 or
 * XGB_CONSERVE_ORDER2, optionally | XGB_GREAT_CIRCLE.  Returns nxgrid (>= 0) or -1.
 * Results stay in HBM, in the reference's emission order (tile, j_in, i_in, then j_out*nx+i_out). */
long long xgb_plan_generate(xgb_plan *p, unsigned int opcode);
/* The same without the wait.  xgb_plan_generate_async enqueues the window on the plan's stream with the buffer sizes the
 * last completed xgb_plan_generate settled on (one synchronous call has to come first); the result stays in HBM and more
 * work can be queued behind it.  xgb_plan_generate_finish waits, checks what the kernels reported and returns the count
 * (a buffer that turned out too small makes it repeat the window synchronously).  xgb_plan_window_counts_device writes the
 * per-window exchange-cell counts of the window last enqueued to a device array (nwin int64) on the plan's stream, so the
 * multi-GPU count all-gather needs no trip through the host. */
int xgb_plan_generate_async(xgb_plan *p, unsigned int opcode);
long long xgb_plan_generate_finish(xgb_plan *p);
int xgb_plan_window_counts_device(xgb_plan *p, long long *counts_dev);

/* The same for callers that want the result in HOST arrays (what setup_conserve_interp's callers need,
 * conserve_interp.c:236-257): the window is generated in nchunks consecutive pieces of source cells and each piece is
 * copied out on a second stream while the next is computed, so the PCIe transfer overlaps the kernels.  The arrays
 * hold `capacity` entries (di, dj, xgrid_clon, xgrid_clat may be NULL; all are ignored for order 1 except area);
 * use pinned memory for real overlap.  Not available with XGB_GREAT_CIRCLE.  Returns nxgrid or -1. */
long long xgb_plan_generate_to_host(xgb_plan *p, unsigned int opcode, int nchunks, long long capacity,
                                    int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                                    double *area, double *di, double *dj, double *xgrid_clon, double *xgrid_clat);

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

/* Order 2 with SEVERAL output tiles (setup_conserve_interp, conserve_interp.c:148-227 and :319-358): the reference adds the
 * exchange cells of every output tile into one (area, clon, clat) record per source cell — output tiles in order, list
 * order inside a tile — before the AREA_RATIO test and the centroid subtraction.  Between _begin and _end every order-2
 * xgb_plan_generate adds its cells to those sums (kept on the device) and leaves di/dj unset; the caller keeps each
 * tile's lists, areas and raw centroids (xgb_plan_result_host with di = dj = NULL, xgb_plan_result_centroids_host).
 * _end turns the sums into source-cell centroids; _distance then computes one tile's tile1_distance from its lists (host
 * arrays, n entries); _reset returns the plan to the one-tile behaviour.  xgb_plan_set_src also resets. */
int xgb_plan_order2_begin(xgb_plan *p);
int xgb_plan_order2_end(xgb_plan *p);
int xgb_plan_order2_distance(xgb_plan *p, long long n, const int *t_in, const int *i_in, const int *j_in, const double *area,
                             const double *xgrid_clon, const double *xgrid_clat, double *di, double *dj);
void xgb_plan_order2_reset(xgb_plan *p);

/* get_grid_great_circle_area on the device: which = 0 source cells (concatenated), 1 destination cells */
int xgb_plan_great_circle_area_host(xgb_plan *p, int which, double *area);

/* Cell areas computed on the device (get_grid_area): source cells concatenated / destination cells. */
int xgb_plan_src_area_host(xgb_plan *p, double *area);
int xgb_plan_dst_area_host(xgb_plan *p, double *area);

/* ------------------------------------------------------------------------------------------
 * Part 2a — one process, several GPUs (csrc/multi_gpu.cu).  Replaces the reference's MPI decomposition of
 * setup_conserve_interp: destination row bands per rank (fregrid_util.c:489-492), the order-2 gather of every rank's list
 * to every rank (conserve_interp.c:202-227) and the gather to the root for writing (:404-437).  The source cells are cut
 * into ngpus * windows_per_gpu windows of equal candidate-pair count, dealt round-robin; the windows' results in window order
 * are the serial list (same cells, same order, same bits as one GPU), written to pinned host arrays allocated here.
 * devices: NULL = 0 .. ngpus-1 (a device may be named twice).  Great circle runs on one device (refused here).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int by_size;                       /* 1: regular lat-lon grid built on the device (fregrid --nlon/--nlat, degrees) */
  int nlon, nlat;
  double lonbegin, lonend, latbegin, latend;
  int nx, ny;                        /* 0: vertex arrays [(ny+1)*(nx+1)], radians */
  const double *lon, *lat;
} xgb_dst_spec;
typedef struct {
  long long nxgrid;
  int *t_in, *i_in, *j_in, *i_out, *j_out;
  double *area, *di, *dj, *xgrid_clon, *xgrid_clat;     /* order 1: di .. xgrid_clat are NULL */
} xgb_host_xgrid;
int xgb_generate_multi_gpu(int ngpus, const int *devices, unsigned int opcode, const xgb_dst_spec *dst, int ntiles,
                           const int *nx, const int *ny, const double *lon, const double *lat, const double *mask,
                           int windows_per_gpu, xgb_host_xgrid *out);
void xgb_host_xgrid_free(xgb_host_xgrid *x);

/* ------------------------------------------------------------------------------------------
 * Part 2b — conservative apply (do_scalar_conserve_interp, conserve_interp.c:507-910) and the order-2
 * gradient terms (grad_c2l, gradient_c2l.c:58-118; calc_c2l_grid_info :368-454), batched over field-levels.
 *
 * Field layout ("one field-level" = what one reference call with nz == 1 consumes):
 *   order 1 : all source tiles concatenated, nx*ny each                         (Field_config.data, globals.h:104)
 *   order 2 : all source tiles concatenated, each WITH its one-cell halo, (nx+2)*(ny+2), halo already filled
 *             (fregrid_util.c:2166-2183 update_halo is the caller's job, as it is for the reference routine)
 *   grad_x, grad_y (double), grad_mask (int): tiles concatenated, nx*ny each     (globals.h:106-108)
 *   output  : nx_out*ny_out
 * A batch of nfields field-levels is nfields such arrays back to back.  Every field-level of a batch is
 * remapped exactly as one reference call would remap it: per destination cell the sums run in the order of the
 * exchange-grid list, so results are bit-identical to the reference.  Supported variants: order 1 / order 2,
 * missing values (conserve_interp.c:563-590, :745-789), conserve_order2_monotonic (:617-742, opcode |
 * XGB_MONOTONIC), and through xgb_plan_apply_options cell_methods "sum", cell_measures, a weight field and --target_grid.
 * on_device != 0: every pointer argument is a device pointer on the plan's device, work is only enqueued on
 * the plan's stream.  on_device == 0: host pointers; the call returns when the outputs are in host memory.
 * ---------------------------------------------------------------------------------------- */

/* Regroup the result of the last xgb_plan_generate by destination cell (one-time, device side). */
int xgb_plan_apply_setup(xgb_plan *p);

/* The same from exchange-grid lists the caller already has (the READ branch of setup_conserve_interp,
 * conserve_interp.c:63-128, or lists all-gathered from other GPUs): 0-based indices, area in m^2,
 * di/dj = tile1_distance (NULL for order 1).  No grids are needed for order-1 apply. */
int xgb_plan_set_xgrid(xgb_plan *p, int ntiles, const int *nx, const int *ny, int nx_out, int ny_out, long long nxgrid,
                       const int *t_in, const int *i_in, const int *j_in, const int *i_out, const int *j_out,
                       const double *area, const double *di, const double *dj, int on_device);
long long xgb_plan_apply_nxgrid(xgb_plan *p);

/* calc_c2l_grid_info (gradient_c2l.c:368) for every source tile on the device.  lont/latt: cell centres with a
 * one-cell halo, tiles concatenated, (nx+2)*(ny+2) each (Grid_config.lont/latt, fregrid_util.c:262-292); corners
 * are the plan's source grid.  Metrics agree with the reference to rounding, not bit for bit (see
 * csrc/c2l_grid_info.cu); use xgb_plan_grad_set_metrics to supply the reference's own. */
int xgb_plan_grad_setup(xgb_plan *p, const double *lont, const double *latt, int on_device);
/* Metrics of one tile in the reference's layouts (Grid_config.dx, dy, area, edge_w/e/s/n, en_n, en_e, vlon_t,
 * vlat_t; globals.h:193-209, sizes gradient_c2l.c:30-47). */
int xgb_plan_grad_set_metrics(xgb_plan *p, int tile, const double *dx, const double *dy, const double *area,
                              const double *edge_w, const double *edge_e, const double *edge_s, const double *edge_n,
                              const double *en_n, const double *en_e, const double *vlon, const double *vlat, int on_device);
int xgb_plan_grad_get_metrics(xgb_plan *p, int tile, double *dx, double *dy, double *area, double *edge_w, double *edge_e,
                              double *edge_s, double *edge_n, double *en_n, double *en_e, double *vlon, double *vlat);

/* grad_c2l with all four on_*_edge flags set, as fregrid calls it (fregrid_util.c:2197-2200), for nfields
 * field-levels; grad_mask (may be NULL) as fregrid_util.c:2203-2216. */
int xgb_plan_grad_c2l(xgb_plan *p, int nfields, const double *data, double *grad_x, double *grad_y, int *grad_mask,
                      int has_missing, double missing, int on_device);

/* do_scalar_conserve_interp for nfields field-levels.  opcode: XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2
 * [| XGB_MONOTONIC].  grad_* may be NULL for order 1; grad_mask may be NULL for order 2 without missing values.
 * Cells no exchange cell touches receive `missing` (-1e20 when has_missing == 0, conserve_interp.c:541). */
int xgb_plan_apply(xgb_plan *p, unsigned int opcode, int nfields, const double *data, const double *grad_x,
                   const double *grad_y, const int *grad_mask, int has_missing, double missing, double *out, int on_device);

/* Per-source-cell factors of do_scalar_conserve_interp for the following apply / regrid calls (conserve_interp.c:572-585,
 * :821-865), all optional:
 *   cell_methods   0 = mean, 1 = sum (CELL_METHODS_SUM, globals.h:64): entries are weighted by area / cell_area and the
 *                  destination value is the sum, not the mean
 *   weight         grid_in[].weight (weight_exist): multiplies every entry's area
 *   field_area     field_in[].area with cell_measures: entries are weighted by area * (field_area / cell_area); one
 *                  field-level per call, like the reference (nz == 1); area_missing as Var_config.area_missing
 *   src_cell_area  grid_in[].cell_area, tiles concatenated (NULL: the plan's own get_grid_area values)
 *   target_grid    opcode TARGET: multiply the result by (sum of exchange areas) / dst_cell_area (NULL: the plan's own)
 * Arrays are copied.  Call with (p, 0, NULL, NULL, NULL, 0, 0, NULL, 0) to return to the plain mean. */
int xgb_plan_apply_options(xgb_plan *p, int cell_methods, const double *weight, const double *src_cell_area,
                           const double *field_area, double area_missing, int target_grid, const double *dst_cell_area,
                           int on_device);

/* gradient (order 2) + apply in one call; gradients stay in HBM. */
int xgb_plan_regrid(xgb_plan *p, unsigned int opcode, int nfields, const double *data, int has_missing, double missing,
                    double *out, int on_device);

/* ------------------------------------------------------------------------------------------
 * Part 3 — input synthesis and self-checks (host side; not on the timed path).
 * ---------------------------------------------------------------------------------------- */

/* make_hgrid "gnomonic_ed" cubed sphere as fregrid reads it (make_hgrid/create_gnomonic_cubic_grid.c:101,
 * fregrid_util.c:227-241): lonc/latc 6*(ni+1)^2 vertices, lont/latt 6*ni^2 centres (may be NULL), radians. */
int xgb_cubed_sphere_grid(int ni, double *lonc, double *latc, double *lont, double *latt);
/* fregrid --nlon/--nlat regular output grid (fregrid_util.c:588-603); degrees in, radians out. */
int xgb_latlon_grid(int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend,
                    double *lonc, double *latc);
/* host build of the device great-circle clip (csrc/gc_clip.cuh: clip_2dx2d_great_circle + great_circle_area), so
 * CPU tests can run the product's algorithm against the compiled reference.  Returns the vertex count (< 0: the
 * condition on which the reference aborts). */
int xgb_gc_clip_host(const double *x1, const double *y1, const double *z1, int n1,
                     const double *x2, const double *y2, const double *z2, int n2,
                     double *xo, double *yo, double *zo, double *area);
/* sin/cos/sincos of csrc/ref_trig.cuh evaluated by the host build and by a device kernel (host pointers). */
void xgb_ref_trig_host(long long n, const double *x, double *s, double *c, double *ss, double *sc);
int  xgb_ref_trig_device(long long n, const double *x, double *s, double *c, double *ss, double *sc);
/* the single-evaluation-site variant used inside the clip kernel (ref_trig_site / ref_sin_small): s1 = sin-only mode,
   (ss, sc) = sincos mode, sm = ref_sin_small */
void xgb_ref_trig_site_host(long long n, const double *x, double *s1, double *ss, double *sc, double *sm);
int  xgb_ref_trig_site_device(long long n, const double *x, double *s1, double *ss, double *sc, double *sm);
/* host build of the clip kernel's one-pass poly_area / poly_ctrlon / poly_ctrlat (csrc/xgrid_geom.cuh poly_moments_site):
   out = {area, ctrlon, ctrlat} of the n-vertex polygon (order 1: area only) */
void xgb_poly_moments_site_host(int order, int n, const double *x, const double *y, double clon, double *out);

/* ------------------------------------------------------------------------------------------
 * Part 4 — remap files (host side).  The file fregrid writes after weight generation and reads back with
 * --remap_file (conserve_interp.c:382-438; read_mosaic.c:330-560), in the classic netCDF formats fregrid's
 * --format option names (classic, 64bit_offset; mpp_io.c:1526-1540) plus CDF-5.  netCDF-4/HDF5 files (the reference's default,
 * mpp_io.c:52) are read through csrc/h5r.c behind the same calls; writing them is refused with a message.  Layout: dims string=255, ncells, two=2; int tile1(ncells), int tile1_cell(ncells,two),
 * int tile2_cell(ncells,two), double xgrid_area(ncells) [m2], double tile1_distance(ncells,two) (order 2); 1-based on
 * disk, 0-based in memory.  setup_conserve_interp honours READ / WRITE through these.
 * ---------------------------------------------------------------------------------------- */
/* "classic" | "64bit_offset" | "cdf5" for the files written from now on (set_in_format); nonzero + xgb_last_error otherwise */
int xgb_set_nc_format(const char *name);
int xgb_get_nc_format(void);                               /* 1, 2 or 5 */
/* lists as setup_conserve_interp leaves them in Interp_config (0-based; i_out/j_out relative to isc/jsc) */
int xgb_remap_write(const char *path, int order, long long nxgrid, const int *t_in, const int *i_in, const int *j_in,
                    const int *i_out, const int *j_out, int isc, int jsc, const double *area, const double *di, const double *dj);
long long xgb_remap_size(const char *path);                /* read_mosaic_xgrid_size; < 0 on error */
/* read_mosaic_xgrid_order1/2 + the tile1 read and area rescale of conserve_interp.c:81-90: 0-based lists, area in m2 */
int xgb_remap_read(const char *path, int order, long long cap, int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                   double *area, double *di, double *dj);

/* gc_acos: acosl(x) rounded to double as spherical_angle produces it (mosaic_util.c:834), host build / device build of
 * csrc/gc_clip.cuh on n arguments (self-checks; the device version copies through temporary buffers) */
void xgb_gc_acos_host(long long n, const double *x, double *out);
int  xgb_gc_acos_device(long long n, const double *x_host, double *out_host);
/* the same for fn = 0 acos, 1 asin, 2 atan2(x[i], y[i]) as glibc returns them (calc_c2l_grid_info's great_circle_distance and
 * xyz2latlon, mosaic_util.c:228-252, :754-757); y may be NULL for fn 0 and 1 */
void xgb_gc_math_host(int fn, long long n, const double *x, const double *y, double *out);
int  xgb_gc_math_device(int fn, long long n, const double *x_host, const double *y_host, double *out_host);

/* ============================================================================================
 * Part 5 — make_coupler_mosaic's exchange grids (tools/make_coupler_mosaic/make_coupler_mosaic.c)
 * ============================================================================================
 * One call replaces the tool's exchange-grid loops and the sums behind its land_mask / ocean_mask files and its
 * tile1_distance / tile2_distance variables: the per-atmosphere-cell loop (:1250-1720: atm x lnd polygons, atm x ocn cells
 * scaled by ocn_frac, the land share 1 - ocn_frac of every ocean cell under an atm x lnd polygon), the land x ocean loop
 * (:2556-2692) and the parent-cell sums (:1739-2030, :2694-2808).  Default clip method only (clip_2dx2d); the tool keeps
 * reading the mosaics, deriving omask from the topography (:939-960), adding its artificial southern ocean row (:845-876) and
 * writing the files (:2131-2480, :2810-2960).  Lists come back in the order the tool writes them: atmosphere (land) cell
 * tile-major / row-major, then land / ocean cell tile-major / row-major; a file of the tool is the sub-list of one (tile1, tile2). */
typedef struct {
  int ntiles;
  const int *nx, *ny;            /* cells per tile */
  const double *lon, *lat;       /* radians; tile after tile, (ny+1) x (nx+1) vertices each, as get_global_grid returns them */
} xgb_mosaic_grid;

typedef struct {
  long long n;
  int *t1, *i1, *j1;             /* parent cell in the first mosaic: tile, i, j (0-based; the tool writes i+1, j+1) */
  int *t2, *i2, *j2;             /* parent cell in the second mosaic */
  double *area;                  /* xgrid_area, m^2 */
  double *d1i, *d1j, *d2i, *d2j; /* interp_order 2: tile1_distance / tile2_distance (lon, lat); NULL for order 1 */
} xgb_coupler_list;

typedef struct {
  xgb_coupler_list atmxlnd, atmxocn, lndxocn;      /* lndxocn is empty when lnd_same_as_atm */
  long long ncell_atm, ncell_lnd, ncell_ocn;
  double *area_atm, *area_lnd, *area_ocn;          /* get_grid_area of every cell (:277) */
  double *lnd_xarea;                               /* per land cell: sum of its atm x lnd areas (land_mask = lnd_xarea / area_lnd, :2079) */
  double *ocn_xarea;                               /* per ocean cell: sum of its atm x ocn areas (ocean_mask "areaX", :2037-2046) */
} xgb_coupler_result;

/* tile_nest: atmosphere tile left out of the parent sums (-1: none).  lnd_same_as_atm: the land model runs on the atmosphere
 * mosaic (lnd may be NULL; an atmosphere cell then meets its own land cell only, :1474-1486).  ocn_same_as_atm: ocean tile n is
 * searched for atmosphere tile n only (:1338-1345).  omask: ocean fraction of every ocean cell, tile after tile.  Results are
 * malloc'ed; release them with xgb_coupler_result_free. */
int  xgb_make_coupler_xgrid(int device, int interp_order, double area_ratio_thresh, int tile_nest, int lnd_same_as_atm,
                            int ocn_same_as_atm, const xgb_mosaic_grid *atm, const xgb_mosaic_grid *lnd,
                            const xgb_mosaic_grid *ocn, const double *omask, xgb_coupler_result *out);
void xgb_coupler_result_free(xgb_coupler_result *r);

#ifdef __cplusplus
}
#endif
#endif /* XGRID_B200_H */
