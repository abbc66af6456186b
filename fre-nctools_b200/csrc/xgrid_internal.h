// Internal (C++) interface between the kernel TUs and the C-ABI layer.  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "xgrid_geom.cuh"

namespace xgb {

// The source cells one generate call works on: up to kMaxWindows contiguous ranges of the concatenated (tile-major,
// row-major) cell index space, visited in the order given.  Kernels index them by a window-relative number t in
// [0, total()); a single window is the common case.  Several windows let one GPU take an interleaved share of the
// mosaic (polar and mid-latitude cells cost differently per pair), and let the host-download path work in pieces.
constexpr int kMaxWindows = 64;
struct SrcMap {
  int nwin;
  long long begin[kMaxWindows];      // first source cell of window w
  long long cum[kMaxWindows + 1];    // window-relative index of the first cell of window w
  __host__ __device__ long long total() const { return cum[nwin]; }
  __host__ __device__ long long cell(long long t) const {
    if (nwin == 1) return begin[0] + t;
    int lo = 0, hi = nwin - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (cum[mid] <= t) lo = mid; else hi = mid - 1; }
    return begin[lo] + (t - cum[lo]);
  }
};

// One structured tile of a mosaic: nx*ny cells, (nx+1)*(ny+1) vertices, row-major.
struct TileDesc {
  int nx, ny;
  long long cell_off;   // first cell of this tile in the concatenated cell index space
  long long vert_off;   // first vertex in the concatenated vertex arrays
};

// fix_lon'd cells of a whole mosaic, struct-of-arrays in HBM.  Vertex k of cell c is
// vx[k*ncell + c]: a warp reading vertex k of 32 neighbouring cells reads 256 contiguous
// bytes, and planes 4..7 (only pole cells use them) are never touched by ordinary cells.
// latitude range of the 4 raw corners (create_xgrid.c:727-728) and longitude range after fix_lon
// (:731-732): one 32-byte record so a box test costs two 16-byte loads
struct __align__(16) Box { double ymin, ymax, xmin, xmax; };

struct CellSet {
  long long ncell;
  Box* box;                   // [ncell]
  double* xavg;               // mean longitude after fix_lon          (create_xgrid.c:733)
  double* area;               // poly_area of the fix_lon'd cell       (create_xgrid.c:66-88)
  unsigned char* nv;          // vertex count after fix_lon (3..8)
  double *vx, *vy;            // [kMaxV][ncell]
};

// Min/max pyramid over the destination tile's index space: level 0 = cells, level l+1 node
// (ix,iy) bounds level-l nodes (2ix..2ix+1, 2iy..2iy+1).  Upper levels only prune; the exact
// reference predicates run at level 0.
constexpr int kMaxLevels = 24;
struct PyrLevel { int nx, ny; const Box* box; };
struct Pyramid { int nlev; PyrLevel lev[kMaxLevels]; };

// error bits raised by kernels (read back by the C-ABI layer, which turns them into the
// reference's fatal errors)
enum : int {
  kErrTooManyVertices = 1,   // "n2_in is greater than MAX_V"        (create_xgrid.c:730)
  kErrParallelEdges   = 2,   // clip_2dx2d determinant < EPSLN30      (create_xgrid.c:1314)
  kErrClipOverflow    = 4,   // clipped polygon exceeded MV vertices
  kErrStackOverflow   = 8,   // candidate traversal stack exhausted (internal)
  kErrGcNotConvex     = 16,  // great-circle: grid box is not convex  (create_xgrid.c:1576-1579)
  kErrGcWalk          = 32,  // great-circle: polygon walk failed     (create_xgrid.c:1795,1822,1825)
  kErrGcNodePool      = 64,  // great-circle: node pool exhausted     (mosaic_util.c:1075)
  kErrHeavyOverflow   = 128, // heavy-cell work buffer exhausted (internal capacity)
  kErrBadIndex        = 256, // an exchange-grid list names a cell outside its grid
};

// device-side control block and work buffers of the level-synchronous "heavy cell" candidate path
// Separable destination tile (every regular lat-lon grid): the latitude range of a cell depends on its row only, the
// longitude range and mean longitude on its column only.  Checked cell by cell, bit for bit, when the destination is set
// (rect_* kernels); the candidate search then needs two 1-D searches instead of a walk of the 2-D pyramid.
struct RectDst {
  int valid;                       // 0: use the pyramid
  int nx, ny;
  const double *ymin, *ymax;       // [ny], non-decreasing
  const double *xmin, *xmax, *xavg;   // [nx], strictly increasing by more than kRectMinStep
  const unsigned char* row_ok;     // [ny] 0: some cell of the row departs from the column values (next to a pole): pyramid
};

struct HeavyCtl { unsigned long long total; unsigned nheavy; unsigned npairs; unsigned nitems[kMaxLevels]; };
struct HeavyWork {
  HeavyCtl* ctl;
  unsigned char* flag;   // [ns]
  int* list;             // [ns] window-relative source cells handed to the heavy path
  int2* items[2];        // ping-pong (heavy cell, node) work items, cap entries each
  int2* pairs;           // (source cell, destination cell) candidates found by the heavy path, cap entries
  unsigned cap;
};

// ---- launchers (xgrid_kernels.cu) ----------------------------------------------------------
// cells [c0, c1) of the tile (default: all of it)
void launch_cell_precompute(const TileDesc& tile, const double* lon, const double* lat,
                            CellSet cells, int* err, cudaStream_t st, long long c0 = 0, long long c1 = -1);
void launch_pyramid_level(const PyrLevel& child, Box* out, int nx, int ny, cudaStream_t st);
void launch_candidates_count(const CellSet& src, const SrcMap& sm, const double* mask,
                             const Pyramid& pyr, const CellSet& dst, uint32_t* cnt, const HeavyWork& hw, int* err, cudaStream_t st);
void launch_candidates_single(const CellSet& src, const SrcMap& sm, const double* mask,
                              const Pyramid& pyr, const RectDst& rect, const CellSet& dst, uint32_t* pair_off, uint32_t* pair_cnt, int2* pairs,
                              unsigned long long pair_cap, uint32_t* cursor, const HeavyWork& hw, int* err, cudaStream_t st);
void launch_clip(int order, const CellSet& src, const CellSet& dst, const double* mask,
                 const int2* pairs, unsigned long long npairs, const unsigned long long* npairs_dev, const SrcMap& sm,
                 double* parea, double* pclon, double* pclat, uint32_t* cnt, int* err, cudaStream_t st,
                 double* gvx = nullptr, double* gvy = nullptr, unsigned short* gmeta = nullptr);
// vertices of the scratch arrays gvx / gvy the two-kernel clip needs for a launch over npairs pairs (gmeta: npairs entries)
size_t clip_scratch_vertices(unsigned long long npairs);
void launch_scatter(int order, const int2* pairs, unsigned long long npairs,
                    const double* parea, const double* pclon, const double* pclat,
                    const uint32_t* pair_off, const uint32_t* pair_cnt, const uint32_t* out_off,
                    const TileDesc* tiles, int ntiles, const SrcMap& sm, int nx2,
                    int* t_in, int* i_in, int* j_in, int* i_out, int* j_out,
                    double* area, double* clon, double* clat, const HeavyWork* hw, cudaStream_t st,
                    cudaStream_t aux, cudaEvent_t fork, cudaEvent_t join, const unsigned long long* npairs_dev = nullptr);
void launch_order2_finalize(const CellSet& src, const SrcMap& sm, const uint32_t* out_off,
                            const double* area, const double* clon, const double* clat,
                            double* di, double* dj, const int* heavy_list, const unsigned* nheavy,
                            cudaStream_t st, cudaStream_t aux, cudaEvent_t fork, cudaEvent_t join);
// order 2 over several output tiles (conserve_interp.c:204-221, :319-358); acc = [area | clon | clat] x ncell source cells
void launch_order2_accumulate(const SrcMap& sm, const uint32_t* out_off, const double* area, const double* clon, const double* clat,
                              double* acc, long long ncell, cudaStream_t st);
void launch_order2_centroids(const CellSet& src, double* acc, cudaStream_t st);
void launch_order2_distance(long long n, const int* t_in, const int* i_in, const int* j_in, const TileDesc* tiles, int ntiles,
                            const double* area, const double* clon, const double* clat, const double* acc, long long ncell,
                            double* di, double* dj, int* err, cudaStream_t st);
// nwords 32-bit words from device memory to pinned host memory, by a kernel (not the copy engine)
void launch_publish(void* host_dst, const void* dev_src, int nwords, cudaStream_t st);
void launch_latlon_fill(int nlon, int nlat, double lonbegin, double lonend, double latbegin, double latend, double* lon, double* lat,
                        cudaStream_t st);
void launch_rect_setup(const CellSet& dst, int nx, int ny, double* store, unsigned char* row_ok, int* invalid, RectDst* out, cudaStream_t st);
void launch_window_counts(const uint32_t* out_off, const SrcMap& sm, const unsigned long long* total, long long* counts, cudaStream_t st);
void launch_publish_windows(void* host_dst, const void* out_off, const SrcMap& sm, cudaStream_t st);
// exclusive prefix sum of n uint32 counts; out has n+1 entries (out[n] = total, must fit 32 bits);
// the 64-bit total is also written to *total_dev.  tmp must hold scan_tmp_bytes(n).
size_t scan_tmp_bytes(long long n);
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, long long n, unsigned long long* total_dev,
                           void* tmp, cudaStream_t st);

}  // namespace xgb
