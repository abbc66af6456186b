// Internal interface of the great-circle path (xgrid_gc_kernels.cu <-> xgrid_gc_capi.cu).  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "xgrid_internal.h"

namespace xgb {

struct __align__(16) Box3 { double lo[3], hi[3]; };

// cells of a mosaic for the great-circle path: corner k (reference's clockwise order), coordinate c of cell g at
// v[(3*k + c)*ncell + g]; grown xyz bounding box; spherical-excess area (get_grid_great_circle_area)
struct GcCells {
  long long ncell;
  double* v;        // [12][ncell]
  Box3* box;        // [ncell]
  double* area;     // [ncell]
};

struct Pyr3Level { int nx, ny; const Box3* box; };
struct Pyramid3 { int nlev; Pyr3Level lev[kMaxLevels]; };

void launch_gc_cell_precompute(const TileDesc& tile, const double* lon, const double* lat, GcCells cells, int* err, cudaStream_t st);
void launch_gc_pyramid_level(const Pyr3Level& child, Box3* out, int nx, int ny, cudaStream_t st);
size_t gc_slot_bytes(long long ns);
void launch_gc_candidates(bool fill, const GcCells& src, const GcCells& dst, long long s0, long long ns, const double* mask, const Pyramid3& pyr,
                          const uint32_t* pair_off, uint32_t* cnt, int2* pairs, int* err, cudaStream_t st, int* slots = nullptr);
// box candidates -> pairs no side separates: flag[p] per pair; then (after an exclusive scan of the flags into pos) the kept
// pairs in order and the per-source-cell offsets / counts rewritten for them
void launch_gc_filter(const GcCells& src, const GcCells& dst, const int2* pairs, unsigned long long npairs, long long s0, uint32_t* flag,
                      cudaStream_t st);
void launch_gc_compact(const int2* pairs, unsigned long long npairs, const uint32_t* flag, const uint32_t* pos, int2* kept, long long ns,
                       uint32_t* pair_off, uint32_t* pair_cnt, cudaStream_t st);
void launch_gc_clip(const GcCells& src, const GcCells& dst, const double* mask, const int2* pairs, unsigned long long npairs,
                    long long s0, double* parea, uint32_t* cnt, int* err, cudaStream_t st);

}  // namespace xgb
