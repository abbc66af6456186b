// Plan object behind the opaque xgb_plan handle (include/xgrid_b200.h).  Internal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "xgrid_internal.h"
#include "xgrid_gc.h"

// grow-only device allocation
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);   // 0 on success
  void release();
};

struct xgb_apply_state;        // apply_capi.cu

struct xgb_plan {
  int device = 0;
  cudaStream_t st = nullptr;
  cudaStream_t copy_st = nullptr;             // result download overlapped with generation (xgb_plan_generate_to_host)
  cudaEvent_t copy_ev = nullptr;
  cudaStream_t aux_st = nullptr;              // heavy-cell kernels beside the bulk scatter / finalize kernels
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;

  // destination tile
  bool have_dst = false;
  int nx2 = 0, ny2 = 0;
  DevBuf dst_lon, dst_lat, dst_store, pyr_store;
  xgb::CellSet dst{};
  xgb::Pyramid pyr{};
  DevBuf rect_store, rect_rows, rect_invalid;  // separable destination tile: 1-D row / column boxes (xgrid_internal.h RectDst)
  xgb::RectDst rect{};
  bool rect_pending = false;                   // set_dst_latlon: the separability check's verdict is read at the next generate's sync
  int* rect_host = nullptr;                    // pinned

  // source mosaic
  bool have_src = false, has_mask = false;
  std::vector<xgb::TileDesc> tiles;
  DevBuf tiles_dev, src_lon, src_lat, mask, src_store;
  xgb::CellSet src{};
  long long s0 = 0, ns = 0;          // active window of source cells (first window, total count)
  xgb::SrcMap map{};                 // all active windows

  // great-circle extras (cartesian vertices, spherical-excess areas)
  bool gc_src_ready = false, gc_dst_ready = false;
  DevBuf gc_src_xyz, gc_dst_xyz, gc_pyr_store;
  xgb::GcCells gc_src{}, gc_dst{};
  xgb::Pyramid3 gc_pyr{};

  // work space
  DevBuf cnt, pair_off, pair_cnt, out_off, pairs, parea, pclon, pclat, scan_tmp, bounds_dev;
  DevBuf clip_vx, clip_vy, clip_meta;   // polygons between the two clip kernels (xgrid_kernels.cu clip_sh_kernel / clip_mom_kernel)
  size_t heavy_cap = 0;              // entries of the heavy-cell work lists (grown on overflow)
  size_t pairs_cap = 0;              // entries the pair buffers were last sized for (single-pass candidate search)
  DevBuf heavy_ctl, heavy_flag, heavy_list, heavy_items, heavy_pairs;

  // result (Interp_config layout)
  DevBuf t_in, i_in, j_in, i_out, j_out, area, clon, clat, di, dj;
  long long nxgrid = -1;
  unsigned long long npairs = 0;
  int order = 0;
  // order 2 over several output tiles: per-source-cell sums carried across generate calls (xgb_plan_order2_begin/_end)
  DevBuf o2_acc, o2_tmp;
  int o2_state = 0;                  // 0 off, 1 accumulating (generate leaves di/dj unset), 2 centroids ready

  // a window enqueued by xgb_plan_generate_async and not yet finished
  bool pending = false;
  int pending_order = 0;
  size_t pending_cap = 0;
  xgb::SrcMap pending_map{};
  xgb::HeavyWork pending_hw{};

  int* err_dev = nullptr;
  int* err_host = nullptr;                    // pinned
  unsigned* win_host = nullptr;               // pinned: out_off at window boundaries
  long long win_nx[xgb::kMaxWindows] = {0};   // exchange cells per window of the last generate
  int win_nx_n = 0;
  unsigned long long* total_dev = nullptr;    // [2]
  unsigned long long* total_host = nullptr;   // pinned [2]

  // per-phase device timing of generate (CUDA events on the plan stream)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float phase_ms[5] = {0, 0, 0, 0, 0};
  double phase_ms_sum[5] = {0, 0, 0, 0, 0};
  long long generates = 0;

  xgb_apply_state* apply = nullptr;
};

void xgb_set_error(const char* fmt, ...);
long long xgb_generate_great_circle(xgb_plan* p, int order);   // xgrid_gc_capi.cu
void xgb_apply_release(xgb_plan* p);                           // apply_capi.cu

namespace xgb {
extern long long g_launches;     // kernels launched by this library since load (xgrid_kernels.cu)
void launch_partition(const uint32_t* pair_off, long long ncell, unsigned long long total, int nparts,
                      long long* bounds, cudaStream_t st, const unsigned long long* targets = nullptr);
}
