// Internal interface of the conservative apply path (apply_kernels.cu <-> apply_capi.cu).  Not installed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "xgrid_internal.h"

namespace xgb {

constexpr int kApplyBT = 8;    // field-levels per thread in apply: weights are read once per kApplyBT fields
constexpr int kGradBT = 4;     // field-levels per thread in grad_c2l: metrics are read once per kGradBT fields

enum : int {
  kErrApplyIndex  = 1 << 10,   // exchange-grid entry points outside the source mosaic
  kErrMonotoneMax = 1 << 11,   // " xdata is greater than f_bar_max "   (conserve_interp.c:693)
  kErrMonotoneMin = 1 << 12,   // " xdata is less than f_bar_min "      (conserve_interp.c:707)
  kErrAreaMissing = 1 << 13,   // "data is not missing but area is missing" (conserve_interp.c:578, :772)
};

// where a source tile's cells live in the concatenated field arrays
struct ApplyTile {
  int nx, ny;
  long long cell_off;   // first cell in the no-halo concatenation (order-1 data, grad_x, grad_y, grad_mask)
  long long halo_off;   // first element in the concatenation of (nx+2)*(ny+2) haloed tiles (order-2 data)
};

// exchange grid regrouped by destination cell; entry q of cell d: off[d] <= q < off[d+1]
struct ApplyCsr {
  const uint32_t* off;      // [ndst + 1]
  const uint32_t* perm;     // [nxgrid] position of entry q in the original list
  int* cell;                // concatenated source cell (no halo)
  int* hidx;                // element in the haloed concatenation
  double* area;
  double* di;               // order 2 only
  double* dj;
};

// grad_c2l metrics of one tile, reference layouts (gradient_c2l.c:30-47)
struct GradTile {
  int nx, ny;
  long long cell_off, halo_off;
  const double *dx, *dy, *area, *edge_w, *edge_e, *edge_s, *edge_n, *en_n, *en_e, *vlon, *vlat;
};

void launch_dst_count(long long n, const int* i_out, const int* j_out, int nx2, int ny2, uint32_t* cnt, int* err, cudaStream_t st);
void launch_dst_fill(long long n, const int* i_out, const int* j_out, int nx2, int ny2, const uint32_t* off, uint32_t* cursor,
                     uint32_t* perm, cudaStream_t st);
void launch_dst_sort_gather(long long ndst, const uint32_t* off, uint32_t* perm, const int* t_in, const int* i_in, const int* j_in,
                            const double* area, const double* di, const double* dj, const ApplyTile* tiles, int ntiles,
                            ApplyCsr csr, int* err, cudaStream_t st);
void launch_apply(int order, bool has_missing, bool from_xdata, const ApplyCsr& csr, long long ndst, int nf,
                  const double* data, long long data_stride, const double* gx, const double* gy, const int* gmask,
                  long long ncell_src, const double* xdata, long long nxgrid, double missing, double* out, cudaStream_t st,
                  int sum_mode = 0);
void launch_grad_c2l(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride,
                     double* gx, double* gy, int* gmask, bool has_missing, double missing, cudaStream_t st);
// fused path: gradient kernel writes (value, grad_x, grad_y, grad_mask) per cell as one double4, apply gathers it
void launch_grad_c2l_packed(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride,
                            double* packed, bool has_missing, double missing, cudaStream_t st);
void launch_apply_packed(bool has_missing, const ApplyCsr& csr, long long ndst, int nf, const double* packed, long long ncell_src,
                         double missing, double* out, cudaStream_t st, int sum_mode, int nx_out);
// field-transposed order-2 path: gradient + transpose into per-source-cell records, then the warp-per-destination-cell apply
int shared_div_check(long long n, const double* a_host, const double* b_host, unsigned long long* nbad_host);
size_t apply_rec_doubles(long long ncell, int nf, bool has_missing);
void launch_regrid_rec(const GradTile* tiles, int ntiles, long long ncell, int nf, const double* data, long long data_stride, double* rec,
                       bool has_missing, double missing, const ApplyCsr& csr, long long ndst, double apply_missing, int sum_mode,
                       double* out, cudaStream_t st, int nx2 = 0);
void launch_effective_area(const ApplyCsr& csr, long long n, const double* weight, const double* carea, const double* farea,
                           int sum_mode, double* eff, cudaStream_t st);
void launch_measure_check(const ApplyCsr& csr, long long n, int order, const double* data, const double* farea, double missing,
                          double area_missing, int* err, cudaStream_t st);
void launch_target_scale(const ApplyCsr& csr, long long ndst, int nf, const double* farea, const double* carea, const double* dst_carea,
                         double missing, double* out, cudaStream_t st);
void launch_monotone(long long nxgrid, const int* t_in, const int* i_in, const int* j_in, const double* di, const double* dj,
                     const ApplyTile* tiles, int ntiles, long long ncell, const double* data, const double* gx, const double* gy,
                     const int* gmask, double missing, double* fbmax, double* fbmin, unsigned long long* fmax_key,
                     unsigned long long* fmin_key, double* xdata, int* err, cudaStream_t st);
// calc_c2l_grid_info (gradient_c2l.c:368-454) for one tile
void launch_c2l_grid_info(int nx, int ny, const double* xt, const double* yt, const double* xc, const double* yc,
                          double* dx, double* dy, double* area, double* edge_w, double* edge_e, double* edge_s, double* edge_n,
                          double* en_n, double* en_e, double* vlon, double* vlat, cudaStream_t st);

}  // namespace xgb
