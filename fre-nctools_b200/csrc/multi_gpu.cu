// One process, several GPUs: exchange-grid generation sharded over the devices of a box, in C++ behind the C ABI (the fregrid_b200
// command line's --gpus N).  Replaces the reference's MPI decomposition of setup_conserve_interp: destination row bands per rank
// (fregrid_util.c:489-492, layout {1,npes}), the order-2 gather of every rank's list to every rank "for bitwise reproducing"
// (conserve_interp.c:202-227) and the gather of the pieces to the root for writing (:404-437).
//
// Sharding is by SOURCE cells: the source cells are cut into ngpus * windows_per_gpu contiguous windows of equal candidate-pair
// count (count pass on the first device), dealt round-robin so every device gets polar and mid-latitude pieces alike.  A window
// is a contiguous piece of the reference's serial emission order, so the windows' results in window order ARE the serial list,
// and a source cell's order-2 centroid sums live on the one device that owns the cell: no exchange of list data between devices
// at all.  What is exchanged is one count per window (here: through host memory, the devices being driven by threads of one
// process), which gives every piece its offset in the caller's host arrays; the devices then copy their pieces there
// concurrently.  The first device holds the whole source mosaic (it runs the count pass); the others receive only the vertex rows
// of their own windows (xgb_plan_set_src_sharded).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

#include "../../include/xgrid_b200.h"
#include "xgrid_plan.h"

namespace {

struct Worker {
  int device = 0;
  xgb_plan* plan = nullptr;
  bool owns_plan = false;
  std::vector<long long> wb, we;       // this device's windows (in global window order)
  std::vector<int> wid;                // their global window numbers
  std::vector<long long> counts;       // exchange cells per window
  long long nx = 0;
  std::string error;
};

int set_dst(xgb_plan* p, const xgb_dst_spec* d)
{
  if (d->by_size) return xgb_plan_set_dst_latlon(p, d->nlon, d->nlat, d->lonbegin, d->lonend, d->latbegin, d->latend);
  return xgb_plan_set_dst(p, d->nx, d->ny, d->lon, d->lat, 0);
}

}  // namespace

extern "C" void xgb_host_xgrid_free(xgb_host_xgrid* x)
{
  if (!x) return;
  void* ptrs[] = {x->t_in, x->i_in, x->j_in, x->i_out, x->j_out, x->area, x->di, x->dj, x->xgrid_clon, x->xgrid_clat};
  for (void* q : ptrs) if (q) cudaFreeHost(q);
  memset(x, 0, sizeof(*x));
}

extern "C" int xgb_generate_multi_gpu(int ngpus, const int* devices, unsigned int opcode, const xgb_dst_spec* dst, int ntiles,
                                      const int* nx, const int* ny, const double* lon, const double* lat, const double* mask,
                                      int windows_per_gpu, xgb_host_xgrid* out)
{
  if (ngpus <= 0 || !dst || ntiles <= 0 || !nx || !ny || !lon || !lat || !out) { xgb_set_error("xgb_generate_multi_gpu: bad arguments"); return 1; }
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)) || (opcode & XGB_GREAT_CIRCLE)) {
    xgb_set_error("xgb_generate_multi_gpu: needs CONSERVE_ORDER1/2 (the great-circle generator runs on one device)");
    return 1;
  }
  memset(out, 0, sizeof(*out));
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  if (windows_per_gpu <= 0) windows_per_gpu = (ngpus > 1) ? 8 : 1;
  if (windows_per_gpu > 64) windows_per_gpu = 64;
  const int nwin = ngpus * windows_per_gpu;
  std::vector<Worker> w(ngpus);
  for (int g = 0; g < ngpus; ++g) w[g].device = devices ? devices[g] : g;

  // ---- the first device: whole mosaic, count pass, window bounds
  w[0].plan = xgb_plan_create(w[0].device);
  if (!w[0].plan) return 1;
  w[0].owns_plan = true;
  auto cleanup = [&]() { for (auto& k : w) if (k.plan && k.owns_plan) { xgb_plan_destroy(k.plan); k.plan = nullptr; } };
  std::vector<long long> bounds(nwin + 1);
  if (set_dst(w[0].plan, dst) || xgb_plan_set_src(w[0].plan, ntiles, nx, ny, lon, lat, mask, 0) ||
      xgb_plan_partition(w[0].plan, nwin, bounds.data())) {
    cleanup();
    return 1;
  }
  for (int k = 0; k < nwin; ++k) {
    Worker& o = w[k % ngpus];
    o.wb.push_back(bounds[k]); o.we.push_back(bounds[k + 1]); o.wid.push_back(k);
  }

  // ---- every device generates its windows (one host thread per device; the library keeps its error text per thread)
  auto generate = [&](int g) {
    Worker& me = w[g];
    const int n = (int)me.wb.size();
    if (g == 0) {
      if (xgb_plan_set_src_windows(me.plan, n, me.wb.data(), me.we.data())) { me.error = xgb_last_error(); return; }
    } else {
      me.plan = xgb_plan_create(me.device);
      if (!me.plan) { me.error = xgb_last_error(); return; }
      me.owns_plan = true;
      if (set_dst(me.plan, dst) || xgb_plan_set_src_sharded(me.plan, ntiles, nx, ny, lon, lat, mask, n, me.wb.data(), me.we.data())) {
        me.error = xgb_last_error();
        return;
      }
    }
    me.nx = xgb_plan_generate(me.plan, opcode);
    if (me.nx < 0) { me.error = xgb_last_error(); return; }
    me.counts.resize(n);
    if (xgb_plan_window_counts(me.plan, me.counts.data())) me.error = xgb_last_error();
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < ngpus; ++g) th.emplace_back(generate, g);
    generate(0);
    for (auto& t : th) t.join();
  }
  for (auto& k : w)
    if (!k.error.empty()) { xgb_set_error("xgb_generate_multi_gpu (device %d): %s", k.device, k.error.c_str()); cleanup(); return 1; }

  // ---- the one exchange: per-window counts -> offset of every window in the serial list
  std::vector<long long> woff(nwin + 1, 0);
  for (int g = 0; g < ngpus; ++g)
    for (size_t q = 0; q < w[g].wid.size(); ++q) woff[w[g].wid[q] + 1] = w[g].counts[q];
  for (int k = 0; k < nwin; ++k) woff[k + 1] += woff[k];
  const long long total = woff[nwin];
  out->nxgrid = total;
  const size_t cap = (size_t)(total > 0 ? total : 1);
  bool ok = true;
  auto halloc = [&](size_t bytes) { void* q = nullptr; if (cudaMallocHost(&q, bytes) != cudaSuccess) { ok = false; q = nullptr; } return q; };
  cudaSetDevice(w[0].device);
  out->t_in = (int*)halloc(cap * sizeof(int)); out->i_in = (int*)halloc(cap * sizeof(int)); out->j_in = (int*)halloc(cap * sizeof(int));
  out->i_out = (int*)halloc(cap * sizeof(int)); out->j_out = (int*)halloc(cap * sizeof(int)); out->area = (double*)halloc(cap * sizeof(double));
  if (order == 2) {
    out->di = (double*)halloc(cap * sizeof(double)); out->dj = (double*)halloc(cap * sizeof(double));
    out->xgrid_clon = (double*)halloc(cap * sizeof(double)); out->xgrid_clat = (double*)halloc(cap * sizeof(double));
  }
  if (!ok) { xgb_set_error("xgb_generate_multi_gpu: cannot allocate pinned host memory for %lld exchange cells", total); xgb_host_xgrid_free(out); cleanup(); return 1; }

  // ---- every device copies its pieces to their places (concurrently)
  auto gather = [&](int g) {
    Worker& me = w[g];
    if (cudaSetDevice(me.device) != cudaSuccess) { me.error = "cudaSetDevice failed"; return; }
    xgb_xgrid_view v;
    if (xgb_plan_result_device(me.plan, &v)) { me.error = xgb_last_error(); return; }
    cudaStream_t st = (cudaStream_t)xgb_plan_stream(me.plan);
    long long local = 0;
    for (size_t q = 0; q < me.wid.size(); ++q) {
      const size_t n = (size_t)me.counts[q], go = (size_t)woff[me.wid[q]];
      struct { void* d; const void* s; size_t esz; } cp[] = {
          {out->t_in, v.t_in, 4}, {out->i_in, v.i_in, 4}, {out->j_in, v.j_in, 4}, {out->i_out, v.i_out, 4}, {out->j_out, v.j_out, 4},
          {out->area, v.area, 8}, {out->di, v.di, 8}, {out->dj, v.dj, 8}, {out->xgrid_clon, v.xgrid_clon, 8}, {out->xgrid_clat, v.xgrid_clat, 8}};
      for (auto& c : cp)
        if (c.d && c.s && n)
          cudaMemcpyAsync((char*)c.d + go * c.esz, (const char*)c.s + (size_t)local * c.esz, n * c.esz, cudaMemcpyDeviceToHost, st);
      local += (long long)n;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) me.error = cudaGetErrorString(cudaGetLastError());
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < ngpus; ++g) th.emplace_back(gather, g);
    gather(0);
    for (auto& t : th) t.join();
  }
  for (auto& k : w)
    if (!k.error.empty()) { xgb_set_error("xgb_generate_multi_gpu (device %d): %s", k.device, k.error.c_str()); xgb_host_xgrid_free(out); cleanup(); return 1; }
  cleanup();
  return 0;
}
