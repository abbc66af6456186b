// Reference-signature L3 entry points (tools/fregrid/conserve_interp.h): setup_conserve_interp (conserve_interp.c:42)
// and do_scalar_conserve_interp (:507) over the reference's own Grid_config / Interp_config / Field_config structs, on top
// of the plan interface.  Error behaviour is the reference's: "FATAL Error: ..." on stderr and exit(1).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/xgrid_b200.h"
#include "fregrid_abi.h"

#define XGB_READ 256u            /* globals.h:54-60 */
#define XGB_WRITE 512u
#define XGB_TARGET 16u
#define XGB_CELL_METHODS_SUM 1

[[noreturn]] static void die(const char* msg)
{
  fprintf(stderr, "FATAL Error: %s\n", msg);          // mpp_error, mpp.c:290-298
  exit(1);
}

static xgb_plan* g_plan = nullptr;                      // one process-wide plan, like the reference's global state
static const void* g_csr_key = nullptr;                 // interp[m].i_in the apply CSR was built from
static size_t g_csr_n = 0;
static int g_csr_order = 0;                            // ... and whether it carries di/dj (order 2)

static xgb_plan* plan()
{
  if (!g_plan) {
    const char* env = getenv("XGB_DEVICE");
    g_plan = xgb_plan_create(env ? atoi(env) : 0);
    if (!g_plan) die(xgb_last_error());
  }
  return g_plan;
}

static void concat_grids(int ntiles, const xgb_Grid_config* g, std::vector<int>& nx, std::vector<int>& ny,
                         std::vector<double>& lon, std::vector<double>& lat)
{
  for (int m = 0; m < ntiles; ++m) {
    nx.push_back(g[m].nx); ny.push_back(g[m].ny);
    const size_t nv = (size_t)(g[m].nx + 1) * (g[m].ny + 1);
    lon.insert(lon.end(), g[m].lonc, g[m].lonc + nv);
    lat.insert(lat.end(), g[m].latc, g[m].latc + nv);
  }
}

extern "C" void setup_conserve_interp(int ntiles_in, const void* grid_in_v, int ntiles_out, void* grid_out_v, void* interp_v,
                                      unsigned int opcode)
{
  const xgb_Grid_config* grid_in = (const xgb_Grid_config*)grid_in_v;
  xgb_Grid_config* grid_out = (xgb_Grid_config*)grid_out_v;
  xgb_Interp_config* interp = (xgb_Interp_config*)interp_v;
  const bool o2 = (opcode & XGB_CONSERVE_ORDER2) != 0;
  if (opcode & XGB_READ) {
    // conserve_interp.c:62-125: every output tile whose remap file exists is read instead of computed
    for (int n = 0; n < ntiles_out; ++n) {
      if (!interp[n].file_exist) continue;
      const long long nall = xgb_remap_size(interp[n].remap_file);
      if (nall < 0) die(xgb_last_error());
      const size_t k = (size_t)(nall > 0 ? nall : 1);
      std::vector<int> t(k), i1(k), j1(k), i2(k), j2(k);
      std::vector<double> a(k), di, dj;
      if (o2) { di.resize(k); dj.resize(k); }
      if (xgb_remap_read(interp[n].remap_file, o2 ? 2 : 1, nall, t.data(), i1.data(), j1.data(), i2.data(), j2.data(), a.data(),
                         o2 ? di.data() : nullptr, o2 ? dj.data() : nullptr))
        die(xgb_last_error());
      // keep the cells of this process's part of the output tile (:93-98), indices relative to it (:107-115)
      const xgb_Grid_config& g = grid_out[n];
      std::vector<size_t> ind;
      for (size_t q = 0; q < (size_t)nall; ++q)
        if (i2[q] <= g.iec && i2[q] >= g.isc && j2[q] <= g.jec && j2[q] >= g.jsc) ind.push_back(q);
      const size_t m = ind.size();
      interp[n].nxgrid = m;
      const size_t mm = m ? m : 1;
      interp[n].i_in = (int*)malloc(mm * sizeof(int));   interp[n].j_in = (int*)malloc(mm * sizeof(int));
      interp[n].i_out = (int*)malloc(mm * sizeof(int));  interp[n].j_out = (int*)malloc(mm * sizeof(int));
      interp[n].t_in = (int*)malloc(mm * sizeof(int));   interp[n].area = (double*)malloc(mm * sizeof(double));
      if (o2) { interp[n].di_in = (double*)malloc(mm * sizeof(double)); interp[n].dj_in = (double*)malloc(mm * sizeof(double)); }
      for (size_t q = 0; q < m; ++q) {
        const size_t s = ind[q];
        interp[n].i_in[q] = i1[s]; interp[n].j_in[q] = j1[s]; interp[n].t_in[q] = t[s];
        interp[n].i_out[q] = i2[s] - g.isc; interp[n].j_out[q] = j2[s] - g.jsc;
        interp[n].area[q] = a[s];
        if (o2) { interp[n].di_in[q] = di[s]; interp[n].dj_in[q] = dj[s]; }
      }
    }
    g_csr_key = nullptr;
    printf("NOTE: Finish reading index and weight for conservative interpolation from file.\n");   // :125
    return;
  }
  if (!(opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2)))
    die("conserve_interp: interp_method should be CONSERVE_ORDER1 or CONSERVE_ORDER2");     // conserve_interp.c:230
  xgb_plan* p = plan();
  std::vector<int> nx, ny;
  std::vector<double> lon, lat;
  concat_grids(ntiles_in, grid_in, nx, ny, lon, lat);
  const unsigned op = opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2 | XGB_GREAT_CIRCLE);
  if (xgb_plan_set_src(p, ntiles_in, nx.data(), ny.data(), lon.data(), lat.data(), nullptr, 0)) die(xgb_last_error());
  // Order 2 with several output tiles: the reference sums every output tile's exchange cells per source cell before the
  // AREA_RATIO test and the centroid subtraction (conserve_interp.c:204-221, :319-358), so the distances can only be
  // finished after the last tile; the per-cell sums stay on the device across the generate calls.
  const bool multi = o2 && ntiles_out > 1 && !(opcode & XGB_GREAT_CIRCLE);
  std::vector<std::vector<double>> rawlon(multi ? ntiles_out : 0), rawlat(multi ? ntiles_out : 0);
  if (multi && xgb_plan_order2_begin(p)) die(xgb_last_error());
  for (int n = 0; n < ntiles_out; ++n) {
    if (xgb_plan_set_dst(p, grid_out[n].nxc, grid_out[n].nyc, grid_out[n].lonc, grid_out[n].latc, 0)) die(xgb_last_error());
    const long long nxg = xgb_plan_generate(p, op);
    if (nxg < 0) die(xgb_last_error());
    interp[n].nxgrid = (size_t)nxg;
    if (nxg == 0) continue;                                                      // the reference allocates nothing either
    const size_t k = (size_t)nxg;
    interp[n].i_in = (int*)malloc(k * sizeof(int));   interp[n].j_in = (int*)malloc(k * sizeof(int));
    interp[n].i_out = (int*)malloc(k * sizeof(int));  interp[n].j_out = (int*)malloc(k * sizeof(int));
    interp[n].t_in = (int*)malloc(k * sizeof(int));   interp[n].area = (double*)malloc(k * sizeof(double));
    if (o2) { interp[n].di_in = (double*)malloc(k * sizeof(double)); interp[n].dj_in = (double*)malloc(k * sizeof(double)); }
    if (xgb_plan_result_host(p, interp[n].t_in, interp[n].i_in, interp[n].j_in, interp[n].i_out, interp[n].j_out, interp[n].area,
                             (o2 && !multi) ? interp[n].di_in : nullptr, (o2 && !multi) ? interp[n].dj_in : nullptr))
      die(xgb_last_error());
    if (multi) {
      rawlon[n].resize(k); rawlat[n].resize(k);
      if (xgb_plan_result_centroids_host(p, rawlon[n].data(), rawlat[n].data())) die(xgb_last_error());
    }
  }
  if (multi) {
    if (xgb_plan_order2_end(p)) die(xgb_last_error());
    for (int n = 0; n < ntiles_out; ++n)
      if (interp[n].nxgrid &&
          xgb_plan_order2_distance(p, (long long)interp[n].nxgrid, interp[n].t_in, interp[n].i_in, interp[n].j_in, interp[n].area,
                                   rawlon[n].data(), rawlat[n].data(), interp[n].di_in, interp[n].dj_in))
        die(xgb_last_error());
    xgb_plan_order2_reset(p);
  }
  if (opcode & XGB_WRITE)                                                        // conserve_interp.c:368-443
    for (int n = 0; n < ntiles_out; ++n) {
      if (interp[n].nxgrid == 0) continue;
      if (xgb_remap_write(interp[n].remap_file, o2 ? 2 : 1, (long long)interp[n].nxgrid, interp[n].t_in, interp[n].i_in, interp[n].j_in,
                          interp[n].i_out, interp[n].j_out, grid_out[n].isc, grid_out[n].jsc, interp[n].area,
                          o2 ? interp[n].di_in : nullptr, o2 ? interp[n].dj_in : nullptr))
        die(xgb_last_error());
    }
  g_csr_key = nullptr;
  if (opcode & XGB_GREAT_CIRCLE) return;
  printf("NOTE: done calculating index and weight for conservative interpolation\n");   // conserve_interp.c:446
}

extern "C" void do_scalar_conserve_interp(void* interp_v, int varid, int ntiles_in, const void* grid_in_v, int ntiles_out,
                                          const void* grid_out_v, const void* field_in_v, void* field_out_v, unsigned int opcode, int nz)
{
  xgb_Interp_config* interp = (xgb_Interp_config*)interp_v;
  const xgb_Grid_config* grid_in = (const xgb_Grid_config*)grid_in_v;
  const xgb_Grid_config* grid_out = (const xgb_Grid_config*)grid_out_v;
  const xgb_Field_config* field_in = (const xgb_Field_config*)field_in_v;
  xgb_Field_config* field_out = (xgb_Field_config*)field_out_v;
  const xgb_Var_config& v = field_in->var[varid];
  const int method = v.interp_method;
  const int order = (method == (int)XGB_CONSERVE_ORDER2) ? 2 : 1;
  const bool has_missing = v.has_missing != 0;
  if (nz > 1 && has_missing) die("conserve_interp: has_missing should be false when nz > 1");                  // :544
  if (nz > 1 && v.cell_measures) die("conserve_interp: cell_measures should be false when nz > 1");            // :545
  if (nz > 1 && v.cell_methods == XGB_CELL_METHODS_SUM) die("conserve_interp: cell_methods should not be sum when nz > 1");
  const bool use_weight = grid_in[0].weight_exist != 0;                                            // :535
  const bool use_sum = v.cell_methods == XGB_CELL_METHODS_SUM, use_meas = v.cell_measures != 0 && !use_sum;
  const bool use_target = (opcode & XGB_TARGET) && !v.use_volume;                                  // :538-539
  xgb_plan* p = plan();
  std::vector<int> nx, ny;
  size_t ncell = 0, nhalo = 0;
  for (int m = 0; m < ntiles_in; ++m) {
    nx.push_back(grid_in[m].nx); ny.push_back(grid_in[m].ny);
    ncell += (size_t)grid_in[m].nx * grid_in[m].ny;
    nhalo += (size_t)(grid_in[m].nx + 2) * (grid_in[m].ny + 2);
  }
  const size_t per = (order == 2) ? nhalo : ncell;
  std::vector<double> data(per * nz), gx, gy;
  std::vector<int> gm;
  if (order == 2) { gx.resize(ncell * nz); gy.resize(ncell * nz); gm.assign(ncell * nz, 0); }
  // Field_config holds one array per tile with the nz levels inside; the batched layout is level-major over all tiles
  size_t off = 0, offc = 0;
  for (int m = 0; m < ntiles_in; ++m) {
    const size_t nt = (order == 2) ? (size_t)(grid_in[m].nx + 2) * (grid_in[m].ny + 2) : (size_t)grid_in[m].nx * grid_in[m].ny;
    const size_t nc = (size_t)grid_in[m].nx * grid_in[m].ny;
    for (int k = 0; k < nz; ++k) {
      memcpy(&data[k * per + off], field_in[m].data + k * nt, nt * sizeof(double));
      if (order == 2) {
        memcpy(&gx[k * ncell + offc], field_in[m].grad_x + k * nc, nc * sizeof(double));
        memcpy(&gy[k * ncell + offc], field_in[m].grad_y + k * nc, nc * sizeof(double));
        if (field_in[m].grad_mask) memcpy(&gm[k * ncell + offc], field_in[m].grad_mask + k * nc, nc * sizeof(int));
      }
    }
    off += nt; offc += nc;
  }
  std::vector<double> wgt, carea, farea;
  if (use_weight || use_sum || use_meas) {
    for (int m = 0; m < ntiles_in; ++m) {
      const size_t nc = (size_t)grid_in[m].nx * grid_in[m].ny;
      if (use_weight) wgt.insert(wgt.end(), grid_in[m].weight, grid_in[m].weight + nc);
      if (use_sum || use_meas) carea.insert(carea.end(), grid_in[m].cell_area, grid_in[m].cell_area + nc);
      if (use_meas) farea.insert(farea.end(), field_in[m].area, field_in[m].area + nc);
    }
  }
  for (int m = 0; m < ntiles_out; ++m) {
    const int nx2 = grid_out[m].nxc, ny2 = grid_out[m].nyc;
    if (interp[m].nxgrid == 0) {
      const double miss = has_missing ? v.missing : -1.e20;
      for (size_t q = 0; q < (size_t)nx2 * ny2 * nz; ++q) field_out[m].data[q] = miss;
      continue;
    }
    if (g_csr_key != (const void*)interp[m].i_in || g_csr_n != interp[m].nxgrid || g_csr_order != order) {
      if (xgb_plan_set_xgrid(p, ntiles_in, nx.data(), ny.data(), nx2, ny2, (long long)interp[m].nxgrid, interp[m].t_in, interp[m].i_in,
                             interp[m].j_in, interp[m].i_out, interp[m].j_out, interp[m].area,
                             order == 2 ? interp[m].di_in : nullptr, order == 2 ? interp[m].dj_in : nullptr, 0))
        die(xgb_last_error());
      g_csr_key = interp[m].i_in; g_csr_n = interp[m].nxgrid; g_csr_order = order;
    }
    if (xgb_plan_apply_options(p, use_sum ? 1 : 0, use_weight ? wgt.data() : nullptr, (use_sum || use_meas) ? carea.data() : nullptr,
                               use_meas ? farea.data() : nullptr, v.area_missing, use_target ? 1 : 0,
                               use_target ? grid_out[m].cell_area : nullptr, 0))
      die(xgb_last_error());
    unsigned op = (order == 2) ? XGB_CONSERVE_ORDER2 : XGB_CONSERVE_ORDER1;
    if (order == 2 && (opcode & XGB_MONOTONIC)) op |= XGB_MONOTONIC;
    if (xgb_plan_apply(p, op, nz, data.data(), order == 2 ? gx.data() : nullptr, order == 2 ? gy.data() : nullptr,
                       order == 2 ? gm.data() : nullptr, has_missing, v.missing, field_out[m].data, 0))
      die(xgb_last_error());
  }
}

// layout self-description for the ABI test: sizeof the structs and offsetof the members this file reads
extern "C" int xgb_abi_layout(size_t* out, int cap)
{
  const size_t v[] = {
      sizeof(xgb_Var_config), offsetof(xgb_Var_config, missing), offsetof(xgb_Var_config, has_missing), offsetof(xgb_Var_config, interp_method),
      offsetof(xgb_Var_config, cell_measures), offsetof(xgb_Var_config, cell_methods), offsetof(xgb_Var_config, use_volume),
      sizeof(xgb_Field_config), offsetof(xgb_Field_config, data), offsetof(xgb_Field_config, grad_x), offsetof(xgb_Field_config, grad_y),
      offsetof(xgb_Field_config, grad_mask), offsetof(xgb_Field_config, var),
      sizeof(xgb_Interp_config), offsetof(xgb_Interp_config, nxgrid), offsetof(xgb_Interp_config, i_in), offsetof(xgb_Interp_config, t_in),
      offsetof(xgb_Interp_config, di_in), offsetof(xgb_Interp_config, area), offsetof(xgb_Interp_config, file_exist),
      sizeof(xgb_Grid_config), offsetof(xgb_Grid_config, nx), offsetof(xgb_Grid_config, ny), offsetof(xgb_Grid_config, nxc),
      offsetof(xgb_Grid_config, nyc), offsetof(xgb_Grid_config, lonc), offsetof(xgb_Grid_config, latc), offsetof(xgb_Grid_config, cell_area),
      offsetof(xgb_Grid_config, weight_exist), offsetof(xgb_Grid_config, domain)};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  for (int k = 0; k < n && k < cap; ++k) out[k] = v[k];
  return n;
}
