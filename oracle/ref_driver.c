/* TEST INFRASTRUCTURE — not product code.
 *
 * Thin flat-array entry points over the UNMODIFIED reference sources (compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libfrenc_ref.so).  Python (ctypes)
 * cannot conveniently build the reference's Grid_config/Interp_config/Field_config structs
 * (tools/libfrencutils/globals.h:66-222), so this file builds them in C and calls the real
 * functions.  Nothing here re-implements reference math.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <netcdf.h>
#include "constant.h"
#include "mpp.h"
#include "mpp_domain.h"
#include "globals.h"
#include "mosaic_util.h"
#include "create_xgrid.h"
#include "gradient_c2l.h"
#include "create_hgrid.h"
#include "conserve_interp.h"

void cell_center(int ni, int nj, const double *lonc, const double *latc, double *lont, double *latt);

/* serial mpp shim state (npes = 1, pe = 0): fregrid's main does this at fregrid.c:414-415 */
static void ref_init_once(void)
{
  static int done = 0;
  if(!done) {
    int argc = 0; char **argv = NULL;
    mpp_init(&argc, &argv);
    mpp_domain_init();
    done = 1;
  }
}

/* Cubed-sphere C<ni> "gnomonic_ed" supergrid exactly as `make_hgrid --grid_type gnomonic_ed
 * --nlon 2*ni` produces it (make_hgrid.c:1155), then subsampled/converted to radians exactly as
 * fregrid's get_input_grid does (fregrid_util.c:227-241).
 *   lonc/latc : 6*(ni+1)*(ni+1)  cell corners, radians
 *   lont/latt : 6*ni*ni          cell centres, radians (may be NULL)
 * Returns 0 on success. */
int ref_cubed_sphere_grid(int ni, double *lonc, double *latc, double *lont, double *latt)
{
  int nlon[6], nlat[6], n, i, j;
  int parent_tile[MAX_NESTS], refine_ratio[MAX_NESTS], is_n[MAX_NESTS], ie_n[MAX_NESTS], js_n[MAX_NESTS], je_n[MAX_NESTS];
  long nx = 2L*ni, nxp = nx+1;
  size_t size1 = (size_t)nxp*nxp*6;
  double *x, *y, *area;
  FILE *saved;

  for(n=0; n<6; n++) { nlon[n] = (int)nx; nlat[n] = (int)nx; }
  memset(parent_tile, 0, sizeof(parent_tile));
  memset(refine_ratio, 0, sizeof(refine_ratio));
  memset(is_n, 0, sizeof(is_n)); memset(ie_n, 0, sizeof(ie_n));
  memset(js_n, 0, sizeof(js_n)); memset(je_n, 0, sizeof(je_n));
  x    = (double *)malloc(size1*sizeof(double));
  y    = (double *)malloc(size1*sizeof(double));
  area = (double *)malloc((size_t)nx*nx*6*sizeof(double));
  if(!x || !y || !area) return 1;

  /* the generator is chatty on stderr ([INFO] lines); silence it for the duration */
  saved = stderr;
  stderr = fopen("/dev/null", "w");
  /* output_length_angle = 0: dx/dy/angle are not needed by fregrid's conservative path */
  create_gnomonic_cubic_grid("gnomonic_ed", nlon, nlat, x, y, NULL, NULL, area, NULL, NULL,
                             18.0, 0, 0, 1.0, 0.0, 0.0, 0, parent_tile, refine_ratio,
                             is_n, ie_n, js_n, je_n, 0, 0);
  if(stderr) fclose(stderr);
  stderr = saved;

  for(n=0; n<6; n++) {
    const double *xs = x + (size_t)n*nxp*nxp, *ys = y + (size_t)n*nxp*nxp;
    for(j=0; j<=ni; j++) for(i=0; i<=ni; i++) {
      size_t ind1 = (size_t)n*(ni+1)*(ni+1) + (size_t)j*(ni+1) + i;
      size_t ind2 = (size_t)2*j*nxp + 2*i;
      lonc[ind1] = xs[ind2]*D2R;
      latc[ind1] = ys[ind2]*D2R;
    }
    if(lont && latt) {
      for(j=0; j<ni; j++) for(i=0; i<ni; i++) {
        size_t ind1 = (size_t)n*ni*ni + (size_t)j*ni + i;
        size_t ind2 = (size_t)(2*j+1)*nxp + 2*i+1;
        lont[ind1] = xs[ind2]*D2R;
        latt[ind1] = ys[ind2]*D2R;
      }
    }
  }
  free(x); free(y); free(area);
  return 0;
}

/* Tripolar ocean grid as `make_hgrid --grid_type tripolar_grid --nxbnd 2 --nybnd 2 --xbnd x0,x1
 * --ybnd y0,y1 --nlon nlon --nlat nlat --lat_join lat_join` (make_hgrid.c:1126), subsampled as
 * fregrid_util.c:227-232.  nlon/nlat are SUPERGRID sizes.  lonc/latc: (nlon/2+1)*(nlat/2+1). */
int ref_tripolar_grid(int nlon_s, int nlat_s, double x0, double x1, double y0, double y1, double lat_join,
                      double *lonc, double *latc)
{
  int nxbnds = 2, nybnds = 2, nlon[1], nlat[1], isc, iec, jsc, jec, i, j;
  double xbnds[2], ybnds[2], dxb[2] = {0,0}, dyb[2] = {0,0};
  int nx = nlon_s, ny = nlat_s, nxp = nx+1, nyp = ny+1, ni = nx/2, nj = ny/2;
  double *x, *y, *dx, *dy, *area, *angle;
  FILE *saved;

  xbnds[0] = x0; xbnds[1] = x1; ybnds[0] = y0; ybnds[1] = y1;
  nlon[0] = nlon_s; nlat[0] = nlat_s;
  x     = (double *)malloc((size_t)nxp*nyp*sizeof(double));
  y     = (double *)malloc((size_t)nxp*nyp*sizeof(double));
  dx    = (double *)malloc((size_t)nxp*(nyp+1)*sizeof(double));
  dy    = (double *)malloc((size_t)(nxp+1)*nyp*sizeof(double));
  area  = (double *)malloc((size_t)nxp*nyp*sizeof(double));
  angle = (double *)malloc((size_t)nxp*nyp*sizeof(double));
  isc = 0; iec = nx-1; jsc = 0; jec = ny-1;
  saved = stdout;
  stdout = fopen("/dev/null", "w");
  create_tripolar_grid(&nxbnds, &nybnds, xbnds, ybnds, nlon, nlat, dxb, dyb, 0, &lat_join,
                       &isc, &iec, &jsc, &jec, x, y, dx, dy, area, angle, "none", 0, 0);
  if(stdout) fclose(stdout);
  stdout = saved;
  for(j=0; j<=nj; j++) for(i=0; i<=ni; i++) {
    lonc[(size_t)j*(ni+1)+i] = x[(size_t)2*j*nxp+2*i]*D2R;
    latc[(size_t)j*(ni+1)+i] = y[(size_t)2*j*nxp+2*i]*D2R;
  }
  free(x); free(y); free(dx); free(dy); free(area); free(angle);
  return 0;
}

/* ---- conservative setup + apply through the reference's own L3 code ------------------------ */

typedef struct {
  int ntiles_in, ntiles_out;
  Grid_config *gin, *gout;
  Interp_config *interp;
  unsigned int opcode;
} RefRegrid;

/* Build Grid_config arrays the way fregrid's main does for the conservative path
 * (get_input_grid fregrid_util.c:157-360 without file I/O, get_output_grid_by_size :564-659,
 * get_input_output_cell_area :363-408) and run the real setup_conserve_interp (conserve_interp.c:42).
 *   lonc_in/latc_in: concatenated tiles, each (nx_in[n]+1)*(ny_in[n]+1), radians
 *   lont_in/latt_in: concatenated tiles, each nx*ny cell centres (only needed for order 2 apply; may be NULL)
 *   lonc_out/latc_out: single output tile (nx_out+1)*(ny_out+1)
 *   jsc_out/jec_out: destination row window [jsc,jec] owned by this "rank" (fregrid_util.c:592-603);
 *                    pass 0, ny_out-1 for the serial case.
 */
/* the next ref_regrid_setup runs setup_conserve_interp with WRITE (mode 1) or READ (mode 2) on interp[0].remap_file = name */
static unsigned int ref_remap_mode = 0;
static char ref_remap_name[STRING];
void ref_next_setup_remap(const char *name, int mode)
{
  strncpy(ref_remap_name, name, STRING-1);
  ref_remap_mode = (mode == 1) ? WRITE : (mode == 2) ? READ : 0;
}

RefRegrid *ref_regrid_setup(int ntiles_in, const int *nx_in, const int *ny_in,
                            const double *lonc_in, const double *latc_in,
                            const double *lont_in, const double *latt_in,
                            int nx_out, int ny_out, const double *lonc_out, const double *latc_out,
                            int jsc_out, int jec_out, unsigned int opcode)
{
  RefRegrid *r = (RefRegrid *)calloc(1, sizeof(RefRegrid));
  size_t offc = 0, offt = 0;
  int n, i, j;

  ref_init_once();

  r->ntiles_in = ntiles_in; r->ntiles_out = 1; r->opcode = opcode;
  r->gin    = (Grid_config *)calloc(ntiles_in, sizeof(Grid_config));
  r->gout   = (Grid_config *)calloc(1, sizeof(Grid_config));
  r->interp = (Interp_config *)calloc(1, sizeof(Interp_config));

  for(n=0; n<ntiles_in; n++) {
    int nx = nx_in[n], ny = ny_in[n];
    Grid_config *g = r->gin + n;
    g->halo = 0; g->nx = nx; g->ny = ny; g->nxc = nx; g->nyc = ny;
    g->lonc = (double *)malloc((size_t)(nx+1)*(ny+1)*sizeof(double));
    g->latc = (double *)malloc((size_t)(nx+1)*(ny+1)*sizeof(double));
    memcpy(g->lonc, lonc_in+offc, (size_t)(nx+1)*(ny+1)*sizeof(double));
    memcpy(g->latc, latc_in+offc, (size_t)(nx+1)*(ny+1)*sizeof(double));
    offc += (size_t)(nx+1)*(ny+1);
    if(lont_in && latt_in) {
      g->lont = (double *)calloc((size_t)(nx+2)*(ny+2), sizeof(double));
      g->latt = (double *)calloc((size_t)(nx+2)*(ny+2), sizeof(double));
      for(j=0; j<ny; j++) for(i=0; i<nx; i++) {
        g->lont[(size_t)(j+1)*(nx+2)+i+1] = lont_in[offt+(size_t)j*nx+i];
        g->latt[(size_t)(j+1)*(nx+2)+i+1] = latt_in[offt+(size_t)j*nx+i];
      }
      offt += (size_t)nx*ny;
    }
  }
  {
    Grid_config *g = r->gout;
    int nyc = jec_out - jsc_out + 1;
    g->nx = nx_out; g->ny = ny_out; g->nxc = nx_out; g->nyc = nyc;
    g->isc = 0; g->iec = nx_out-1; g->jsc = jsc_out; g->jec = jec_out;
    g->lonc = (double *)malloc((size_t)(nx_out+1)*(nyc+1)*sizeof(double));
    g->latc = (double *)malloc((size_t)(nx_out+1)*(nyc+1)*sizeof(double));
    memcpy(g->lonc, lonc_out+(size_t)jsc_out*(nx_out+1), (size_t)(nx_out+1)*(nyc+1)*sizeof(double));
    memcpy(g->latc, latc_out+(size_t)jsc_out*(nx_out+1), (size_t)(nx_out+1)*(nyc+1)*sizeof(double));
  }
  /* get_input_output_cell_area (fregrid_util.c:363-408), halo == 0 */
  for(n=0; n<ntiles_in; n++) {
    Grid_config *g = r->gin + n;
    g->cell_area = (double *)malloc((size_t)g->nx*g->ny*sizeof(double));
    if(opcode & GREAT_CIRCLE) get_grid_great_circle_area(&g->nx, &g->ny, g->lonc, g->latc, g->cell_area);
    else                      get_grid_area(&g->nx, &g->ny, g->lonc, g->latc, g->cell_area);
  }
  {
    Grid_config *g = r->gout;
    g->cell_area = (double *)malloc((size_t)g->nxc*g->nyc*sizeof(double));
    if(opcode & GREAT_CIRCLE) get_grid_great_circle_area(&g->nxc, &g->nyc, g->lonc, g->latc, g->cell_area);
    else                      get_grid_area(&g->nxc, &g->nyc, g->lonc, g->latc, g->cell_area);
  }
  r->interp[0].file_exist = 0;
  if(ref_remap_mode) {   /* WRITE into / READ from the in-memory file store of shim/io_stubs.c */
    strncpy(r->interp[0].remap_file, ref_remap_name, STRING-1);
    r->interp[0].file_exist = (ref_remap_mode == READ);
    setup_conserve_interp(ntiles_in, r->gin, 1, r->gout, r->interp, (opcode & ~(WRITE|READ|CHECK_CONSERVE)) | ref_remap_mode);
    ref_remap_mode = 0;
    return r;
  }
  setup_conserve_interp(ntiles_in, r->gin, 1, r->gout, r->interp, opcode & ~(WRITE|READ|CHECK_CONSERVE));
  return r;
}

long ref_regrid_nxgrid(const RefRegrid *r) { return (long)r->interp[0].nxgrid; }

/* copy the Interp_config lists out (globals.h:149-163); di/dj may be NULL for order 1 */
void ref_regrid_get(const RefRegrid *r, int *t_in, int *i_in, int *j_in, int *i_out, int *j_out,
                    double *area, double *di, double *dj)
{
  size_t n = r->interp[0].nxgrid;
  memcpy(t_in,  r->interp[0].t_in,  n*sizeof(int));
  memcpy(i_in,  r->interp[0].i_in,  n*sizeof(int));
  memcpy(j_in,  r->interp[0].j_in,  n*sizeof(int));
  memcpy(i_out, r->interp[0].i_out, n*sizeof(int));
  memcpy(j_out, r->interp[0].j_out, n*sizeof(int));
  memcpy(area,  r->interp[0].area,  n*sizeof(double));
  if(di && (r->opcode & CONSERVE_ORDER2)) memcpy(di, r->interp[0].di_in, n*sizeof(double));
  if(dj && (r->opcode & CONSERVE_ORDER2)) memcpy(dj, r->interp[0].dj_in, n*sizeof(double));
}

void ref_regrid_cell_area(const RefRegrid *r, double *area_in /* concatenated */, double *area_out)
{
  size_t off = 0; int n;
  for(n=0; n<r->ntiles_in; n++) {
    size_t sz = (size_t)r->gin[n].nx*r->gin[n].ny;
    if(area_in) memcpy(area_in+off, r->gin[n].cell_area, sz*sizeof(double));
    off += sz;
  }
  if(area_out) memcpy(area_out, r->gout->cell_area, (size_t)r->gout->nxc*r->gout->nyc*sizeof(double));
}

/* One call of the reference's do_scalar_conserve_interp (conserve_interp.c:507) for one 2-D (or nz-level)
 * field.  data_in: concatenated tiles; order 1: nx*ny*nz each; order 2: (nx+2)*(ny+2)*nz each (halo 1).
 * grad_x/grad_y: concatenated nx*ny*nz, grad_mask: nx*ny ints (order 2 only).
 * data_out: nxc*nyc*nz.  */
void ref_regrid_apply(RefRegrid *r, int interp_method, int has_missing, double missing,
                      int cell_methods, int nz, unsigned int extra_opcode,
                      const double *data_in, const double *grad_x, const double *grad_y,
                      const int *grad_mask, double *data_out)
{
  Field_config *fin  = (Field_config *)calloc(r->ntiles_in, sizeof(Field_config));
  Field_config *fout = (Field_config *)calloc(1, sizeof(Field_config));
  Var_config *var = (Var_config *)calloc(1, sizeof(Var_config));
  size_t offd = 0, offg = 0, offm = 0;
  int n, halo = (interp_method == CONSERVE_ORDER2) ? 1 : 0;

  strcpy(var->name, "f");
  var->interp_method = interp_method;
  var->has_missing = has_missing;
  var->missing = missing;
  var->cell_methods = cell_methods;
  var->cell_measures = 0;
  var->use_volume = 0;
  for(n=0; n<r->ntiles_in; n++) {
    size_t nx = r->gin[n].nx, ny = r->gin[n].ny;
    fin[n].var = var;
    fin[n].data = (double *)(data_in + offd);
    offd += (nx+2*halo)*(ny+2*halo)*nz;
    if(halo) {
      fin[n].grad_x = (double *)(grad_x + offg);
      fin[n].grad_y = (double *)(grad_y + offg);
      fin[n].grad_mask = (int *)(grad_mask + offm);
      offg += nx*ny*nz;
      offm += nx*ny;
    }
  }
  fout[0].var = var;
  fout[0].data = data_out;
  do_scalar_conserve_interp(r->interp, 0, r->ntiles_in, r->gin, 1, r->gout, fin, fout,
                            (r->opcode | extra_opcode) & ~CHECK_CONSERVE, nz);
  free(fin); free(fout); free(var);
}

void ref_regrid_free(RefRegrid *r)
{
  int n;
  if(!r) return;
  for(n=0; n<r->ntiles_in; n++) {
    free(r->gin[n].lonc); free(r->gin[n].latc); free(r->gin[n].lont); free(r->gin[n].latt);
    free(r->gin[n].cell_area);
  }
  free(r->gout->lonc); free(r->gout->latc); free(r->gout->cell_area);
  if(r->interp[0].nxgrid > 0) {
    free(r->interp[0].i_in); free(r->interp[0].j_in); free(r->interp[0].i_out); free(r->interp[0].j_out);
    free(r->interp[0].t_in); free(r->interp[0].area);
    if(r->opcode & CONSERVE_ORDER2) { free(r->interp[0].di_in); free(r->interp[0].dj_in); }
  }
  free(r->gin); free(r->gout); free(r->interp); free(r);
}

/* mpp_compute_extent (mpp_domain.c:101) — the row-band decomposition fregrid_parallel uses */
void ref_compute_extent(int npts, int ndivs, int *ibegin, int *iend)
{
  mpp_compute_extent(npts, ndivs, ibegin, iend);
}


/* ---- ABI checks of the product's reference-signature entry points ------------------------------------------ */
#include <stddef.h>
/* sizeof / offsetof of the REAL structs, in the order libxgrid_b200's xgb_abi_layout() reports its mirrors */
int ref_abi_layout(size_t *out, int cap)
{
  const size_t v[] = {
      sizeof(Var_config), offsetof(Var_config, missing), offsetof(Var_config, has_missing), offsetof(Var_config, interp_method),
      offsetof(Var_config, cell_measures), offsetof(Var_config, cell_methods), offsetof(Var_config, use_volume),
      sizeof(Field_config), offsetof(Field_config, data), offsetof(Field_config, grad_x), offsetof(Field_config, grad_y),
      offsetof(Field_config, grad_mask), offsetof(Field_config, var),
      sizeof(Interp_config), offsetof(Interp_config, nxgrid), offsetof(Interp_config, i_in), offsetof(Interp_config, t_in),
      offsetof(Interp_config, di_in), offsetof(Interp_config, area), offsetof(Interp_config, file_exist),
      sizeof(Grid_config), offsetof(Grid_config, nx), offsetof(Grid_config, ny), offsetof(Grid_config, nxc),
      offsetof(Grid_config, nyc), offsetof(Grid_config, lonc), offsetof(Grid_config, latc), offsetof(Grid_config, cell_area),
      offsetof(Grid_config, weight_exist), offsetof(Grid_config, domain)};
  const int n = (int)(sizeof(v)/sizeof(v[0]));
  int k;
  for(k=0; k<n && k<cap; k++) out[k] = v[k];
  return n;
}

typedef void (*setup_fn_t)(int, const Grid_config *, int, Grid_config *, Interp_config *, unsigned int);
typedef void (*apply_fn_t)(Interp_config *, int, int, const Grid_config *, int, const Grid_config *, const Field_config *,
                           Field_config *, unsigned int, int);

/* Re-run the setup of an existing RefRegrid through ANOTHER implementation of setup_conserve_interp (a function pointer
 * taken from libxgrid_b200.so), handing it the real Grid_config structs; returns a new handle sharing the grids. */
RefRegrid *ref_regrid_setup_through(const RefRegrid *r, void *setup_fn)
{
  RefRegrid *q = (RefRegrid *)calloc(1, sizeof(RefRegrid));
  *q = *r;
  q->interp = (Interp_config *)calloc(1, sizeof(Interp_config));
  ((setup_fn_t)setup_fn)(q->ntiles_in, q->gin, 1, q->gout, q->interp, q->opcode & ~(WRITE|READ|CHECK_CONSERVE|LEGACY_CLIP));
  return q;
}

/* the same, with WRITE (mode 1) or READ (mode 2) set and interp[0].remap_file = name: the product writes / reads a real file */
RefRegrid *ref_regrid_setup_through_remap(const RefRegrid *r, void *setup_fn, const char *name, int mode)
{
  RefRegrid *q = (RefRegrid *)calloc(1, sizeof(RefRegrid));
  *q = *r;
  q->interp = (Interp_config *)calloc(1, sizeof(Interp_config));
  strncpy(q->interp[0].remap_file, name, STRING-1);
  q->interp[0].file_exist = (mode == 2);
  ((setup_fn_t)setup_fn)(q->ntiles_in, q->gin, 1, q->gout, q->interp,
                         (q->opcode & ~(WRITE|READ|CHECK_CONSERVE|LEGACY_CLIP)) | (mode == 1 ? WRITE : mode == 2 ? READ : 0));
  return q;
}

/* ref_regrid_apply, but through another implementation of do_scalar_conserve_interp */
void ref_regrid_apply_through(RefRegrid *r, void *apply_fn, int interp_method, int has_missing, double missing,
                              int nz, unsigned int extra_opcode,
                              const double *data_in, const double *grad_x, const double *grad_y,
                              const int *grad_mask, double *data_out)
{
  Field_config *fin  = (Field_config *)calloc(r->ntiles_in, sizeof(Field_config));
  Field_config *fout = (Field_config *)calloc(1, sizeof(Field_config));
  Var_config *var = (Var_config *)calloc(1, sizeof(Var_config));
  size_t offd = 0, offg = 0, offm = 0;
  int n, halo = (interp_method == CONSERVE_ORDER2) ? 1 : 0;
  strcpy(var->name, "f");
  var->interp_method = interp_method;
  var->has_missing = has_missing;
  var->missing = missing;
  for(n=0; n<r->ntiles_in; n++) {
    size_t nx = r->gin[n].nx, ny = r->gin[n].ny;
    fin[n].var = var;
    fin[n].data = (double *)(data_in + offd);
    offd += (nx+2*halo)*(ny+2*halo)*nz;
    if(halo) {
      fin[n].grad_x = (double *)(grad_x + offg);
      fin[n].grad_y = (double *)(grad_y + offg);
      fin[n].grad_mask = (int *)(grad_mask + offm);
      offg += nx*ny*nz;
      offm += nx*ny;
    }
  }
  fout[0].var = var;
  fout[0].data = data_out;
  ((apply_fn_t)apply_fn)(r->interp, 0, r->ntiles_in, r->gin, 1, r->gout, fin, fout,
                         (r->opcode | extra_opcode) & ~(CHECK_CONSERVE|LEGACY_CLIP), nz);
  free(fin); free(fout); free(var);
}


/* do_scalar_conserve_interp with the optional per-cell factors: weight (grid_in[].weight + weight_exist), cell_methods,
 * cell_measures (field_in[].area, area_missing), TARGET (grid_out cell_area is already set by ref_regrid_setup).
 * weight / farea: concatenated over tiles, NULL when unused.  nz == 1. */
void ref_regrid_apply_ex(RefRegrid *r, int interp_method, int has_missing, double missing, int cell_methods,
                         const double *weight, const double *farea, double area_missing, unsigned int extra_opcode,
                         const double *data_in, const double *grad_x, const double *grad_y, const int *grad_mask,
                         double *data_out)
{
  Field_config *fin  = (Field_config *)calloc(r->ntiles_in, sizeof(Field_config));
  Field_config *fout = (Field_config *)calloc(1, sizeof(Field_config));
  Var_config *var = (Var_config *)calloc(1, sizeof(Var_config));
  size_t offd = 0, offc = 0;
  int n, halo = (interp_method == CONSERVE_ORDER2) ? 1 : 0;
  strcpy(var->name, "f");
  var->interp_method = interp_method;
  var->has_missing = has_missing;
  var->missing = missing;
  var->cell_methods = cell_methods;
  var->cell_measures = farea ? 1 : 0;
  var->area_missing = area_missing;
  for(n=0; n<r->ntiles_in; n++) {
    size_t nx = r->gin[n].nx, ny = r->gin[n].ny;
    fin[n].var = var;
    fin[n].data = (double *)(data_in + offd);
    offd += (nx+2*halo)*(ny+2*halo);
    if(halo) {
      fin[n].grad_x = (double *)(grad_x + offc);
      fin[n].grad_y = (double *)(grad_y + offc);
      fin[n].grad_mask = (int *)(grad_mask + offc);
    }
    if(farea) fin[n].area = (double *)(farea + offc);
    r->gin[n].weight = weight ? (double *)(weight + offc) : NULL;
    r->gin[n].weight_exist = weight ? 1 : 0;
    offc += nx*ny;
  }
  fout[0].var = var;
  fout[0].data = data_out;
  do_scalar_conserve_interp(r->interp, 0, r->ntiles_in, r->gin, 1, r->gout, fin, fout,
                            (r->opcode | extra_opcode) & ~CHECK_CONSERVE, 1);
  for(n=0; n<r->ntiles_in; n++) { r->gin[n].weight = NULL; r->gin[n].weight_exist = 0; }
  free(fin); free(fout); free(var);
}


/* ref_regrid_apply_ex through another implementation of do_scalar_conserve_interp (function pointer) */
void ref_regrid_apply_ex_through(RefRegrid *r, void *apply_fn, int interp_method, int has_missing, double missing, int cell_methods,
                                 const double *weight, const double *farea, double area_missing, unsigned int extra_opcode,
                                 const double *data_in, const double *grad_x, const double *grad_y, const int *grad_mask,
                                 double *data_out)
{
  Field_config *fin  = (Field_config *)calloc(r->ntiles_in, sizeof(Field_config));
  Field_config *fout = (Field_config *)calloc(1, sizeof(Field_config));
  Var_config *var = (Var_config *)calloc(1, sizeof(Var_config));
  size_t offd = 0, offc = 0;
  int n, halo = (interp_method == CONSERVE_ORDER2) ? 1 : 0;
  strcpy(var->name, "f");
  var->interp_method = interp_method;
  var->has_missing = has_missing;
  var->missing = missing;
  var->cell_methods = cell_methods;
  var->cell_measures = farea ? 1 : 0;
  var->area_missing = area_missing;
  for(n=0; n<r->ntiles_in; n++) {
    size_t nx = r->gin[n].nx, ny = r->gin[n].ny;
    fin[n].var = var;
    fin[n].data = (double *)(data_in + offd);
    offd += (nx+2*halo)*(ny+2*halo);
    if(halo) {
      fin[n].grad_x = (double *)(grad_x + offc);
      fin[n].grad_y = (double *)(grad_y + offc);
      fin[n].grad_mask = (int *)(grad_mask + offc);
    }
    if(farea) fin[n].area = (double *)(farea + offc);
    r->gin[n].weight = weight ? (double *)(weight + offc) : NULL;
    r->gin[n].weight_exist = weight ? 1 : 0;
    offc += nx*ny;
  }
  fout[0].var = var;
  fout[0].data = data_out;
  ((apply_fn_t)apply_fn)(r->interp, 0, r->ntiles_in, r->gin, 1, r->gout, fin, fout,
                         (r->opcode | extra_opcode) & ~(CHECK_CONSERVE|LEGACY_CLIP), 1);
  for(n=0; n<r->ntiles_in; n++) { r->gin[n].weight = NULL; r->gin[n].weight_exist = 0; }
  free(fin); free(fout); free(var);
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Several OUTPUT tiles (conserve_interp.c:148-227, :319-358: order 2 sums every output tile's exchange cells per source
 * cell before the centroid correction).  Builds Grid_config[ntiles_in] / Grid_config[ntiles_out] and runs
 * setup_conserve_interp — the reference's own when setup_fn is NULL, otherwise the implementation behind the pointer
 * (libxgrid_b200.so's export) on the very same structs.
 *   lonc_out/latc_out: the output tiles' vertex arrays back to back, each (nx_out[n]+1)*(ny_out[n]+1).
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct { int ntiles_in, ntiles_out; unsigned int opcode; Grid_config *gin, *gout; Interp_config *interp; } RefMulti;

RefMulti *ref_multi_setup(int ntiles_in, const int *nx_in, const int *ny_in, const double *lonc_in, const double *latc_in,
                          int ntiles_out, const int *nx_out, const int *ny_out, const double *lonc_out, const double *latc_out,
                          unsigned int opcode, void *setup_fn)
{
  RefMulti *r = (RefMulti *)calloc(1, sizeof(RefMulti));
  size_t off = 0;
  int n;
  ref_init_once();
  r->ntiles_in = ntiles_in; r->ntiles_out = ntiles_out; r->opcode = opcode;
  r->gin = (Grid_config *)calloc(ntiles_in, sizeof(Grid_config));
  r->gout = (Grid_config *)calloc(ntiles_out, sizeof(Grid_config));
  r->interp = (Interp_config *)calloc(ntiles_out, sizeof(Interp_config));
  for(n=0; n<ntiles_in; n++) {
    Grid_config *g = r->gin + n;
    size_t nv = (size_t)(nx_in[n]+1)*(ny_in[n]+1);
    g->nx = g->nxc = nx_in[n]; g->ny = g->nyc = ny_in[n];
    g->lonc = (double *)malloc(nv*sizeof(double)); g->latc = (double *)malloc(nv*sizeof(double));
    memcpy(g->lonc, lonc_in+off, nv*sizeof(double)); memcpy(g->latc, latc_in+off, nv*sizeof(double));
    off += nv;
    g->cell_area = (double *)malloc((size_t)g->nx*g->ny*sizeof(double));
    get_grid_area(&g->nx, &g->ny, g->lonc, g->latc, g->cell_area);
  }
  off = 0;
  for(n=0; n<ntiles_out; n++) {
    Grid_config *g = r->gout + n;
    size_t nv = (size_t)(nx_out[n]+1)*(ny_out[n]+1);
    g->nx = g->nxc = nx_out[n]; g->ny = g->nyc = ny_out[n];
    g->isc = 0; g->iec = nx_out[n]-1; g->jsc = 0; g->jec = ny_out[n]-1;
    g->lonc = (double *)malloc(nv*sizeof(double)); g->latc = (double *)malloc(nv*sizeof(double));
    memcpy(g->lonc, lonc_out+off, nv*sizeof(double)); memcpy(g->latc, latc_out+off, nv*sizeof(double));
    off += nv;
    g->cell_area = (double *)malloc((size_t)g->nx*g->ny*sizeof(double));
    get_grid_area(&g->nx, &g->ny, g->lonc, g->latc, g->cell_area);
  }
  if(setup_fn) ((setup_fn_t)setup_fn)(ntiles_in, r->gin, ntiles_out, r->gout, r->interp, opcode & ~(WRITE|READ|CHECK_CONSERVE|LEGACY_CLIP));
  else setup_conserve_interp(ntiles_in, r->gin, ntiles_out, r->gout, r->interp, opcode & ~(WRITE|READ|CHECK_CONSERVE));
  return r;
}

long ref_multi_nxgrid(const RefMulti *r, int n) { return (long)r->interp[n].nxgrid; }

void ref_multi_get(const RefMulti *r, int n, int *t_in, int *i_in, int *j_in, int *i_out, int *j_out, double *area, double *di, double *dj)
{
  size_t k = r->interp[n].nxgrid;
  memcpy(t_in, r->interp[n].t_in, k*sizeof(int));   memcpy(i_in, r->interp[n].i_in, k*sizeof(int));
  memcpy(j_in, r->interp[n].j_in, k*sizeof(int));   memcpy(i_out, r->interp[n].i_out, k*sizeof(int));
  memcpy(j_out, r->interp[n].j_out, k*sizeof(int)); memcpy(area, r->interp[n].area, k*sizeof(double));
  if(di && (r->opcode & CONSERVE_ORDER2)) memcpy(di, r->interp[n].di_in, k*sizeof(double));
  if(dj && (r->opcode & CONSERVE_ORDER2)) memcpy(dj, r->interp[n].dj_in, k*sizeof(double));
}

/* do_scalar_conserve_interp on a handle from ref_multi_setup: one variable (interp_method), nz = 1, no missing values;
 * apply_fn NULL = the reference's.  data_in: source tiles back to back, (nx+2*halo)*(ny+2*halo) each (halo 1 for order 2);
 * data_out: output tiles back to back. */
void ref_multi_apply(RefMulti *r, void *apply_fn, int interp_method, const double *data_in, const double *grad_x, const double *grad_y,
                     const int *grad_mask, double *data_out)
{
  Field_config *fin  = (Field_config *)calloc(r->ntiles_in, sizeof(Field_config));
  Field_config *fout = (Field_config *)calloc(r->ntiles_out, sizeof(Field_config));
  Var_config *var = (Var_config *)calloc(1, sizeof(Var_config));
  size_t offd = 0, offg = 0, offo = 0;
  int n, halo = (interp_method == CONSERVE_ORDER2) ? 1 : 0;
  strcpy(var->name, "f");
  var->interp_method = interp_method;
  for(n=0; n<r->ntiles_in; n++) {
    size_t nx = r->gin[n].nx, ny = r->gin[n].ny;
    fin[n].var = var;
    fin[n].data = (double *)(data_in + offd);
    offd += (nx+2*halo)*(ny+2*halo);
    if(halo) {
      fin[n].grad_x = (double *)(grad_x + offg); fin[n].grad_y = (double *)(grad_y + offg); fin[n].grad_mask = (int *)(grad_mask + offg);
      offg += nx*ny;
    }
  }
  for(n=0; n<r->ntiles_out; n++) {
    fout[n].var = var;
    fout[n].data = data_out + offo;
    offo += (size_t)r->gout[n].nxc*r->gout[n].nyc;
  }
  if(apply_fn) ((apply_fn_t)apply_fn)(r->interp, 0, r->ntiles_in, r->gin, r->ntiles_out, r->gout, fin, fout, r->opcode & ~(CHECK_CONSERVE|LEGACY_CLIP), 1);
  else do_scalar_conserve_interp(r->interp, 0, r->ntiles_in, r->gin, r->ntiles_out, r->gout, fin, fout, r->opcode & ~CHECK_CONSERVE, 1);
  free(fin); free(fout); free(var);
}
