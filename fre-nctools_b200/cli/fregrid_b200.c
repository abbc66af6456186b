/* fregrid_b200 — the conservative path of the fregrid command line (reference tools/fregrid/fregrid.c) on top of
 * libxgrid_b200: same option names, same mosaic / grid / field / remap file conventions, weight generation and remapping
 * on the GPU.  Host code is plain C; there is no CPU regridding path (without the library and a device it stops).
 *
 * What is mirrored (reference file:line):
 *   options and their checks                       fregrid.c:368-640
 *   input mosaic -> per-tile corner / centre grids get_input_grid, fregrid_util.c:157-360 (supergrid every 2nd point, x D2R)
 *   cubed-sphere halos from the mosaic contacts    read_mosaic_contact read_mosaic.c:657-762, setup_boundary / update_halo
 *                                                  fregrid_util.c:2420-2660
 *   output grid from --nlon/--nlat or a mosaic     get_output_grid_by_size :564-659, get_output_grid_from_mosaic :414
 *   remap file naming, READ when it exists         set_remap_file :1946-1995
 *   field metadata, scale/offset/missing           get_input_metadata :824-1424, get_field_attribute :1890
 *   output axes, bounds, attributes, history       set_output_metadata :1472-1880, print_provenance tool_util.c:780
 *   per time / level remapping loop                fregrid.c:1010-1085, get_input_data :2036, write_field_data :2339
 *   --check_conserve report                        conserve_interp.c:448-487
 * Not built (refused with a message): bilinear and vector remapping, --extrapolate / --dst_vgrid, --test_case,
 * cell_measures area files, MPI decomposition, netCDF-4 OUTPUT (input files may be classic, 64-bit offset, CDF-5 or
 * netCDF-4/HDF5: csrc/nc3.c, csrc/h5r.c).
 * Unlike the reference, all levels of a time step go to the device as one batch; every level is still remapped exactly as
 * one reference call would remap it (include/xgrid_b200.h, Part 2b).
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/xgrid_b200.h"
#include "../csrc/nc3.h"

#define STRING 255                 /* constant.h:25 */
#define MAXTILE 64
#define MAXVAR 128
#define MAXAXIS 16
#define D2R (M_PI / 180.)
#define R2D (180. / M_PI)
#define RADIUS 6371000.0
enum { ROT_ZERO = 0, ROT_NINETY = 1, ROT_MINUS_NINETY = -1, ROT_180 = 2 };      /* globals.h:33-36 */
enum { EAST = 4, NORTH = 5, WEST = 6, SOUTH = 7 };                               /* globals.h:39-42 */
enum { METHOD_MEAN = 0, METHOD_SUM = 1 };

static void die(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
static void die(const char *fmt, ...)
{
  va_list ap;
  fprintf(stderr, "FATAL Error: ");                                              /* mpp_error, mpp.c:290-298 */
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n");
  exit(1);
}
static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) die("fregrid: out of memory"); return p; }
#define XGB(call) do { if (call) die("%s", xgb_last_error()); } while (0)

/* ------------------------------------------------------------------------------------------------------------- */
/* netCDF helpers */
static nc3_file *nc_open_or_die(const char *path)
{
  char err[400];
  nc3_file *f = nc3_open(path, err, sizeof err);
  if (!f) die("mpp_io(mpp_open): error in opening file %s: %s", path, err);
  return f;
}
static int need_var(nc3_file *f, const char *file, const char *name)
{
  int v = nc3_var_id(f, name);
  if (v < 0) die("mpp_io(mpp_get_varid): error in get field_id of variable %s from file %s", name, file);
  return v;
}
static long long need_dim(nc3_file *f, const char *file, const char *name)
{
  int d = nc3_dim_id(f, name);
  if (d < 0) die("mpp_io(mpp_get_dimlen): error in inquiring dimension %s from file %s", name, file);
  return nc3_dim_len(f, d);
}
static void read_string_row(nc3_file *f, const char *file, const char *var, int row, char *out)
{
  const int v = need_var(f, file, var);
  const int *dd = nc3_var_dimids(f, v);
  size_t start[2] = {(size_t)row, 0}, count[2] = {1, 0};
  char buf[1024];
  if (nc3_var_ndims(f, v) == 1) { start[0] = 0; count[0] = (size_t)nc3_dim_len(f, dd[0]); count[1] = 1; }
  else count[1] = (size_t)nc3_dim_len(f, dd[1]);
  if ((nc3_var_ndims(f, v) == 1 ? count[0] : count[1]) >= sizeof buf) die("fregrid: string variable %s in %s is too long", var, file);
  memset(buf, 0, sizeof buf);
  if (nc3_get_vara_text(f, v, start, count, buf)) die("%s: %s", file, nc3_strerror(f));
  strncpy(out, buf, STRING); out[STRING - 1] = 0;
  for (int k = (int)strlen(out) - 1; k >= 0 && (out[k] == ' '); --k) out[k] = 0;
}
static void dir_of(const char *path, char *dir)                                  /* get_file_path */
{
  const char *s = strrchr(path, '/');
  if (!s) strcpy(dir, ".");
  else { size_t n = (size_t)(s - path); memcpy(dir, path, n); dir[n] = 0; if (n == 0) strcpy(dir, "/"); }
}
static int file_exists(const char *p) { struct stat st; return stat(p, &st) == 0; }

/* ------------------------------------------------------------------------------------------------------------- */
/* grids */
typedef struct {
  int nx, ny;
  double *lonc, *latc;            /* (ny+1)*(nx+1) corners, radians */
  double *lont, *latt;            /* (ny+2)*(nx+2) centres with halo (order 2 only) */
  double *lont1D, *latt1D, *lonc1D, *latc1D;
} Tile;

typedef struct { int nbound; int *is1, *ie1, *js1, *je1, *is2, *ie2, *js2, *je2, *rotate, *tile2; } Bound;

typedef struct {
  int ntiles;
  char dir[1024];
  char gridfile[MAXTILE][STRING], gridtile[MAXTILE][STRING];
  int ncontact;
  int *tile, *istart, *iend, *jstart, *jend;       /* 2*ncontact each: first the "1" sides, then the "2" sides */
} Mosaic;

static void tokenize_str(const char *s, const char *seps, char out[][STRING], int maxn, int *n)
{
  static char buf[40960];                                      /* fregrid.c MAXSTRING-sized entries */
  char *save = NULL, *t;
  if (strlen(s) >= sizeof buf) die("fregrid: the entry '%.40s...' is too long", s);
  strcpy(buf, s);
  *n = 0;
  for (t = strtok_r(buf, seps, &save); t; t = strtok_r(NULL, seps, &save)) {
    if (*n >= maxn) die("fregrid: too many tokens in '%s'", s);
    strncpy(out[*n], t, STRING - 1); out[*n][STRING - 1] = 0; ++*n;
  }
}

static int to_model_index(int s_in, int e_in, int *s_out, int *e_out)           /* transfer_to_model_index, refine 2 */
{
  if (s_in == e_in) { *s_out = (s_in + 1) / 2 - 1; *e_out = *s_out; return 0; }
  if (e_in > s_in) { *s_out = s_in - 1; *e_out = e_in - 2; }
  else             { *s_out = s_in - 2; *e_out = e_in - 1; }
  if (*s_out % 2 || *e_out % 2) die("Error from read_mosaic: mismatch between refine_ratio and istart_in/iend_in");
  *s_out /= 2; *e_out /= 2;
  return 1;
}

static void load_mosaic(const char *path, Mosaic *m)
{
  nc3_file *f = nc_open_or_die(path);
  memset(m, 0, sizeof *m);
  m->ntiles = (int)need_dim(f, path, "ntiles");
  if (m->ntiles > MAXTILE) die("fregrid: more than %d tiles in %s", MAXTILE, path);
  dir_of(path, m->dir);
  for (int n = 0; n < m->ntiles; ++n) {
    read_string_row(f, path, "gridfiles", n, m->gridfile[n]);
    read_string_row(f, path, "gridtiles", n, m->gridtile[n]);
  }
  m->ncontact = nc3_dim_id(f, "ncontact") >= 0 ? (int)need_dim(f, path, "ncontact") : 0;   /* read_mosaic_ncontacts */
  if (m->ncontact > 0) {
    const int nc = m->ncontact;
    m->tile = xmalloc(2 * nc * sizeof(int)); m->istart = xmalloc(2 * nc * sizeof(int)); m->iend = xmalloc(2 * nc * sizeof(int));
    m->jstart = xmalloc(2 * nc * sizeof(int)); m->jend = xmalloc(2 * nc * sizeof(int));
    for (int n = 0; n < nc; ++n) {
      char s[STRING], tok[16][STRING];
      int nt, raw[8], t1, t2;
      read_string_row(f, path, "contacts", n, s);
      tokenize_str(s, ":", tok, 16, &nt);
      if (nt != 4) die("Error from read_mosaic: number of elements in contact seperated by :/:: should be 4");
      t1 = t2 = -1;
      for (int k = 0; k < m->ntiles; ++k) { if (!strcmp(m->gridtile[k], tok[1])) t1 = k; if (!strcmp(m->gridtile[k], tok[3])) t2 = k; }
      if (t1 < 0) die("error from read_mosaic: the first tile name specified in contact is not found in tile list");
      if (t2 < 0) die("error from read_mosaic: the second tile name specified in contact is not found in tile list");
      m->tile[n] = t1; m->tile[n + nc] = t2;
      read_string_row(f, path, "contact_index", n, s);
      tokenize_str(s, ":,", tok, 16, &nt);
      if (nt != 8) die("Error from read_mosaic: number of elements in contact_index seperated by :/, should be 8");
      for (int k = 0; k < 8; ++k) {
        for (const char *c = tok[k]; *c; ++c)
          if (*c > '9' || *c < '0') die("Error from read_mosaic: some of the character in contact_indices except token is not digit number");
        raw[k] = atoi(tok[k]);
      }
      const int i1 = to_model_index(raw[0], raw[1], &m->istart[n], &m->iend[n]);
      const int j1 = to_model_index(raw[2], raw[3], &m->jstart[n], &m->jend[n]);
      const int i2 = to_model_index(raw[4], raw[5], &m->istart[n + nc], &m->iend[n + nc]);
      const int j2 = to_model_index(raw[6], raw[7], &m->jstart[n + nc], &m->jend[n + nc]);
      if (i1 == 0 && j1 == 0) die("Error from read_mosaic_contact:istart1==iend1 and jstart1==jend1");
      if (i2 == 0 && j2 == 0) die("Error from read_mosaic_contact:istart2==iend2 and jstart2==jend2");
      if (i1 + j1 != i2 + j2) die("Error from read_mosaic_contact: It is not a line or overlap contact");
    }
  }
  nc3_close(f);
}

static void init_halo(double *v, int nx, int ny)                                /* halo 1, nz 1 */
{
  const int nxd = nx + 2, nyd = ny + 2;
  for (int j = 0; j < nyd; ++j) { v[j * nxd] = 0; v[j * nxd + nx + 1] = 0; }
  for (int i = 0; i < nxd; ++i) { v[i] = 0; v[(ny + 1) * nxd + i] = 0; }
}

/* one tile of a mosaic: corners (and, when asked, halo-ed centres) from the supergrid; returns the great-circle flag */
static int load_tile(const Mosaic *m, int n, int want_centres, Tile *t)
{
  char path[2048], att[64];
  snprintf(path, sizeof path, "%s/%s", m->dir, m->gridfile[n]);
  nc3_file *g = nc_open_or_die(path);
  int gca = 0;
  if (nc3_get_att_text(g, NC3_GLOBAL, "great_circle_algorithm", att, sizeof att) == 0) gca = !strcmp(att, "TRUE");
  int nx = (int)need_dim(g, path, "nx"), ny = (int)need_dim(g, path, "ny");
  if (nx % 2) die("fregrid_util(get_input_grid): the size of dimension nx should be even (on supergrid)");
  if (ny % 2) die("fregrid_util(get_input_grid): the size of dimension ny should be even (on supergrid)");
  nx /= 2; ny /= 2;
  const size_t ns = (size_t)(2 * nx + 1) * (2 * ny + 1);
  double *x = xmalloc(ns * sizeof(double)), *y = xmalloc(ns * sizeof(double));
  if (nc3_get_var_double(g, need_var(g, path, "x"), x) || nc3_get_var_double(g, need_var(g, path, "y"), y)) die("%s: %s", path, nc3_strerror(g));
  memset(t, 0, sizeof *t);
  t->nx = nx; t->ny = ny;
  t->lonc = xmalloc((size_t)(nx + 1) * (ny + 1) * sizeof(double));
  t->latc = xmalloc((size_t)(nx + 1) * (ny + 1) * sizeof(double));
  t->lont1D = xmalloc(nx * sizeof(double)); t->latt1D = xmalloc(ny * sizeof(double));
  for (int i = 0; i < nx; ++i) t->lont1D[i] = x[2 * nx + 1 + 2 * i + 1] * D2R;
  for (int j = 0; j < ny; ++j) t->latt1D[j] = y[(size_t)(2 * j + 1) * (2 * nx + 1) + 1] * D2R;
  for (int j = 0; j <= ny; ++j) for (int i = 0; i <= nx; ++i) {
    const size_t a = (size_t)j * (nx + 1) + i, b = (size_t)2 * j * (2 * nx + 1) + 2 * i;
    t->lonc[a] = x[b] * D2R; t->latc[a] = y[b] * D2R;
  }
  if (want_centres) {
    t->lont = xmalloc((size_t)(nx + 2) * (ny + 2) * sizeof(double));
    t->latt = xmalloc((size_t)(nx + 2) * (ny + 2) * sizeof(double));
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t a = (size_t)(j + 1) * (nx + 2) + i + 1, b = (size_t)(2 * j + 1) * (2 * nx + 1) + 2 * i + 1;
      t->lont[a] = x[b] * D2R; t->latt[a] = y[b] * D2R;
    }
    init_halo(t->lont, nx, ny); init_halo(t->latt, nx, ny);
  }
  free(x); free(y);
  nc3_close(g);
  return gca;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static void setup_boundary(const Mosaic *m, const Tile *grid, Bound *bound)     /* halo = 1, position = CENTER */
{
  const int nc2 = 2 * m->ncontact, halo = 1, shift = 0;
  int *dir = xmalloc((nc2 ? nc2 : 1) * sizeof(int));
  for (int n = 0; n < m->ntiles; ++n) bound[n].nbound = 0;
  if (m->ncontact == 0) { free(dir); return; }
  for (int l = 0; l < nc2; ++l) {                                                /* get_contact_direction */
    if (m->istart[l] == m->iend[l] && m->jstart[l] == m->jend[l]) die("fregrid_util: istart = iend and jstart = jend can not be both true for one contact");
    if (m->istart[l] != m->iend[l] && m->jstart[l] != m->jend[l]) die("fregrid_util: either istart = iend or jstart = jend need to be true");
    if (m->istart[l] == m->iend[l]) dir[l] = (m->istart[l] == 0) ? WEST : EAST;
    else dir[l] = (m->jstart[l] == 0) ? SOUTH : NORTH;
  }
  for (int n = 0; n < m->ntiles; ++n) {
    const int nx = grid[n].nx, ny = grid[n].ny;
    int nb = 0;
    for (int l = 0; l < nc2; ++l) if (m->tile[l] == n) ++nb;
    Bound *b = &bound[n];
    b->nbound = nb;
    if (!nb) continue;
    b->is1 = xmalloc(nb * sizeof(int)); b->ie1 = xmalloc(nb * sizeof(int)); b->js1 = xmalloc(nb * sizeof(int)); b->je1 = xmalloc(nb * sizeof(int));
    b->is2 = xmalloc(nb * sizeof(int)); b->ie2 = xmalloc(nb * sizeof(int)); b->js2 = xmalloc(nb * sizeof(int)); b->je2 = xmalloc(nb * sizeof(int));
    b->rotate = xmalloc(nb * sizeof(int)); b->tile2 = xmalloc(nb * sizeof(int));
    nb = 0;
    for (int l = 0; l < nc2; ++l) {
      if (m->tile[l] != n) continue;
      const int js = imin(m->jstart[l], m->jend[l]) + halo, je = imax(m->jstart[l], m->jend[l]) + halo + shift;
      const int is = imin(m->istart[l], m->iend[l]) + halo, ie = imax(m->istart[l], m->iend[l]) + halo + shift;
      switch (dir[l]) {
        case WEST:  b->is1[nb] = 0; b->ie1[nb] = halo - 1; b->js1[nb] = js; b->je1[nb] = je; break;
        case EAST:  b->is1[nb] = nx + shift + halo; b->ie1[nb] = nx + shift + halo + halo - 1; b->js1[nb] = js; b->je1[nb] = je; break;
        case SOUTH: b->is1[nb] = is; b->ie1[nb] = ie; b->js1[nb] = 0; b->je1[nb] = halo - 1; break;
        default:    b->is1[nb] = is; b->ie1[nb] = ie; b->js1[nb] = ny + shift + halo; b->je1[nb] = ny + shift + halo + halo - 1; break;
      }
      const int l2 = (l + m->ncontact) % nc2;
      b->tile2[nb] = m->tile[l2];
      const int js2 = imin(m->jstart[l2], m->jend[l2]) + halo, je2 = imax(m->jstart[l2], m->jend[l2]) + halo + shift;
      const int is2 = imin(m->istart[l2], m->iend[l2]) + halo, ie2 = imax(m->istart[l2], m->iend[l2]) + halo + shift;
      switch (dir[l2]) {                          /* (the reference uses this tile's nx, ny here too, :2533-2551) */
        case WEST:  b->is2[nb] = halo + shift; b->ie2[nb] = halo + shift + halo - 1; b->js2[nb] = js2; b->je2[nb] = je2; break;
        case EAST:  b->is2[nb] = nx - halo + 1; b->ie2[nb] = nx; b->js2[nb] = js2; b->je2[nb] = je2; break;
        case SOUTH: b->is2[nb] = is2; b->ie2[nb] = ie2; b->js2[nb] = halo + shift; b->je2[nb] = halo + shift + halo - 1; break;
        default:    b->is2[nb] = is2; b->ie2[nb] = ie2; b->js2[nb] = ny - halo + 1; b->je2[nb] = ny; break;
      }
      b->rotate[nb] = ROT_ZERO;
      if (dir[l] == WEST && dir[l2] == NORTH) b->rotate[nb] = ROT_NINETY;
      if (dir[l] == EAST && dir[l2] == SOUTH) b->rotate[nb] = ROT_NINETY;
      if (dir[l] == SOUTH && dir[l2] == EAST) b->rotate[nb] = ROT_MINUS_NINETY;
      if (dir[l] == NORTH && dir[l2] == WEST) b->rotate[nb] = ROT_MINUS_NINETY;
      if (dir[l] == NORTH && dir[l2] == NORTH) b->rotate[nb] = ROT_180;
      if ((b->ie2[nb] - b->is2[nb] + 1) * (b->je2[nb] - b->js2[nb] + 1) != (b->ie1[nb] - b->is1[nb] + 1) * (b->je1[nb] - b->js1[nb] + 1))
        die("fregrid_util: size mismatch between the boundary");
      ++nb;
    }
  }
  free(dir);
}

/* update_halo (fregrid_util.c:2614-2660) for one level: data[t] = this tile's (nx+2)*(ny+2) array, tiles[] = all tiles' arrays */
static void update_halo(const Tile *grid, int n, double *const *tiles, const Bound *b)
{
  const int nx = grid[n].nx + 2;
  double *data = tiles[n];
  for (int k = 0; k < b->nbound; ++k) {
    const int is1 = b->is1[k], ie1 = b->ie1[k], js1 = b->js1[k], je1 = b->je1[k];
    const int is2 = b->is2[k], ie2 = b->ie2[k], js2 = b->js2[k], je2 = b->je2[k];
    const int nx2 = grid[b->tile2[k]].nx + 2;
    const double *src = tiles[b->tile2[k]];
    double *buf = xmalloc((size_t)(ie2 - is2 + 1) * (je2 - js2 + 1) * sizeof(double));
    int l = 0;
    switch (b->rotate[k]) {
      case ROT_ZERO:         for (int j = js2; j <= je2; ++j) for (int i = is2; i <= ie2; ++i) buf[l++] = src[j * nx2 + i]; break;
      case ROT_NINETY:       for (int i = ie2; i >= is2; --i) for (int j = js2; j <= je2; ++j) buf[l++] = src[j * nx2 + i]; break;
      case ROT_MINUS_NINETY: for (int i = is2; i <= ie2; ++i) for (int j = je2; j >= js2; --j) buf[l++] = src[j * nx2 + i]; break;
      default:               for (int j = je2; j >= js2; --j) for (int i = ie2; i >= is2; --i) buf[l++] = src[j * nx2 + i]; break;
    }
    l = 0;
    for (int j = js1; j <= je1; ++j) for (int i = is1; i <= ie1; ++i) data[j * nx + i] = buf[l++];
    free(buf);
  }
}

/* the destination tile on the device: the --nlon/--nlat grid is built there (get_output_grid_by_size's arithmetic, nothing to
   upload), a mosaic tile is uploaded */
static int set_dst_tile(xgb_plan *plan, const Tile *g, int by_size, int nlon, int nlat, double lonbegin, double lonend,
                        double latbegin, double latend)
{
  if (by_size) return xgb_plan_set_dst_latlon(plan, nlon, nlat, lonbegin, lonend, latbegin, latend);
  return xgb_plan_set_dst(plan, g->nx, g->ny, g->lonc, g->latc, 0);
}

/* ------------------------------------------------------------------------------------------------------------- */
/* exchange grid of one output tile, host lists */
typedef struct {
  long long n;
  int *t_in, *i_in, *j_in, *i_out, *j_out;
  double *area, *di, *dj;
  char remap_file[STRING + 8];
  int file_exist;
} Xgrid;

static void xgrid_alloc(Xgrid *x, long long n, int order)
{
  const size_t k = (size_t)(n > 0 ? n : 1);
  x->n = n;
  x->t_in = xmalloc(k * sizeof(int)); x->i_in = xmalloc(k * sizeof(int)); x->j_in = xmalloc(k * sizeof(int));
  x->i_out = xmalloc(k * sizeof(int)); x->j_out = xmalloc(k * sizeof(int));
  x->area = xmalloc(k * sizeof(double));
  x->di = order == 2 ? xmalloc(k * sizeof(double)) : NULL;
  x->dj = order == 2 ? xmalloc(k * sizeof(double)) : NULL;
}

/* ------------------------------------------------------------------------------------------------------------- */
/* fields */
typedef struct {
  char name[STRING], bndname[STRING];
  char cart;
  int type, vid_in, size, bndtype, bnd_vid_in;
  double *data, *bnddata;
  int dimid, vid, bndid;          /* in the output file */
} Axis;

typedef struct {
  char name[STRING];
  int vid_in, type, ndim, index[5];
  int do_regrid, has_taxis, has_zaxis, has_naxis, nz, nn, kstart, kend, lstart;
  int has_missing, order, cell_methods;
  double missing, scale, offset;
  int vid_out;
} Var;

static char var_cart(nc3_file *f, int vid)                                      /* mpp_get_var_cart */
{
  char s[16];
  if (vid < 0) return 'N';
  if (nc3_get_att_text(f, vid, "cartesian_axis", s, sizeof s) == 0 && s[0]) return s[0];
  if (nc3_get_att_text(f, vid, "axis", s, sizeof s) == 0 && s[0]) return s[0];
  return 'N';
}
static void var_bndname(nc3_file *f, int vid, char *out)                        /* mpp_get_var_bndname */
{
  if (nc3_get_att_text(f, vid, "climatology", out, STRING) == 0) return;
  if (nc3_get_att_text(f, vid, "bounds", out, STRING) == 0) return;
  if (nc3_get_att_text(f, vid, "edges", out, STRING) == 0) return;
  strcpy(out, "none");
}

static const char *usage =
  "fregrid_b200 --input_mosaic input_mosaic [--output_mosaic output_mosaic | --nlon #lon --nlat #lat]\n"
  "             [--input_dir dir] [--input_file file] [--scalar_field a,b,...] [--output_dir dir] [--output_file file]\n"
  "             [--remap_file file] [--interp_method conserve_order1|conserve_order2|conserve_order2_monotonic]\n"
  "             [--lonBegin #] [--lonEnd #] [--latBegin #] [--latEnd #] [--KlevelBegin #] [--KlevelEnd #]\n"
  "             [--LstepBegin #] [--LstepEnd #] [--check_conserve] [--target_grid] [--weight_file f --weight_field w]\n"
  "             [--standard_dimension] [--format classic|64bit_offset|cdf5] [--debug] [--gpus N | --gpu_list d0,d1,...]\n"
  "Conservative remapping of scalar fields between mosaics (the conservative path of FRE-NCtools fregrid) on B200 GPUs.\n"
  "--gpus N: exchange-grid generation sharded over N devices of this box (one process; replaces mpirun fregrid_parallel).\n";

int main(int argc, char **argv)
{
  const char *mosaic_in = NULL, *mosaic_out = NULL, *dir_in = NULL, *dir_out = NULL, *remap_file = NULL;
  const char *weight_file = NULL, *weight_field = NULL, *format = NULL;
  char input_file[STRING] = "", output_file[STRING] = "", interp_method[STRING] = "conserve_order1";
  char scalar_name[MAXVAR][STRING];
  int nscalar = 0, nfiles = 0, nfiles_out = 0, nlon = 0, nlat = 0, check_conserve = 0, debug = 0, target_grid = 0;
  int kbegin = 0, kend = -1, lbegin = 0, lend = -1, standard_dimension = 0;
  int ngpus = 1, gpu_list[64], have_gpu_list = 0;
  double lonbegin = 0, lonend = 360, latbegin = -90, latend = 90;
  unsigned opcode = 0;
  int c, idx = 0, errflg = argc == 1;
  static struct option opts[] = {
    {"input_mosaic", required_argument, NULL, 'a'}, {"output_mosaic", required_argument, NULL, 'b'},
    {"input_dir", required_argument, NULL, 'c'}, {"output_dir", required_argument, NULL, 'd'},
    {"input_file", required_argument, NULL, 'e'}, {"output_file", required_argument, NULL, 'f'},
    {"remap_file", required_argument, NULL, 'g'}, {"test_case", required_argument, NULL, 'i'},
    {"interp_method", required_argument, NULL, 'j'}, {"test_parameter", required_argument, NULL, 'k'},
    {"symmetry", no_argument, NULL, 'l'}, {"grid_type", required_argument, NULL, 'm'},
    {"target_grid", no_argument, NULL, 'n'}, {"finer_step", required_argument, NULL, 'o'},
    {"fill_missing", no_argument, NULL, 'p'}, {"nlon", required_argument, NULL, 'q'}, {"nlat", required_argument, NULL, 'r'},
    {"scalar_field", required_argument, NULL, 's'}, {"check_conserve", no_argument, NULL, 't'},
    {"u_field", required_argument, NULL, 'u'}, {"v_field", required_argument, NULL, 'v'},
    {"center_y", no_argument, NULL, 'y'}, {"lonBegin", required_argument, NULL, 'A'}, {"lonEnd", required_argument, NULL, 'B'},
    {"latBegin", required_argument, NULL, 'C'}, {"latEnd", required_argument, NULL, 'D'},
    {"KlevelBegin", required_argument, NULL, 'E'}, {"KlevelEnd", required_argument, NULL, 'F'},
    {"LstepBegin", required_argument, NULL, 'G'}, {"LstepEnd", required_argument, NULL, 'H'},
    {"weight_file", required_argument, NULL, 'I'}, {"weight_field", required_argument, NULL, 'J'},
    {"extrapolate", no_argument, NULL, 'L'}, {"dst_vgrid", required_argument, NULL, 'M'},
    {"stop_crit", required_argument, NULL, 'N'}, {"standard_dimension", no_argument, NULL, 'O'},
    {"debug", no_argument, NULL, 'P'}, {"nthreads", required_argument, NULL, 'Q'},
    {"associated_file_dir", required_argument, NULL, 'R'}, {"deflation", required_argument, NULL, 'S'},
    {"shuffle", required_argument, NULL, 'T'}, {"format", required_argument, NULL, 'U'}, {"help", no_argument, NULL, 'h'},
    {"gpus", required_argument, NULL, 'V'}, {"gpu_list", required_argument, NULL, 'W'},
    {0, 0, 0, 0}};
  char tok[MAXVAR][STRING];
  int ntok;

  while ((c = getopt_long(argc, argv, "", opts, &idx)) != -1) {
    switch (c) {
      case 'a': mosaic_in = optarg; break;
      case 'b': mosaic_out = optarg; break;
      case 'c': dir_in = optarg; break;
      case 'd': dir_out = optarg; break;
      case 'e': tokenize_str(optarg, ",", tok, MAXVAR, &ntok); nfiles = ntok; if (ntok) strcpy(input_file, tok[0]); break;
      case 'f': tokenize_str(optarg, ",", tok, MAXVAR, &ntok); nfiles_out = ntok; if (ntok) strcpy(output_file, tok[0]); break;
      case 'g': remap_file = optarg; break;
      case 's': tokenize_str(optarg, ",", scalar_name, MAXVAR, &nscalar); break;
      case 'j': strncpy(interp_method, optarg, STRING - 1); break;
      case 'n': target_grid = 1; break;
      case 'q': nlon = atoi(optarg); break;
      case 'r': nlat = atoi(optarg); break;
      case 't': check_conserve = 1; break;
      case 'y': break;                                       /* centre-y is what the conservative path always uses (:729) */
      case 'A': lonbegin = atof(optarg); break;
      case 'B': lonend = atof(optarg); break;
      case 'C': latbegin = atof(optarg); break;
      case 'D': latend = atof(optarg); break;
      case 'E': kbegin = atoi(optarg); break;
      case 'F': kend = atoi(optarg); break;
      case 'G': lbegin = atoi(optarg); break;
      case 'H': lend = atoi(optarg); break;
      case 'I': weight_file = optarg; break;
      case 'J': weight_field = optarg; break;
      case 'O': standard_dimension = 1; break;
      case 'P': debug = 1; break;
      case 'Q': case 'S': case 'T': case 'R': case 'N': break;  /* threads / deflation / shuffle do not apply here */
      case 'U': format = optarg; break;
      case 'V': ngpus = atoi(optarg); if (ngpus < 1 || ngpus > 64) die("fregrid_b200: --gpus must be between 1 and 64"); break;
      case 'W': tokenize_str(optarg, ",", tok, MAXVAR, &ntok);
                if (ntok < 1 || ntok > 64) die("fregrid_b200: --gpu_list takes 1 to 64 device numbers");
                ngpus = ntok; have_gpu_list = 1;
                for (int q = 0; q < ntok; ++q) gpu_list[q] = atoi(tok[q]);
                break;
      case 'u': case 'v': case 'm':
        die("fregrid: conservative interpolation of vector fields is not supported. \n"
            "Use bilinear interpolation or regrid the vector components independently as scalars.");
      case 'i': case 'k': die("fregrid_b200: --test_case is not built");
      case 'l': die("fregrid_b200: --symmetry applies to bilinear remapping, which is not built");
      case 'o': case 'p': die("fregrid_b200: --finer_step / --fill_missing apply to bilinear remapping, which is not built");
      case 'L': case 'M': die("fregrid_b200: --extrapolate / --dst_vgrid are not built");
      case 'h': fputs(usage, stdout); return 0;
      default: errflg++;
    }
  }
  if (errflg) { fputs(usage, stderr); return 2; }
  /* fregrid.c:571-640 */
  if (!mosaic_in) die("fregrid: input_mosaic is not specified");
  if (!mosaic_out) {
    if (nlon == 0 || nlat == 0) die("fregrid: when output_mosaic is not specified, nlon and nlat should be specified");
    if (lonend <= lonbegin) die("fregrid: when output_mosaic is not specified, lonEnd should be larger than lonBegin");
    if (latend <= latbegin) die("fregrid: when output_mosaic is not specified, latEnd should be larger than latBegin");
  } else if (nlon != 0 || nlat != 0) die("fregrid: when output_mosaic is specified, nlon and nlat should not be specified");
  if (!strcmp(interp_method, "conserve_order1")) { printf("****fregrid: first order conservative scheme will be used for regridding.\n"); opcode |= XGB_CONSERVE_ORDER1; }
  else if (!strcmp(interp_method, "conserve_order2")) { printf("****fregrid: second order conservative scheme will be used for regridding.\n"); opcode |= XGB_CONSERVE_ORDER2; }
  else if (!strcmp(interp_method, "conserve_order2_monotonic")) {
    printf("****fregrid: second order monotonic conservative scheme will be used for regridding.\n");
    opcode |= XGB_CONSERVE_ORDER2 | XGB_MONOTONIC;
  } else if (!strcmp(interp_method, "bilinear")) die("fregrid_b200: bilinear remapping is not built; use the reference fregrid for it");
  else die("fregrid: interp_method must be 'conserve_order1', 'conserve_order2', 'conserve_order2_monotonic'  or 'bilinear'");
  int save_weight_only = 0;
  if (nfiles == 0) {
    if (nscalar > 0) die("fregrid: when --input_file is not specified, --scalar_field, --u_field and --v_field should also not be specified");
    if (!remap_file) die("fregrid: when --input_file is not specified, remap_file must be specified to save weight information");
    save_weight_only = 1;
    printf("NOTE: No input file specified in this run, no data file will be regridded and only weight information is calculated.\n");
  } else if (nfiles == 1) {
    if (nscalar == 0) die("fregrid: both scalar_field and vector_field are not specified");
    if (nfiles_out == 0) strcpy(output_file, input_file);
    else if (nfiles_out != nfiles) die("fregrid:number of input file is not equal to number of output file");
  } else die("fregrid: when scalar_field is specified, number of files must be 1");
  if (kbegin != 0 || kend != -1) if (kbegin < 1 || kend < kbegin) die("fregrid:KlevelBegin should be a positive integer and no larger than KlevelEnd when you want pick certain klevel");
  if (lbegin != 0 || lend != -1) if (lbegin < 1 || lend < lbegin) die("fregrid:LstepBegin should be a positive integer and no larger than LstepEnd when you want pick certain Lstep");
  char wt_file_obj[2 * STRING];
  if (weight_field && !weight_file) {
    if (nfiles == 0) die("fregrid: weight_field is specified, but both weight_file and input_file are not specified");
    snprintf(wt_file_obj, sizeof wt_file_obj, "%s/%s", dir_in ? dir_in : ".", input_file);
    weight_file = wt_file_obj;
  }
  if (format && xgb_set_nc_format(format)) die("%s", xgb_last_error());
  char history[4096] = "";
  for (int i = 0; i < argc; ++i) { if (i) strncat(history, " ", sizeof history - strlen(history) - 1); strncat(history, argv[i], sizeof history - strlen(history) - 1); }

  /* ---- grids ---- */
  Mosaic min, mout;
  load_mosaic(mosaic_in, &min);
  const int ntiles_in = min.ntiles;
  const int order = (opcode & XGB_CONSERVE_ORDER2) ? 2 : 1;
  if (ntiles_in != 6 && order == 2) die("fregrid: when the input grid is not cubic sphere grid, interp_method can not be conserve_order2");
  Tile *gin = xmalloc(ntiles_in * sizeof(Tile));
  Bound *bound = xmalloc(ntiles_in * sizeof(Bound));
  int gca_in = 0, gca_out = 0;
  const int read_tgrid = !(save_weight_only || order == 1);                     /* fregrid_util.c:181-183 */
  for (int n = 0; n < ntiles_in; ++n) { const int g = load_tile(&min, n, read_tgrid, &gin[n]); if (n == 0) gca_in = g; }
  setup_boundary(&min, gin, bound);
  if (read_tgrid) {
    double *lt[MAXTILE], *la[MAXTILE];
    for (int n = 0; n < ntiles_in; ++n) { lt[n] = gin[n].lont; la[n] = gin[n].latt; }
    for (int n = 0; n < ntiles_in; ++n) { update_halo(gin, n, lt, &bound[n]); update_halo(gin, n, la, &bound[n]); }
  }
  int ntiles_out = 1;
  Tile *gout;
  if (mosaic_out) {
    load_mosaic(mosaic_out, &mout);
    ntiles_out = mout.ntiles;
    gout = xmalloc(ntiles_out * sizeof(Tile));
    for (int n = 0; n < ntiles_out; ++n) { const int g = load_tile(&mout, n, 0, &gout[n]); if (n == 0) gca_out = g; }
  } else {
    gout = xmalloc(sizeof(Tile));
    memset(gout, 0, sizeof(Tile));
    Tile *g = gout;
    g->nx = nlon; g->ny = nlat;
    g->lont1D = xmalloc(nlon * sizeof(double)); g->latt1D = xmalloc(nlat * sizeof(double));
    g->lonc1D = xmalloc((nlon + 1) * sizeof(double)); g->latc1D = xmalloc((nlat + 1) * sizeof(double));
    const double dlon = (lonend - lonbegin) / nlon, dlat = (latend - latbegin) / nlat;   /* get_output_grid_by_size, centre-y */
    for (int i = 0; i < nlon; ++i) g->lont1D[i] = (lonbegin + (i + 0.5) * dlon) * D2R;
    for (int i = 0; i <= nlon; ++i) g->lonc1D[i] = (lonbegin + i * dlon) * D2R;
    for (int j = 0; j < nlat; ++j) g->latt1D[j] = (latbegin + (j + 0.5) * dlat) * D2R;
    for (int j = 0; j <= nlat; ++j) g->latc1D[j] = (latbegin + j * dlat) * D2R;
    g->lonc = xmalloc((size_t)(nlon + 1) * (nlat + 1) * sizeof(double));
    g->latc = xmalloc((size_t)(nlon + 1) * (nlat + 1) * sizeof(double));
    for (int j = 0; j <= nlat; ++j) for (int i = 0; i <= nlon; ++i) {
      g->lonc[(size_t)j * (nlon + 1) + i] = g->lonc1D[i];
      g->latc[(size_t)j * (nlon + 1) + i] = g->latc1D[j];
    }
  }
  if (gca_in || gca_out) {
    opcode |= XGB_GREAT_CIRCLE;
    if (order != 1) die("fregrid: when clip_method is 'conserve_great_circle', interp_methos need to be 'conserve_order1', contact developer");
  }

  /* ---- the device plan: source mosaic once ---- */
  xgb_plan *plan = xgb_plan_create(getenv("XGB_DEVICE") ? atoi(getenv("XGB_DEVICE")) : 0);
  if (!plan) die("%s", xgb_last_error());
  int nxs[MAXTILE], nys[MAXTILE];
  size_t ncorner = 0, ncell = 0, nhalo = 0;
  for (int n = 0; n < ntiles_in; ++n) {
    nxs[n] = gin[n].nx; nys[n] = gin[n].ny;
    ncorner += (size_t)(gin[n].nx + 1) * (gin[n].ny + 1); ncell += (size_t)gin[n].nx * gin[n].ny; nhalo += (size_t)(gin[n].nx + 2) * (gin[n].ny + 2);
  }
  double *lon_cat = xmalloc(ncorner * sizeof(double)), *lat_cat = xmalloc(ncorner * sizeof(double));
  { size_t o = 0; for (int n = 0; n < ntiles_in; ++n) { const size_t k = (size_t)(gin[n].nx + 1) * (gin[n].ny + 1);
      memcpy(lon_cat + o, gin[n].lonc, k * sizeof(double)); memcpy(lat_cat + o, gin[n].latc, k * sizeof(double)); o += k; } }
  XGB(xgb_plan_set_src(plan, ntiles_in, nxs, nys, lon_cat, lat_cat, NULL, 0));

  /* ---- remap files (set_remap_file) and the exchange grids ---- */
  Xgrid *xg = xmalloc(ntiles_out * sizeof(Xgrid));
  memset(xg, 0, ntiles_out * sizeof(Xgrid));
  int do_write = 0;
  if (remap_file) {
    char base[STRING];
    const size_t len = strlen(remap_file);
    if (len >= STRING) die("setoutput_remap_file(fregrid_util): length of remap_file should be less than STRING");
    strcpy(base, remap_file);
    if (len > 3 && !strcmp(remap_file + len - 3, ".nc")) base[len - 3] = 0;
    do_write = 1;
    for (int m = 0; m < ntiles_out; ++m) {
      if (ntiles_out > 1) snprintf(xg[m].remap_file, sizeof xg[m].remap_file, "%s.%s.nc", base, mout.gridtile[m]);
      else snprintf(xg[m].remap_file, sizeof xg[m].remap_file, "%s.nc", base);
      if (!save_weight_only && file_exists(xg[m].remap_file)) xg[m].file_exist = 1;
    }
  }
  int any_read = 0;
  for (int m = 0; m < ntiles_out; ++m) any_read |= xg[m].file_exist;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  /* order 2 onto several output tiles: sums per source cell across the tiles (one device); several devices: one output tile at a time */
  const int o2_multi = (order == 2 && ntiles_out > 1 && !(opcode & XGB_GREAT_CIRCLE) && !any_read);
  const int multi_gpu = ((ngpus > 1 || have_gpu_list) && !(opcode & XGB_GREAT_CIRCLE) && !o2_multi);
  double *raw_clon[MAXTILE] = {0}, *raw_clat[MAXTILE] = {0};
  if (o2_multi) XGB(xgb_plan_order2_begin(plan));
  if ((ngpus > 1 || have_gpu_list) && !multi_gpu && !any_read)
    printf("NOTE: --gpus applies to conserve_order1/2 weight generation onto one output tile at a time; this run uses one device.\n");
  for (int m = 0; m < ntiles_out; ++m) {
    if (any_read) {                                         /* conserve_interp.c:62-125: READ reads, and only reads */
      if (!xg[m].file_exist) continue;
      const long long n = xgb_remap_size(xg[m].remap_file);
      if (n < 0) die("%s", xgb_last_error());
      xgrid_alloc(&xg[m], n, order);
      XGB(xgb_remap_read(xg[m].remap_file, order, n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out, xg[m].area, xg[m].di, xg[m].dj));
      continue;
    }
    if (multi_gpu) {
      /* one process, ngpus devices: source-cell windows dealt to the devices, pieces put back in the serial order on the host
         (csrc/multi_gpu.cu) — the reference's mpirun -n N fregrid_parallel (fregrid_util.c:489-492, conserve_interp.c:404-437) */
      xgb_dst_spec ds;
      memset(&ds, 0, sizeof ds);
      if (!mosaic_out) { ds.by_size = 1; ds.nlon = nlon; ds.nlat = nlat; ds.lonbegin = lonbegin; ds.lonend = lonend; ds.latbegin = latbegin; ds.latend = latend; }
      else { ds.nx = gout[m].nx; ds.ny = gout[m].ny; ds.lon = gout[m].lonc; ds.lat = gout[m].latc; }
      xgb_host_xgrid hx;
      XGB(xgb_generate_multi_gpu(ngpus, have_gpu_list ? gpu_list : NULL, opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2), &ds, ntiles_in, nxs, nys,
                                 lon_cat, lat_cat, NULL, 0, &hx));
      const long long n = hx.nxgrid;
      xgrid_alloc(&xg[m], n, order);
      if (n > 0) {
        memcpy(xg[m].t_in, hx.t_in, n * sizeof(int)); memcpy(xg[m].i_in, hx.i_in, n * sizeof(int)); memcpy(xg[m].j_in, hx.j_in, n * sizeof(int));
        memcpy(xg[m].i_out, hx.i_out, n * sizeof(int)); memcpy(xg[m].j_out, hx.j_out, n * sizeof(int)); memcpy(xg[m].area, hx.area, n * sizeof(double));
        if (order == 2) { memcpy(xg[m].di, hx.di, n * sizeof(double)); memcpy(xg[m].dj, hx.dj, n * sizeof(double)); }
      }
      xgb_host_xgrid_free(&hx);
      if (do_write && n > 0)
        XGB(xgb_remap_write(xg[m].remap_file, order, n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out, 0, 0, xg[m].area, xg[m].di, xg[m].dj));
      continue;
    }
    XGB(set_dst_tile(plan, &gout[m], !mosaic_out, nlon, nlat, lonbegin, lonend, latbegin, latend));
    const long long n = xgb_plan_generate(plan, opcode & (XGB_CONSERVE_ORDER1 | XGB_CONSERVE_ORDER2 | XGB_GREAT_CIRCLE));
    if (n < 0) die("%s", xgb_last_error());
    xgrid_alloc(&xg[m], n, order);
    if (n > 0) XGB(xgb_plan_result_host(plan, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out, xg[m].area,
                                        o2_multi ? NULL : xg[m].di, o2_multi ? NULL : xg[m].dj));
    if (o2_multi && n > 0) {                                /* raw centroids; the distances follow after the last output tile */
      raw_clon[m] = xmalloc(n * sizeof(double)); raw_clat[m] = xmalloc(n * sizeof(double));
      XGB(xgb_plan_result_centroids_host(plan, raw_clon[m], raw_clat[m]));
    }
    if (do_write && n > 0 && !o2_multi)
      XGB(xgb_remap_write(xg[m].remap_file, order, n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out, 0, 0, xg[m].area, xg[m].di, xg[m].dj));
  }
  if (o2_multi && !any_read) {
    /* conserve_interp.c:204-221, :319-358: the per-source-cell sums run over ALL output tiles before the AREA_RATIO test */
    XGB(xgb_plan_order2_end(plan));
    for (int m = 0; m < ntiles_out; ++m) {
      if (xg[m].n > 0) {
        XGB(xgb_plan_order2_distance(plan, xg[m].n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].area, raw_clon[m], raw_clat[m], xg[m].di, xg[m].dj));
        if (do_write)
          XGB(xgb_remap_write(xg[m].remap_file, order, xg[m].n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out, 0, 0, xg[m].area, xg[m].di, xg[m].dj));
      }
      free(raw_clon[m]); free(raw_clat[m]);
    }
    xgb_plan_order2_reset(plan);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (any_read) printf("NOTE: Finish reading index and weight for conservative interpolation from file.\n");
  else printf("NOTE: done calculating index and weight for conservative interpolation\n");
  if (debug) printf("setup_interp took %.3f s wall (%lld exchange cells in tile 1)\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), xg[0].n);

  if (check_conserve) {                                     /* conserve_interp.c:448-487 */
    for (int m = 0; m < ntiles_out; ++m) {
      const int nx1 = gout[m].nx, ny1 = gout[m].ny;
      double *area2 = xmalloc((size_t)nx1 * ny1 * sizeof(double)), *cell_area = xmalloc((size_t)nx1 * ny1 * sizeof(double));
      XGB(set_dst_tile(plan, &gout[m], !mosaic_out, nlon, nlat, lonbegin, lonend, latbegin, latend));
      XGB(xgb_plan_dst_area_host(plan, cell_area));
      for (size_t i = 0; i < (size_t)nx1 * ny1; ++i) area2[i] = 0;
      for (long long i = 0; i < xg[m].n; ++i) area2[(size_t)xg[m].j_out[i] * nx1 + xg[m].i_out[i]] += xg[m].area[i];
      double max_ratio = 0; int max_i = 0, max_j = 0;
      for (int j = 0; j < ny1; ++j) for (int i = 0; i < nx1; ++i) {
        const size_t ii = (size_t)j * nx1 + i;
        const double r = fabs(cell_area[ii] - area2[ii]) / cell_area[ii];
        if (r > max_ratio) { max_ratio = r; max_i = i; max_j = j; }
        if (r > 1.e-4) printf("(i,j)=(%d,%d), change = %g, area1=%g, area2=%g\n", i, j, r, cell_area[ii], area2[ii]);
      }
      const size_t ii = (size_t)max_j * nx1 + max_i;
      printf("The maximum ratio change at (%d,%d) = %g, area1=%g, area2=%g\n", max_i, max_j, max_ratio, cell_area[ii], area2[ii]);
      free(area2); free(cell_area);
    }
  }
  if (save_weight_only) {
    printf("NOTE: Successfully running fregrid and the following files which store weight information are generated.\n");
    for (int m = 0; m < ntiles_out; ++m) printf("****%s\n", xg[m].remap_file);
    xgb_plan_destroy(plan);
    return 0;
  }

  /* ---- input files, field metadata (tile 0 decides, like the reference) ---- */
  char base_in[2 * STRING], base_out[2 * STRING], name_in[MAXTILE][3 * STRING], name_out[MAXTILE][3 * STRING];
  { char s[STRING]; strcpy(s, input_file); size_t l = strlen(s); if (l > 3 && !strcmp(s + l - 3, ".nc")) s[l - 3] = 0;
    if (dir_in) snprintf(base_in, sizeof base_in, "%s/%s", dir_in, s); else strcpy(base_in, s);
    strcpy(s, output_file); l = strlen(s); if (l > 3 && !strcmp(s + l - 3, ".nc")) s[l - 3] = 0;
    if (dir_out) snprintf(base_out, sizeof base_out, "%s/%s", dir_out, s); else strcpy(base_out, s); }
  nc3_file *fin[MAXTILE];
  for (int n = 0; n < ntiles_in; ++n) {
    if (ntiles_in > 1) snprintf(name_in[n], sizeof name_in[n], "%s.%s.nc", base_in, min.gridtile[n]);
    else snprintf(name_in[n], sizeof name_in[n], "%s.nc", base_in);
    fin[n] = nc_open_or_die(name_in[n]);
  }
  for (int m = 0; m < ntiles_out; ++m) {
    if (ntiles_out > 1) snprintf(name_out[m], sizeof name_out[m], "%s.%s.nc", base_out, mout.gridtile[m]);
    else snprintf(name_out[m], sizeof name_out[m], "%s.nc", base_out);
  }
  if (!format) { const int fmt = nc3_format(fin[0]); xgb_set_nc_format(fmt == 1 ? "classic" : fmt == 5 ? "cdf5" : "64bit_offset"); }  /* fregrid.c:896-903; netCDF-4
                                                   input (format 4): the output is 64-bit offset, HDF5 is read here but not written */

  Var var[MAXVAR];
  int nvar = 0;
  Axis axis[MAXAXIS];
  int naxis = 0, nt = 1, has_tavg_info = 0;
  memset(var, 0, sizeof var); memset(axis, 0, sizeof axis);
  for (int l = 0; l < nscalar; ++l) {
    const int vid = need_var(fin[0], name_in[0], scalar_name[l]);
    char att[STRING];
    if (nc3_get_att_text(fin[0], vid, "interp_method", att, sizeof att) == 0 && (!strcmp(att, "none") || !strcmp(att, "NONE") || !strcmp(att, "None")))
      continue;                                                                  /* fregrid.c:838-853 */
    Var *v = &var[nvar++];
    strcpy(v->name, scalar_name[l]);
    v->vid_in = vid;
    v->type = nc3_var_type(fin[0], vid);
    if (v->type != NC3_SHORT && v->type != NC3_INT && v->type != NC3_FLOAT && v->type != NC3_DOUBLE)
      die("fregrid_util(get_input_metadata): field %s in file %s has an invalid type, the type should be NC_DOUBLE, NC_FLOAT, NC_INT or NC_SHORT", v->name, name_in[0]);
    v->ndim = nc3_var_ndims(fin[0], vid);
    if (v->ndim > 5) die("get_input_metadata(fregrid_util.c): ndim should be no larger than 5");
    v->order = order; v->nz = 1; v->nn = 1; v->cell_methods = METHOD_MEAN;
    const int *dd = nc3_var_dimids(fin[0], vid);
    char cart[5], dname[5][STRING], bname[5][STRING];
    int dsize[5], dtype[5], dvid[5];
    for (int i = 0; i < v->ndim; ++i) {
      strcpy(dname[i], nc3_dim_name(fin[0], dd[i]));
      dsize[i] = (int)nc3_dim_len(fin[0], dd[i]);
      dvid[i] = nc3_var_id(fin[0], dname[i]);
      if (dvid[i] < 0) die("mpp_io(mpp_get_varid): error in get field_id of variable %s from file %s", dname[i], name_in[0]);
      cart[i] = var_cart(fin[0], dvid[i]);
      dtype[i] = nc3_var_type(fin[0], dvid[i]);
      var_bndname(fin[0], dvid[i], bname[i]);
    }
    v->do_regrid = v->ndim > 1 && cart[v->ndim - 1] == 'X' && cart[v->ndim - 2] == 'Y';
    if (!v->do_regrid) printf("fregrid_util: Field %s will not be remapped for file :\n  %s\n", v->name, name_in[0]);
    if (nc3_att_inq(fin[0], vid, "time_avg_info", NULL, NULL) == 0) has_tavg_info = 1;
    if (nc3_get_att_text(fin[0], vid, "cell_methods", att, sizeof att) == 0) {
      const char *p = strstr(att, "area:");
      if (p) { char w[STRING]; if (sscanf(p + 5, "%254s", w) == 1) {
        if (!strcmp(w, "mean")) v->cell_methods = METHOD_MEAN;
        else if (!strcmp(w, "sum")) v->cell_methods = METHOD_SUM;
        else die("fregrid_util(get_input_metadata): field %s in file %s attribute cell_methods should have value 'mean' or 'sum' after area: ", v->name, name_in[0]); } }
    }
    if (nc3_att_inq(fin[0], vid, "cell_measures", NULL, NULL) == 0 && v->do_regrid)
      printf("NOTE from fregrid_b200: cell_measures of field %s is ignored (associated area files are not read)\n", v->name);
    if (!(opcode & XGB_MONOTONIC) && nc3_get_att_text(fin[0], vid, "interp_method", att, sizeof att) == 0) {
      if (!strcmp(att, "conserve_order1")) v->order = 1;
      else if (!strcmp(att, "conserve_order2")) v->order = 2;
      else die("get_input_metadata(fregrid_util.c): in file %s, attribute interp_method of field %s has value = %s is not suitable, it should be conserve_order1, conserve_order2 or bilinear", name_in[0], v->name, att);
    }
    if (v->do_regrid) {
      if (dsize[v->ndim - 1] != gin[0].nx) die("get_input_metadata(fregrid_util.c): x-size in grid file in not the same as in data file");
      if (dsize[v->ndim - 2] != gin[0].ny) die("get_input_metadata(fregrid_util.c): y-size in grid file in not the same as in data file");
      if (v->ndim > 2) {
        if (cart[v->ndim - 3] == 'Z') {
          v->has_zaxis = 1; v->nz = dsize[v->ndim - 3];
          if (kend > v->nz) die("get_input_metadata(fregrid_util.c): KlevelEnd should be no larger than number of vertical levels of field %s in file %s.", v->name, name_in[0]);
          if (kbegin > 0) { v->kstart = kbegin - 1; v->kend = kend - 1; v->nz = kend - kbegin + 1; } else { v->kstart = 0; v->kend = v->nz - 1; }
        } else if (cart[v->ndim - 3] == 'N') { v->has_naxis = 1; v->nn = dsize[v->ndim - 3]; }
      }
      if (v->ndim > 3) {
        if (cart[v->ndim - 4] == 'Z') die("get_input_metadata(fregrid_util.c): the Z-axis must be the third dimension");
        if (cart[v->ndim - 4] == 'N') { v->has_naxis = 1; v->nn = dsize[v->ndim - 4]; }
      }
      if (cart[0] == 'T') {
        v->has_taxis = 1;
        if (lend > dsize[0]) die("get_input_metadata(fregrid_util.c): LstepEnd should be no larger than number of time levels of field %s in file %s.", v->name, name_in[0]);
        if (lbegin > 0) { v->lstart = lbegin - 1; nt = lend - lbegin + 1; } else { v->lstart = 0; nt = dsize[0]; }
      }
    }
    for (int i = 0; i < v->ndim; ++i) {
      int j;
      for (j = 0; j < naxis; ++j) if (!strcmp(dname[i], axis[j].name)) break;
      v->index[i] = j;
      if (j < naxis) continue;
      if (naxis >= MAXAXIS) die("get_input_metadata(fregrid_util.c):ndim is greater than MAXDIM");
      Axis *a = &axis[naxis++];
      a->cart = cart[i]; a->type = dtype[i]; a->vid_in = dvid[i];
      strcpy(a->name, dname[i]); strcpy(a->bndname, bname[i]);
      size_t st[2] = {0, 0}, cn[2] = {0, 2};
      if (v->do_regrid && cart[i] == 'T') { st[0] = (size_t)v->lstart; a->size = nt; }
      else if (v->do_regrid && cart[i] == 'Z') { st[0] = (size_t)v->kstart; a->size = v->nz; }
      else a->size = dsize[i];
      a->data = xmalloc((size_t)(a->size + 1) * sizeof(double));
      cn[0] = (size_t)a->size;
      if (nc3_get_vara_double(fin[0], a->vid_in, st, cn, a->data)) die("%s: %s", name_in[0], nc3_strerror(fin[0]));
      a->bndtype = 0;
      if (strcmp(a->bndname, "none")) {
        a->bnd_vid_in = need_var(fin[0], name_in[0], a->bndname);
        if (nc3_var_ndims(fin[0], a->bnd_vid_in) == 1) {
          a->bndtype = 1; a->bnddata = xmalloc((size_t)(a->size + 1) * sizeof(double)); cn[0] = (size_t)a->size + 1;
        } else { a->bndtype = 2; a->bnddata = xmalloc((size_t)2 * a->size * sizeof(double)); cn[0] = (size_t)a->size; cn[1] = 2; }
        if (nc3_get_vara_double(fin[0], a->bnd_vid_in, st, cn, a->bnddata)) die("%s: %s", name_in[0], nc3_strerror(fin[0]));
      } else if (a->cart == 'X' || a->cart == 'Y') { char nm[STRING + 8]; snprintf(nm, sizeof nm, "%s_bnds", dname[i]); strncpy(a->bndname, nm, STRING - 1); }   /* fregrid_util.c:1333 */
    }
    /* get_field_attribute */
    v->missing = 0; v->scale = 0; v->offset = 0;
    if (v->do_regrid) {
      v->has_missing = nc3_get_att_double(fin[0], vid, "missing_value", &v->missing, 1) == 1;
      if (!v->has_missing) v->has_missing = nc3_get_att_double(fin[0], vid, "_FillValue", &v->missing, 1) == 1;
      nc3_get_att_double(fin[0], vid, "scale_factor", &v->scale, 1);
      nc3_get_att_double(fin[0], vid, "add_offset", &v->offset, 1);
    }
  }
  if (nvar == 0) { printf("NOTE from fregrid: no scalar and vector field need to be regridded.\n"); return 0; }
  int need2 = 0;
  for (int l = 0; l < nvar; ++l) if (var[l].do_regrid && var[l].order == 2) need2 = 1;
  if (need2 && order == 1)
    die("fregrid_b200: a field asks for conserve_order2 through its interp_method attribute; rerun with --interp_method conserve_order2");
  (void)has_tavg_info;

  /* weight field (set_weight_inf, fregrid_util.c:126-150) */
  double *weight = NULL;
  if (weight_field) {
    weight = xmalloc(ncell * sizeof(double));
    size_t o = 0;
    char s[STRING]; strcpy(s, weight_file); size_t l = strlen(s); if (l > 3 && !strcmp(s + l - 3, ".nc")) s[l - 3] = 0;
    for (int n = 0; n < ntiles_in; ++n) {
      char p[3 * STRING];
      if (ntiles_in > 1) snprintf(p, sizeof p, "%s.%s.nc", s, min.gridtile[n]); else snprintf(p, sizeof p, "%s.nc", s);
      nc3_file *w = nc_open_or_die(p);
      const int wv = need_var(w, p, weight_field);
      const int nd = nc3_var_ndims(w, wv);
      size_t st[5] = {0, 0, 0, 0, 0}, cn[5] = {1, 1, 1, 1, 1};
      if (nd < 2) die("fregrid_util(set_weight_inf): weight field %s should have at least two dimensions", weight_field);
      cn[nd - 2] = (size_t)gin[n].ny; cn[nd - 1] = (size_t)gin[n].nx;
      if (nc3_get_vara_double(w, wv, st, cn, weight + o)) die("%s: %s", p, nc3_strerror(w));
      o += (size_t)gin[n].nx * gin[n].ny;
      nc3_close(w);
    }
  }

  /* ---- remap, one output tile at a time ---- */
  double *lont_cat = NULL, *latt_cat = NULL;
  if (need2) {
    lont_cat = xmalloc(nhalo * sizeof(double)); latt_cat = xmalloc(nhalo * sizeof(double));
    size_t o = 0;
    for (int n = 0; n < ntiles_in; ++n) { const size_t k = (size_t)(gin[n].nx + 2) * (gin[n].ny + 2);
      memcpy(lont_cat + o, gin[n].lont, k * sizeof(double)); memcpy(latt_cat + o, gin[n].latt, k * sizeof(double)); o += k; }
  }
  for (int m = 0; m < ntiles_out; ++m) {
    const int nx2 = gout[m].nx, ny2 = gout[m].ny;
    /* dst_is_latlon (fregrid_util.c:1489-1515) */
    int dst_is_latlon = 1;
    for (int j = 0; j <= ny2 && dst_is_latlon; ++j) for (int i = 1; i <= nx2; ++i)
      if (gout[m].latc[(size_t)j * (nx2 + 1) + i] != gout[m].latc[(size_t)j * (nx2 + 1)]) { dst_is_latlon = 0; break; }
    for (int i = 0; i <= nx2 && dst_is_latlon; ++i) for (int j = 1; j <= ny2; ++j)
      if (gout[m].lonc[(size_t)j * (nx2 + 1) + i] != gout[m].lonc[i]) { dst_is_latlon = 0; break; }
    if (!gout[m].lonc1D) {                                  /* mosaic output: the 1-D axes are the first row / column */
      gout[m].lonc1D = xmalloc((nx2 + 1) * sizeof(double)); gout[m].latc1D = xmalloc((ny2 + 1) * sizeof(double));
      for (int i = 0; i <= nx2; ++i) gout[m].lonc1D[i] = gout[m].lonc[i];
      for (int j = 0; j <= ny2; ++j) gout[m].latc1D[j] = gout[m].latc[(size_t)j * (nx2 + 1)];
    }
    /* define the output file (set_output_metadata) */
    char err[400];
    nc3_file *fo = nc3_create(name_out[m], xgb_get_nc_format(), err, sizeof err);
    if (!fo) die("mpp_io(mpp_open): error in opening file %s: %s", name_out[m], err);
    nc3_copy_atts(fin[0], NC3_GLOBAL, fo, NC3_GLOBAL);
    { char host[128] = "", tbuf[64]; time_t now = time(NULL); gethostname(host, sizeof host);
      strncpy(tbuf, ctime(&now), sizeof tbuf - 1); tbuf[sizeof tbuf - 1] = 0; if (strchr(tbuf, '\n')) *strchr(tbuf, '\n') = 0;
      nc3_put_att_text(fo, NC3_GLOBAL, "code_release_version", "fregrid_b200 (libxgrid_b200)");
      nc3_put_att_text(fo, NC3_GLOBAL, "creationtime", tbuf);
      nc3_put_att_text(fo, NC3_GLOBAL, "hostname", host);
      nc3_put_att_text(fo, NC3_GLOBAL, "history", history); }
    int dim_bnds = -1, have_bnds_axis = 0;
    for (int i = 0; i < naxis; ++i) {
      Axis *a = &axis[i];
      if (a->cart == 'X') a->size = nx2;
      if (a->cart == 'Y') a->size = ny2;
      if (standard_dimension && (a->cart == 'X' || a->cart == 'Y')) a->bndtype = 3;
      if (a->bndtype == 0 && (a->cart == 'X' || a->cart == 'Y') && dst_is_latlon) a->bndtype = 3;
      if (!strcmp(a->name, "bnds")) have_bnds_axis = 1;
    }
    if (!have_bnds_axis)
      for (int i = 0; i < naxis; ++i)
        if (axis[i].bndtype == 2 || axis[i].bndtype == 3 || (axis[i].bndtype == 1 && standard_dimension)) { dim_bnds = nc3_def_dim(fo, "bnds", 2); break; }
    for (int i = 0; i < naxis; ++i) {
      Axis *a = &axis[i];
      if (a->cart == 'T') a->dimid = nc3_def_dim(fo, a->name, 0);
      else if ((a->type == NC3_INT || standard_dimension) && a->cart == 'X') a->dimid = nc3_def_dim(fo, "lon", a->size);
      else if ((a->type == NC3_INT || standard_dimension) && a->cart == 'Y') a->dimid = nc3_def_dim(fo, "lat", a->size);
      else { a->dimid = nc3_def_dim(fo, a->name, a->size); if (!strcmp(a->name, "bnds")) dim_bnds = a->dimid; }
      if (a->dimid < 0) die("%s: %s", name_out[m], nc3_strerror(fo));
    }
    for (int i = 0; i < naxis; ++i) {
      Axis *a = &axis[i];
      int dims[2] = {a->dimid, dim_bnds};
      const int xy = (a->cart == 'X' || a->cart == 'Y');
      const int isx = a->cart == 'X';
      if (xy && (a->type == NC3_INT || standard_dimension)) {
        a->vid = nc3_def_var(fo, isx ? "lon" : "lat", NC3_DOUBLE, 1, &a->dimid);
        if (a->type == NC3_INT) {
          nc3_put_att_text(fo, a->vid, "units", "degrees"); nc3_put_att_text(fo, a->vid, "axis", isx ? "X" : "Y");
          nc3_put_att_text(fo, a->vid, "standard_name", isx ? "grid_longitude" : "grid_latitude");
          nc3_put_att_text(fo, a->vid, "bounds", isx ? "lon_bnds" : "lat_bnds");
          a->bndid = nc3_def_var(fo, isx ? "lon_bnds" : "lat_bnds", NC3_DOUBLE, 2, dims);
          nc3_put_att_text(fo, a->bndid, "units", "degrees");
          nc3_put_att_text(fo, a->bndid, "standard_name", isx ? "grid_longitude_bounds" : "grid_latitude_bounds");
        } else {
          nc3_put_att_text(fo, a->vid, "long_name", isx ? "longitude" : "latitude");
          nc3_put_att_text(fo, a->vid, "units", isx ? "degrees_E" : "degrees_N");
          nc3_put_att_text(fo, a->vid, "axis", isx ? "X" : "Y");
          nc3_put_att_text(fo, a->vid, "bounds", isx ? "lon_bnds" : "lat_bnds");
          a->bndid = nc3_def_var(fo, isx ? "lon_bnds" : "lat_bnds", NC3_DOUBLE, 2, dims);
          nc3_put_att_text(fo, a->bndid, "long_name", isx ? "longitude bounds" : "latitude bounds");
          nc3_put_att_text(fo, a->bndid, "units", isx ? "degrees_E" : "degrees_N");
          nc3_put_att_text(fo, a->bndid, "axis", isx ? "X" : "Y");
        }
      } else {
        a->vid = nc3_def_var(fo, a->name, a->type, 1, &a->dimid);
        nc3_copy_atts(fin[0], a->vid_in, fo, a->vid);
        if (a->bndtype == 3) {
          nc3_put_att_text(fo, a->vid, "bounds", a->bndname);
          a->bndid = nc3_def_var(fo, a->bndname, a->type, 2, dims);
          nc3_copy_atts(fin[0], a->vid_in, fo, a->bndid);
        } else if (a->bndtype == 1) {
          int j, d1;
          for (j = 0; j < naxis; ++j) if (!strcmp(a->bndname, axis[j].name)) break;
          if (j < naxis) a->bndtype = 0;                    /* the bounds axis is a dimension of its own, written as such */
          else { d1 = nc3_def_dim(fo, a->bndname, a->size + 1); a->bndid = nc3_def_var(fo, a->bndname, a->type, 1, &d1);
                 nc3_copy_atts(fin[0], a->bnd_vid_in, fo, a->bndid); }
        } else if (a->bndtype == 2) {
          a->bndid = nc3_def_var(fo, a->bndname, a->type, 2, dims);
          nc3_copy_atts(fin[0], a->bnd_vid_in, fo, a->bndid);
        }
      }
      if (a->vid < 0) die("%s: %s", name_out[m], nc3_strerror(fo));
    }
    for (int l = 0; l < nvar; ++l) {
      Var *v = &var[l];
      int dims[5], isaxis = 0;
      for (int j = 0; j < naxis; ++j) if (!strcmp(axis[j].name, v->name)) isaxis = 1;
      if (isaxis) { v->vid_out = -1; continue; }
      for (int i = 0; i < v->ndim; ++i) dims[i] = axis[v->index[i]].dimid;
      v->vid_out = nc3_def_var(fo, v->name, v->type, v->ndim, dims);
      if (v->vid_out < 0) die("%s: %s", name_out[m], nc3_strerror(fo));
      const int natts = nc3_var_natts(fin[0], v->vid_in);
      for (int i = 0; i < natts; ++i) {
        const char *an = nc3_att_name(fin[0], v->vid_in, i);
        if (!strcmp(an, "time_avg_info")) continue;
        if (standard_dimension && dst_is_latlon && !strcmp(an, "coordinates")) continue;
        nc3_copy_att(fin[0], v->vid_in, an, fo, v->vid_out);
      }
      if (v->do_regrid) nc3_put_att_text(fo, v->vid_out, "interp_method", v->order == 2 ? "conserve_order2" : "conserve_order1");
    }
    if (nc3_enddef(fo)) die("%s: %s", name_out[m], nc3_strerror(fo));
    for (int i = 0; i < naxis; ++i) {
      Axis *a = &axis[i];
      if (a->cart == 'T') continue;
      double *d = a->data, *b = a->bnddata, *tmpd = NULL, *tmpb = NULL;
      if (a->cart == 'X' || a->cart == 'Y') {
        const int isx = a->cart == 'X';
        const double *t1 = isx ? gout[m].lont1D : gout[m].latt1D, *c1 = isx ? gout[m].lonc1D : gout[m].latc1D;
        tmpd = xmalloc(a->size * sizeof(double)); tmpb = xmalloc((size_t)2 * a->size * sizeof(double) + 16);
        for (int k = 0; k < a->size; ++k) tmpd[k] = t1[k] * R2D;
        if (a->bndtype == 1) for (int k = 0; k <= a->size; ++k) tmpb[k] = c1[k] * R2D;
        else for (int k = 0; k < a->size; ++k) { tmpb[2 * k] = c1[k] * R2D; tmpb[2 * k + 1] = c1[k + 1] * R2D; }
        d = tmpd; b = tmpb;
      }
      if (nc3_put_var_double(fo, a->vid, d)) die("%s: %s", name_out[m], nc3_strerror(fo));
      if (a->bndtype > 0 && nc3_put_var_double(fo, a->bndid, b)) die("%s: %s", name_out[m], nc3_strerror(fo));
      free(tmpd); free(tmpb);
    }
    for (int l = 0; l < nvar; ++l) {                        /* fields that are not remapped are copied (mpp_copy_data) */
      Var *v = &var[l];
      if (v->do_regrid || v->vid_out < 0) continue;
      size_t tot = 1; const int *dd = nc3_var_dimids(fin[0], v->vid_in);
      for (int i = 0; i < v->ndim; ++i) tot *= (size_t)nc3_dim_len(fin[0], dd[i]);
      double *buf = xmalloc(tot * sizeof(double));
      if (nc3_get_var_double(fin[0], v->vid_in, buf)) die("%s: %s", name_in[0], nc3_strerror(fin[0]));
      size_t st[5] = {0, 0, 0, 0, 0}, cn[5];
      for (int i = 0; i < v->ndim; ++i) cn[i] = (size_t)nc3_dim_len(fin[0], dd[i]);
      if (nc3_put_vara_double(fo, v->vid_out, st, cn, buf)) die("%s: %s", name_out[m], nc3_strerror(fo));
      free(buf);
    }

    double *dst_area = NULL;
    if (target_grid) { dst_area = xmalloc((size_t)nx2 * ny2 * sizeof(double));
      XGB(set_dst_tile(plan, &gout[m], !mosaic_out, nlon, nlat, lonbegin, lonend, latbegin, latend)); XGB(xgb_plan_dst_area_host(plan, dst_area)); }
    /* the exchange grid of this output tile on the device */
    if (xg[m].n > 0) {
      XGB(xgb_plan_set_xgrid(plan, ntiles_in, nxs, nys, nx2, ny2, xg[m].n, xg[m].t_in, xg[m].i_in, xg[m].j_in, xg[m].i_out, xg[m].j_out,
                             xg[m].area, xg[m].di, xg[m].dj, 0));
      if (need2) XGB(xgb_plan_grad_setup(plan, lont_cat, latt_cat, 0));
    }

    for (int t = 0; t < nt; ++t) {
      for (int i = 0; i < naxis; ++i) {                     /* write_output_time */
        Axis *a = &axis[i];
        if (a->cart != 'T') continue;
        size_t st[2] = {(size_t)t, 0}, cn[2] = {1, 2};
        if (nc3_put_vara_double(fo, a->vid, st, cn, &a->data[t])) die("%s: %s", name_out[m], nc3_strerror(fo));
        if (a->bndtype == 2 && nc3_put_vara_double(fo, a->bndid, st, cn, &a->bnddata[2 * t])) die("%s: %s", name_out[m], nc3_strerror(fo));
      }
      for (int l = 0; l < nvar; ++l) {
        Var *v = &var[l];
        if (!v->do_regrid || v->vid_out < 0) continue;
        if (!v->has_taxis && t > 0) continue;
        const int level_t = t + v->lstart, nz = v->nz, o2 = v->order == 2;
        const size_t per = o2 ? nhalo : ncell;
        double *data = xmalloc(per * nz * sizeof(double)), *out = xmalloc((size_t)nx2 * ny2 * nz * sizeof(double));
        for (int ln = 0; ln < v->nn; ++ln) {
          size_t st[5] = {0, 0, 0, 0, 0}, cn[5] = {1, 1, 1, 1, 1};
          int pos = 0;
          if (v->has_taxis) st[pos++] = (size_t)level_t;
          if (v->has_naxis) st[pos++] = (size_t)ln;
          if (v->has_zaxis) { st[pos] = (size_t)v->kstart; cn[pos++] = (size_t)nz; }
          if (v->ndim != pos + 2) die("fregrid_util(get_input_data): mimstch between ndim and has_taxis/has_zaxis/has_naxis");
          /* get_input_data: read every tile, scale / offset, place inside the halo */
          size_t off = 0;
          for (int n = 0; n < ntiles_in; ++n) {
            const int nx = gin[n].nx, ny = gin[n].ny;
            const size_t nc = (size_t)nx * ny, nh = (size_t)(nx + 2) * (ny + 2), nt_ = o2 ? nh : nc;
            double *raw = xmalloc(nc * nz * sizeof(double));
            cn[pos] = (size_t)ny; cn[pos + 1] = (size_t)nx;
            if (nc3_get_vara_double(fin[n], need_var(fin[n], name_in[n], v->name), st, cn, raw)) die("%s: %s", name_in[n], nc3_strerror(fin[n]));
            if (v->scale != 0) for (size_t i = 0; i < nc * nz; ++i) if (raw[i] != v->missing) raw[i] *= v->scale;
            if (v->offset != 0) for (size_t i = 0; i < nc * nz; ++i) if (raw[i] != v->missing) raw[i] += v->offset;
            for (int k = 0; k < nz; ++k) {
              double *dstp = data + (size_t)k * per + off;
              if (!o2) memcpy(dstp, raw + (size_t)k * nc, nc * sizeof(double));
              else {
                init_halo(dstp, nx, ny);
                for (int j = 0; j < ny; ++j) memcpy(dstp + (size_t)(j + 1) * (nx + 2) + 1, raw + (size_t)k * nc + (size_t)j * nx, nx * sizeof(double));
              }
            }
            free(raw);
            off += nt_;
          }
          if (o2)
            for (int k = 0; k < nz; ++k) {
              double *tiles[MAXTILE];
              size_t o = 0;
              for (int n = 0; n < ntiles_in; ++n) { tiles[n] = data + (size_t)k * per + o; o += (size_t)(gin[n].nx + 2) * (gin[n].ny + 2); }
              for (int n = 0; n < ntiles_in; ++n) update_halo(gin, n, tiles, &bound[n]);
            }
          const double miss = v->has_missing ? v->missing : -1.e20;
          if (xg[m].n == 0) for (size_t i = 0; i < (size_t)nx2 * ny2 * nz; ++i) out[i] = miss;   /* conserve_interp.c:541-560 */
          else {
            XGB(xgb_plan_apply_options(plan, v->cell_methods == METHOD_SUM, weight, NULL, NULL, 0.0, target_grid, dst_area, 0));
            unsigned op = o2 ? XGB_CONSERVE_ORDER2 : XGB_CONSERVE_ORDER1;
            if (o2 && (opcode & XGB_MONOTONIC)) op |= XGB_MONOTONIC;
            if (o2) XGB(xgb_plan_regrid(plan, op, nz, data, v->has_missing, v->missing, out, 0));
            else XGB(xgb_plan_apply(plan, op, nz, data, NULL, NULL, NULL, v->has_missing, v->missing, out, 0));
          }
          /* write_field_data */
          const size_t no = (size_t)nx2 * ny2 * nz;
          if (v->offset != 0) for (size_t i = 0; i < no; ++i) if (out[i] != v->missing) out[i] -= v->offset;
          if (v->scale != 0) for (size_t i = 0; i < no; ++i) if (out[i] != v->missing) out[i] /= v->scale;
          if (v->type == NC3_SHORT) for (size_t i = 0; i < no; ++i) out[i] = (double)(short)out[i];
          if (v->type == NC3_INT) for (size_t i = 0; i < no; ++i) out[i] = (double)(int)out[i];
          pos = 0;
          size_t so[5] = {0, 0, 0, 0, 0}, co[5] = {1, 1, 1, 1, 1};
          if (v->has_taxis) so[pos++] = (size_t)t;
          if (v->has_naxis) so[pos++] = (size_t)ln;
          if (v->has_zaxis) co[pos++] = (size_t)nz;
          co[pos] = (size_t)ny2; co[pos + 1] = (size_t)nx2;
          if (nc3_put_vara_double(fo, v->vid_out, so, co, out)) die("%s: %s", name_out[m], nc3_strerror(fo));
        }
        free(data); free(out);
      }
    }
    if (nc3_close(fo)) die("fregrid: error closing %s", name_out[m]);
    free(dst_area);
  }
  printf("Successfully running fregrid and the following output file are generated.\n");
  for (int m = 0; m < ntiles_out; ++m) printf("****%s\n", name_out[m]);
  for (int n = 0; n < ntiles_in; ++n) nc3_close(fin[n]);
  xgb_plan_destroy(plan);
  return 0;
}
