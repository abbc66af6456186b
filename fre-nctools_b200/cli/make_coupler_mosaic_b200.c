/* make_coupler_mosaic_b200 — the make_coupler_mosaic command line (tools/make_coupler_mosaic/make_coupler_mosaic.c) with its
 * exchange-grid loops on the GPU.
 *
 * Host side in plain C, like the reference tool: the option table (make_coupler_mosaic.c:409-425), the mosaic / supergrid /
 * topography readers (:515-960: model grid = every second supergrid point, degrees -> radians, the artificial southern ocean
 * row, omask from `area_frac` or `depth > sea_level`), the exchange-grid files (:2131-2480, :2810-2960), land_mask / ocean_mask
 * (:2020-2120) and the coupler mosaic file (:3664-3805).  Between reading and writing sits ONE call,
 * xgb_make_coupler_xgrid (csrc/coupler.cu), instead of the loops of :1250-2017 and :2556-2808.
 *
 * Built: the tool's default clip method.  Refused by name: --wave_mosaic, great-circle grids (great_circle_algorithm = TRUE in a
 * grid file), nested atmosphere mosaics, --rotate_poly, netCDF-4 output.  --check, --verbose and --print_memory are accepted and
 * ignored (they only print).
 */
#include <getopt.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/xgrid_b200.h"
#include "../csrc/nc3.h"

#define STRING 255
#define D2R (M_PI / 180.)
#define TINY_VALUE (1.e-7)                       /* make_coupler_mosaic.c:146 */
#define TOLORENCE (1.e-4)                        /* :147 */
static const char grid_version[] = "0.2";        /* :153 */

static void die(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
static void die(const char *fmt, ...)
{
  va_list ap;
  fprintf(stderr, "FATAL Error: ");              /* mpp_error, mpp.c:290-298 */
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fprintf(stderr, "\n");
  exit(1);
}
static void *xmalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) die("make_coupler_mosaic: out of memory"); return p; }

static nc3_file *open_or_die(const char *path)
{
  char err[400];
  nc3_file *f = nc3_open(path, err, sizeof err);
  if (!f) die("mpp_io(mpp_open): error in opening file %s: %s", path, err);
  return f;
}
static long long need_dim(nc3_file *f, const char *file, const char *name)
{
  const int d = nc3_dim_id(f, name);
  if (d < 0) die("mpp_io(mpp_get_dimlen): error in inquiring dimension %s from file %s", name, file);
  return nc3_dim_len(f, d);
}
static int need_var(nc3_file *f, const char *file, const char *name)
{
  const int v = nc3_var_id(f, name);
  if (v < 0) die("mpp_io(mpp_get_varid): error in get field_id of variable %s from file %s", name, file);
  return v;
}
/* one row of a (n, string) character variable, or a (string) variable */
static void read_string(nc3_file *f, const char *file, const char *var, int row, char *out)
{
  const int v = need_var(f, file, var);
  const int nd = nc3_var_ndims(f, v);
  const int *dd = nc3_var_dimids(f, v);
  size_t start[2] = {(size_t)row, 0}, count[2] = {1, 0};
  char buf[1024];
  long long len = nc3_dim_len(f, dd[nd - 1]);
  if (len >= (long long)sizeof buf) die("make_coupler_mosaic: string variable %s in %s is too long", var, file);
  memset(buf, 0, sizeof buf);
  if (nd == 1) { start[0] = 0; count[0] = (size_t)len; } else count[1] = (size_t)len;
  if (nc3_get_vara_text(f, v, start, count, buf)) die("%s: %s", file, nc3_strerror(f));
  strncpy(out, buf, STRING); out[STRING - 1] = 0;
}
static void dir_and_name(const char *file, char *dir, char *name)      /* get_file_dir_and_name, make_coupler_mosaic.c:157-174 */
{
  const char *s = strrchr(file, '/');
  if (!s) { strcpy(name, file); strcpy(dir, "./"); }
  else { const size_t n = (size_t)(s + 1 - file); strcpy(name, s + 1); memcpy(dir, file, n); dir[n] = 0; }
}
static void path_of(const char *file, char *dir)                       /* get_file_path */
{
  const char *s = strrchr(file, '/');
  if (!s) strcpy(dir, ".");
  else { const size_t n = (size_t)(s - file); memcpy(dir, file, n); dir[n] = 0; if (n == 0) strcpy(dir, "/"); }
}

typedef struct {
  char name[STRING];
  int ntiles;
  char (*tile)[STRING];
  int *nx, *ny;                 /* model cells per tile */
  double **x, **y;              /* model-grid vertices per tile, radians */
} Mosaic;

/* read one mosaic: its name, tiles and the tiles' model grids (:515-600, :662-720, :776-880) */
static void load_mosaic(const char *path, Mosaic *m, const char *what)
{
  char dir[STRING], gridfile[STRING], file[2 * STRING + 2];
  nc3_file *f = open_or_die(path);
  read_string(f, path, "mosaic", 0, m->name);
  m->ntiles = (int)need_dim(f, path, "ntiles");
  m->tile = xmalloc((size_t)m->ntiles * sizeof *m->tile);
  m->nx = xmalloc((size_t)m->ntiles * sizeof(int)); m->ny = xmalloc((size_t)m->ntiles * sizeof(int));
  m->x = xmalloc((size_t)m->ntiles * sizeof(double *)); m->y = xmalloc((size_t)m->ntiles * sizeof(double *));
  path_of(path, dir);
  for (int n = 0; n < m->ntiles; ++n) {
    read_string(f, path, "gridfiles", n, gridfile);
    read_string(f, path, "gridtiles", n, m->tile[n]);
    snprintf(file, sizeof file, "%s/%s", dir, gridfile);
    nc3_file *g = open_or_die(file);
    const long long snx = need_dim(g, file, "nx"), sny = need_dim(g, file, "ny");
    {                                                                   /* get_great_circle_algorithm */
      const int tv = nc3_var_id(g, "tile");
      char att[64] = "";
      if (tv >= 0 && nc3_get_att_text(g, tv, "great_circle_algorithm", att, sizeof att) == 0 && !strcmp(att, "TRUE"))
        die("make_coupler_mosaic_b200: %s asks for the great-circle algorithm; only the default clip method is built", file);
    }
    if (snx % 2) die("make_coupler_mosaic: %s supergrid x-size can not be divided by x_refine", what);
    if (sny % 2) die("make_coupler_mosaic: %s supergrid y-size can not be divided by y_refine", what);
    const int nx = (int)(snx / 2), ny = (int)(sny / 2);
    double *sx = xmalloc((size_t)(snx + 1) * (sny + 1) * sizeof(double)), *sy = xmalloc((size_t)(snx + 1) * (sny + 1) * sizeof(double));
    if (nc3_get_var_double(g, need_var(g, file, "x"), sx) || nc3_get_var_double(g, need_var(g, file, "y"), sy))
      die("%s: %s", file, nc3_strerror(g));
    nc3_close(g);
    m->nx[n] = nx; m->ny[n] = ny;
    m->x[n] = xmalloc((size_t)(nx + 1) * (ny + 1) * sizeof(double));
    m->y[n] = xmalloc((size_t)(nx + 1) * (ny + 1) * sizeof(double));
    for (int j = 0; j <= ny; ++j)
      for (int i = 0; i <= nx; ++i) {                                   /* get_global_grid :177-222, then * D2R :580-583 */
        m->x[n][j * (nx + 1) + i] = sx[(size_t)(2 * j) * (snx + 1) + 2 * i] * D2R;
        m->y[n][j * (nx + 1) + i] = sy[(size_t)(2 * j) * (snx + 1) + 2 * i] * D2R;
      }
    free(sx); free(sy);
  }
  nc3_close(f);
}

static void concat(const Mosaic *m, double **lon, double **lat, xgb_mosaic_grid *g)
{
  size_t nv = 0, at = 0;
  for (int n = 0; n < m->ntiles; ++n) nv += (size_t)(m->nx[n] + 1) * (m->ny[n] + 1);
  *lon = xmalloc(nv * sizeof(double)); *lat = xmalloc(nv * sizeof(double));
  for (int n = 0; n < m->ntiles; ++n) {
    const size_t k = (size_t)(m->nx[n] + 1) * (m->ny[n] + 1);
    memcpy(*lon + at, m->x[n], k * sizeof(double)); memcpy(*lat + at, m->y[n], k * sizeof(double));
    at += k;
  }
  g->ntiles = m->ntiles; g->nx = m->nx; g->ny = m->ny; g->lon = *lon; g->lat = *lat;
}

/* ---- writers ------------------------------------------------------------------------------------------------------ */
static char g_history[1024];
static nc3_file *create_or_die(const char *path)
{
  char err[400], host[128] = "", *t;
  time_t now;
  nc3_file *f = nc3_create(path, 2, err, sizeof err);
  if (!f) die("mpp_io(mpp_open): error in creating file %s: %s", path, err);
  gethostname(host, sizeof host);
  time(&now);
  t = strtok(ctime(&now), "\n");
  /* print_provenance_gv_gca, tool_util.c:780-808 */
  nc3_put_att_text(f, NC3_GLOBAL, "grid_version", grid_version);
  nc3_put_att_text(f, NC3_GLOBAL, "code_release_version", "b200");
  nc3_put_att_text(f, NC3_GLOBAL, "git_hash", "unknown");
  nc3_put_att_text(f, NC3_GLOBAL, "creationtime", t ? t : "");
  nc3_put_att_text(f, NC3_GLOBAL, "hostname", host);
  nc3_put_att_text(f, NC3_GLOBAL, "history", g_history);
  return f;
}
static int def_text_var(nc3_file *f, const char *name, int ndims, const int *dims, int npairs, ...)
{
  va_list ap;
  const int v = nc3_def_var(f, name, NC3_CHAR, ndims, dims);
  va_start(ap, npairs);
  for (int k = 0; k < npairs; ++k) { const char *a = va_arg(ap, const char *), *b = va_arg(ap, const char *); nc3_put_att_text(f, v, a, b); }
  va_end(ap);
  return v;
}
static void put_text(nc3_file *f, int v, int row, int ndims, const char *s)
{
  size_t start[2] = {0, 0}, count[2] = {1, 1};
  if (!*s) return;
  if (ndims == 1) count[0] = strlen(s); else { start[0] = (size_t)row; count[1] = strlen(s); }
  if (nc3_put_vara_text(f, v, start, count, s)) die("make_coupler_mosaic: write failed: %s", nc3_strerror(f));
}

/* one exchange-grid file (:2131-2258 atmXlnd, :2260-2385 atmXocn, :2823-2950 lndXocn): the (t1, t2) sub-list of l */
static void write_xgrid(const char *file, const char *contact, int order, const xgb_coupler_list *l, int t1, int t2, int jshift2,
                        const char *sn1, const char *sn2)
{
  long long n = 0, k = 0;
  for (long long e = 0; e < l->n; ++e) if (l->t1[e] == t1 && l->t2[e] == t2) ++n;
  nc3_file *f = create_or_die(file);
  int dims[2];
  const int d_string = nc3_def_dim(f, "string", STRING), d_ncells = nc3_def_dim(f, "ncells", n), d_two = nc3_def_dim(f, "two", 2);
  int v_contact, v_c1, v_c2, v_area, v_d1 = -1, v_d2 = -1;
  if (order == 2)
    v_contact = def_text_var(f, "contact", 1, &d_string, 7, "standard_name", "grid_contact_spec", "contact_type", "exchange",
                             "parent1_cell", "tile1_cell", "parent2_cell", "tile2_cell", "xgrid_area_field", "xgrid_area",
                             "distant_to_parent1_centroid", "tile1_distance", "distant_to_parent2_centroid", "tile2_distance");
  else
    v_contact = def_text_var(f, "contact", 1, &d_string, 5, "standard_name", "grid_contact_spec", "contact_type", "exchange",
                             "parent1_cell", "tile1_cell", "parent2_cell", "tile2_cell", "xgrid_area_field", "xgrid_area");
  dims[0] = d_ncells; dims[1] = d_two;
  v_c1 = nc3_def_var(f, "tile1_cell", NC3_INT, 2, dims); nc3_put_att_text(f, v_c1, "standard_name", sn1);
  v_c2 = nc3_def_var(f, "tile2_cell", NC3_INT, 2, dims); nc3_put_att_text(f, v_c2, "standard_name", sn2);
  v_area = nc3_def_var(f, "xgrid_area", NC3_DOUBLE, 1, &d_ncells);
  nc3_put_att_text(f, v_area, "standard_name", "exchange_grid_area"); nc3_put_att_text(f, v_area, "units", "m2");
  if (order == 2) {
    v_d1 = nc3_def_var(f, "tile1_distance", NC3_DOUBLE, 2, dims);
    nc3_put_att_text(f, v_d1, "standard_name", "distance_from_parent1_cell_centroid");
    v_d2 = nc3_def_var(f, "tile2_distance", NC3_DOUBLE, 2, dims);
    nc3_put_att_text(f, v_d2, "standard_name", "distance_from_parent2_cell_centroid");
  }
  if (nc3_enddef(f)) die("%s: %s", file, nc3_strerror(f));
  put_text(f, v_contact, 0, 1, contact);
  int *c1 = xmalloc((size_t)n * 2 * sizeof(int)), *c2 = xmalloc((size_t)n * 2 * sizeof(int));
  double *a = xmalloc((size_t)n * sizeof(double)), *d1 = xmalloc((size_t)n * 2 * sizeof(double)), *d2 = xmalloc((size_t)n * 2 * sizeof(double));
  for (long long e = 0; e < l->n; ++e) {
    if (l->t1[e] != t1 || l->t2[e] != t2) continue;
    c1[2 * k] = l->i1[e] + 1; c1[2 * k + 1] = l->j1[e] + 1;             /* 1-based on disk (:2186-2191) */
    c2[2 * k] = l->i2[e] + 1; c2[2 * k + 1] = l->j2[e] + 1 - jshift2;    /* :2309: the artificial ocean row is not counted */
    a[k] = l->area[e];
    if (order == 2) { d1[2 * k] = l->d1i[e]; d1[2 * k + 1] = l->d1j[e]; d2[2 * k] = l->d2i[e]; d2[2 * k + 1] = l->d2j[e]; }
    ++k;
  }
  if (nc3_put_var_int(f, v_c1, c1) || nc3_put_var_int(f, v_c2, c2) || nc3_put_var_double(f, v_area, a) ||
      (order == 2 && (nc3_put_var_double(f, v_d1, d1) || nc3_put_var_double(f, v_d2, d2))))
    die("%s: %s", file, nc3_strerror(f));
  if (nc3_close(f)) die("make_coupler_mosaic: cannot close %s", file);
  free(c1); free(c2); free(a); free(d1); free(d2);
}

static int def_2d(nc3_file *f, const char *name, const int *dims, const char *sn)
{
  const int v = nc3_def_var(f, name, NC3_DOUBLE, 2, dims);
  nc3_put_att_text(f, v, "standard_name", sn); nc3_put_att_text(f, v, "units", "none");
  return v;
}

static const char *usage =
    "make_coupler_mosaic_b200 --atmos_mosaic atmos_mosaic.nc --ocean_mosaic ocean_mosaic.nc --ocean_topog ocean_topog.nc\n"
    "        [--land_mosaic land_mosaic.nc] [--sea_level #] [--interp_order #] [--mosaic_name mosaic_name]\n"
    "        [--area_ratio_thresh #] [--check] [--verbose] [--print_memory] [--gpu #]\n";

int main(int argc, char **argv)
{
  const char *amosaic = NULL, *lmosaic = NULL, *omosaic = NULL, *otopog = NULL;
  char mosaic_name[STRING] = "mosaic", mosaic_file[STRING + 4];
  double sea_level = 0., area_ratio_thresh = 1.0e-6;
  int interp_order = 2, device = 0, c, idx;
  static struct option opts[] = {
      {"atmos_mosaic", required_argument, NULL, 'a'}, {"land_mosaic", required_argument, NULL, 'l'},
      {"ocean_mosaic", required_argument, NULL, 'o'}, {"wave_mosaic", required_argument, NULL, 'w'},
      {"ocean_topog", required_argument, NULL, 't'}, {"sea_level", required_argument, NULL, 's'},
      {"interp_order", required_argument, NULL, 'i'}, {"mosaic_name", required_argument, NULL, 'm'},
      {"area_ratio_thresh", required_argument, NULL, 'r'}, {"check", no_argument, NULL, 'n'},
      {"verbose", no_argument, NULL, 'v'}, {"print_memory", no_argument, NULL, 'p'}, {"rotate_poly", no_argument, NULL, 'u'},
      {"gpu", required_argument, NULL, 'g'}, {"help", no_argument, NULL, 'h'}, {0, 0, 0, 0}};
  while ((c = getopt_long(argc, argv, "", opts, &idx)) != -1) {
    switch (c) {
      case 'a': amosaic = optarg; break;
      case 'l': lmosaic = optarg; break;
      case 'o': omosaic = optarg; break;
      case 'w': die("make_coupler_mosaic_b200: --wave_mosaic is not built");
      case 't': otopog = optarg; break;
      case 's': sea_level = atof(optarg); break;
      case 'i': interp_order = atoi(optarg); break;
      case 'm': strncpy(mosaic_name, optarg, STRING - 1); break;
      case 'r': area_ratio_thresh = atof(optarg); break;
      case 'n': case 'v': case 'p': break;
      case 'u': die("make_coupler_mosaic_b200: --rotate_poly is not built");
      case 'g': device = atoi(optarg); break;
      default: fputs(usage, stderr); return c == 'h' ? 0 : 1;
    }
  }
  /* :483-497 */
  if (!amosaic) die("make_coupler_mosaic: atmos_mosaic is not specified");
  if (!omosaic) die("make_coupler_mosaic: ocean_mosaic is not specified");
  if (!otopog) die("make_coupler_mosaic: ocean_topog is not specified");
  if (interp_order != 1 && interp_order != 2) die("make_coupler_mosaic: interp_order should be 1 or 2");
  if (!lmosaic) lmosaic = amosaic;
  g_history[0] = 0;
  for (int k = 0; k < argc && strlen(g_history) + strlen(argv[k]) + 2 < sizeof g_history; ++k) { if (k) strcat(g_history, " "); strcat(g_history, argv[k]); }
  snprintf(mosaic_file, sizeof mosaic_file, "%s.nc", mosaic_name);
  char adir[STRING], afile[STRING], ldir[STRING], lfile[STRING], odir[STRING], ofile[STRING], tdir[STRING], tfile[STRING];
  dir_and_name(amosaic, adir, afile); dir_and_name(lmosaic, ldir, lfile); dir_and_name(omosaic, odir, ofile); dir_and_name(otopog, tdir, tfile);
  if (!strcmp(mosaic_file, afile) || !strcmp(mosaic_file, lfile) || !strcmp(mosaic_file, ofile))
    die("make_coupler_mosaic: mosaic_file can not have the same name as amosaic, lmosaic or omosaic");
  if (!strcmp(afile, "mosaic.nc") || !strcmp(lfile, "mosaic.nc") || !strcmp(ofile, "mosaic.nc"))
    die("make_coupler_mosaic: the file name of amosaic, lmosaic or omosaic can not be mosaic.nc");

  /* ---- grids */
  Mosaic A, L, O;
  const int lnd_same_as_atm = !strcmp(lmosaic, amosaic), ocn_same_as_atm = !strcmp(omosaic, amosaic);   /* :662, :769 */
  load_mosaic(amosaic, &A, "atmos");
  {                                                                     /* nested atmosphere mosaics: refused (:623-660) */
    nc3_file *f = open_or_die(amosaic);
    const int d = nc3_dim_id(f, "ncontact");
    if (d >= 0 && A.ntiles != 6 && A.ntiles != 1 && nc3_dim_len(f, d) > 0)
      die("make_coupler_mosaic_b200: an atmosphere mosaic of %d tiles with contacts looks nested; nested grids are not built", A.ntiles);
    nc3_close(f);
  }
  if (lnd_same_as_atm) L = A; else load_mosaic(lmosaic, &L, "land");
  load_mosaic(omosaic, &O, "ocean");
  int ocn_south_ext = 0;
  if (O.ntiles == 1) {                                                  /* :840-876 */
    const int nx = O.nx[0], ny = O.ny[0];
    const double min_atm_lat = -90. * D2R;
    if (O.y[0][0] > min_atm_lat + TINY_VALUE) {
      ocn_south_ext = 1;
      printf("make_coupler_mosaic: one row is added to the south end to cover the globe\n");
      double *x = xmalloc((size_t)(nx + 1) * (ny + 2) * sizeof(double)), *y = xmalloc((size_t)(nx + 1) * (ny + 2) * sizeof(double));
      memcpy(x + nx + 1, O.x[0], (size_t)(nx + 1) * (ny + 1) * sizeof(double));
      memcpy(y + nx + 1, O.y[0], (size_t)(nx + 1) * (ny + 1) * sizeof(double));
      for (int i = 0; i <= nx; ++i) { x[i] = x[nx + 1 + i]; y[i] = min_atm_lat; }
      free(O.x[0]); free(O.y[0]);
      O.x[0] = x; O.y[0] = y; O.ny[0] = ny + 1;
    }
  }
  /* ---- ocean topography -> omask (:901-957) */
  long long nocn = 0;
  for (int n = 0; n < O.ntiles; ++n) nocn += (long long)O.nx[n] * O.ny[n];
  double *omask = xmalloc((size_t)nocn * sizeof(double));
  {
    nc3_file *t = open_or_die(otopog);
    const int ntiles = nc3_dim_id(t, "ntiles") >= 0 ? (int)nc3_dim_len(t, nc3_dim_id(t, "ntiles")) : 1;
    if (ntiles != O.ntiles) die("make_coupler_mosaic: dimlen ntiles in mosaic file is not the same as dimlen in topog file");
    long long at = 0;
    for (int n = 0; n < O.ntiles; ++n) {
      char nxn[64] = "nx", nyn[64] = "ny", dn[64] = "depth", mn[64] = "area_frac";
      if (ntiles > 1) { sprintf(nxn, "nx_tile%d", n + 1); sprintf(nyn, "ny_tile%d", n + 1); sprintf(dn, "depth_tile%d", n + 1); sprintf(mn, "area_frac_tile%d", n + 1); }
      const int nx = (int)need_dim(t, otopog, nxn), ny = (int)need_dim(t, otopog, nyn);
      if (nx != O.nx[n] || ny + ocn_south_ext != O.ny[n]) die("make_coupler_mosaic: grid size mismatch between mosaic file and topog file");
      double *m = omask + at + (long long)ocn_south_ext * nx;
      if (nc3_var_id(t, mn) >= 0) {
        if (nc3_get_var_double(t, nc3_var_id(t, mn), m)) die("%s: %s", otopog, nc3_strerror(t));
      } else {
        double *depth = xmalloc((size_t)nx * ny * sizeof(double));
        if (nc3_get_var_double(t, need_var(t, otopog, dn), depth)) die("%s: %s", otopog, nc3_strerror(t));
        for (long long k = 0; k < (long long)nx * ny; ++k) if (depth[k] > sea_level) m[k] = 1;
        free(depth);
      }
      at += (long long)O.nx[n] * O.ny[n];
    }
    nc3_close(t);
  }
  /* :1078-1091 */
  int same_mosaic;
  if (strcmp(O.name, A.name)) {
    if (!strcmp(O.name, L.name)) die("make_coupler_mosaic: omosaic is the same as lmosaic, but different from amosaic.");
    same_mosaic = 0;
  } else {
    if (strcmp(O.name, L.name)) die("make_coupler_mosaic: omosaic is the same as amosaic, but different from lmosaic.");
    same_mosaic = 1;
  }
  for (int n = 0; n < L.ntiles && n < A.ntiles; ++n)
    if (A.nx[n] * A.ny[n] != L.nx[n] * L.ny[n])
      printf("Warning: Number of ATM and LND cells for tile %d are not equal %d, %d.\n", n + 1, A.nx[n] * A.ny[n], L.nx[n] * L.ny[n]);

  /* ---- the exchange grids: one call instead of :1250-2017 and :2556-2808 */
  double *alon, *alat, *llon = NULL, *llat = NULL, *olon, *olat;
  xgb_mosaic_grid ga, gl, go;
  concat(&A, &alon, &alat, &ga);
  if (!lnd_same_as_atm) concat(&L, &llon, &llat, &gl);
  concat(&O, &olon, &olat, &go);
  xgb_coupler_result r;
  if (xgb_make_coupler_xgrid(device, interp_order, area_ratio_thresh, -1, lnd_same_as_atm, ocn_same_as_atm, &ga,
                             lnd_same_as_atm ? NULL : &gl, &go, omask, &r))
    die("%s", xgb_last_error());

  /* ---- ocean_mask (:2020-2070) */
  int nbad = 0;
  {
    long long at = 0;
    for (int n = 0; n < O.ntiles; ++n) {
      const int nx = O.nx[n], ny = O.ny[n] - ocn_south_ext;
      char file[STRING];
      double *mask = xmalloc((size_t)nx * ny * sizeof(double));
      const double *ax = r.ocn_xarea + at + (long long)ocn_south_ext * nx, *ao = r.area_ocn + at + (long long)ocn_south_ext * nx;
      const double *om = omask + at + (long long)ocn_south_ext * nx;
      for (long long k = 0; k < (long long)nx * ny; ++k) {
        mask[k] = ax[k] / ao[k];
        if (fabs(om[k] - mask[k]) > TOLORENCE) {
          ++nbad;
          printf("at ocean point (%d,%d), omask = %f, ocn_frac = %f, diff = %f\n", (int)(k % nx), (int)(k / nx), om[k], mask[k], om[k] - mask[k]);
        }
      }
      if (O.ntiles > 1) sprintf(file, "ocean_mask_tile%d.nc", n + 1); else strcpy(file, "ocean_mask.nc");
      nc3_file *f = create_or_die(file);
      int dims[2];
      dims[1] = nc3_def_dim(f, "nx", nx); dims[0] = nc3_def_dim(f, "ny", ny);
      const int vm = def_2d(f, "mask", dims, "ocean fraction at T-cell centers"), vo = def_2d(f, "areaO", dims, "ocean grid area"),
                vx = def_2d(f, "areaX", dims, "ocean exchange grid area");
      if (nc3_enddef(f) || nc3_put_var_double(f, vm, mask) || nc3_put_var_double(f, vo, ao) || nc3_put_var_double(f, vx, ax) || nc3_close(f))
        die("make_coupler_mosaic: cannot write %s", file);
      free(mask);
      at += (long long)O.nx[n] * O.ny[n];
    }
    if (nbad > 0) printf("make_coupler_mosaic: number of points with omask != ofrac is %d\n", nbad);
  }
  /* ---- land_mask (:2072-2120; its `area_atm` is area_atm[nl][i], the atmosphere tile with the LAND tile's number) */
  {
    long long at = 0, aat = 0;
    for (int n = 0; n < L.ntiles; ++n) {
      const int nx = L.nx[n], ny = L.ny[n];
      const long long nc = (long long)nx * ny;
      char file[STRING];
      double *mask = xmalloc((size_t)nc * sizeof(double)), *aa = xmalloc((size_t)nc * sizeof(double));
      const long long natm = (n < A.ntiles) ? (long long)A.nx[n] * A.ny[n] : 0;
      for (long long k = 0; k < nc; ++k) {
        mask[k] = r.lnd_xarea[at + k] / r.area_lnd[at + k];
        aa[k] = (k < natm) ? r.area_atm[aat + k] : 0.0;                   /* the reference reads past the tile here; not reproduced */
      }
      if (L.ntiles > 1) sprintf(file, "land_mask_tile%d.nc", n + 1); else strcpy(file, "land_mask.nc");
      nc3_file *f = create_or_die(file);
      int dims[2];
      dims[1] = nc3_def_dim(f, "nx", nx); dims[0] = nc3_def_dim(f, "ny", ny);
      const int vm = def_2d(f, "mask", dims, "land fraction at T-cell centers"), va = def_2d(f, "area_atm", dims, "area atm "),
                vl = def_2d(f, "area_lnd", dims, "area land "), vx = def_2d(f, "l_area", dims, "land x area");
      if (nc3_enddef(f) || nc3_put_var_double(f, vm, mask) || nc3_put_var_double(f, vl, r.area_lnd + at) || nc3_put_var_double(f, va, aa) ||
          nc3_put_var_double(f, vx, r.lnd_xarea + at) || nc3_close(f))
        die("make_coupler_mosaic: cannot write %s", file);
      free(mask); free(aa);
      at += nc; aat += natm;
    }
  }
  /* ---- exchange-grid files, in the tool's order: per atmosphere tile its land files, then its ocean files; then land x ocean */
  enum { MAXF = 4096 };
  char (*axl)[STRING] = xmalloc(MAXF * sizeof *axl), (*axo)[STRING] = xmalloc(MAXF * sizeof *axo), (*lxo)[STRING] = xmalloc(MAXF * sizeof *lxo);
  int naxl = 0, naxo = 0, nlxo = 0;
  char contact[4 * STRING];
  for (int na = 0; na < A.ntiles; ++na) {
    for (int nl = 0; nl < L.ntiles; ++nl) {
      int any = 0;
      for (long long e = 0; e < r.atmxlnd.n && !any; ++e) any = (r.atmxlnd.t1[e] == na && r.atmxlnd.t2[e] == nl);
      if (!any || naxl >= MAXF) continue;
      snprintf(axl[naxl], STRING, same_mosaic ? "atm_%s_%sXlnd_%s_%s.nc" : "%s_%sX%s_%s.nc", A.name, A.tile[na], L.name, L.tile[nl]);
      snprintf(contact, sizeof contact, "%s:%s::%s:%s", A.name, A.tile[na], L.name, L.tile[nl]);
      write_xgrid(axl[naxl], contact, interp_order, &r.atmxlnd, na, nl, 0, "parent_cell_indices_in_mosaic1", "parent_cell_indices_in_mosaic2");
      ++naxl;
    }
    for (int no = 0; no < O.ntiles; ++no) {
      int any = 0;
      for (long long e = 0; e < r.atmxocn.n && !any; ++e) any = (r.atmxocn.t1[e] == na && r.atmxocn.t2[e] == no);
      if (!any || naxo >= MAXF) continue;
      snprintf(axo[naxo], STRING, same_mosaic ? "atm_%s_%sXocn_%s_%s.nc" : "%s_%sX%s_%s.nc", A.name, A.tile[na], O.name, O.tile[no]);
      snprintf(contact, sizeof contact, "%s:%s::%s:%s", A.name, A.tile[na], O.name, O.tile[no]);
      write_xgrid(axo[naxo], contact, interp_order, &r.atmxocn, na, no, ocn_south_ext, "parent_cell_indices_in_mosaic1", "parent_cell_indices_in_mosaic2");
      ++naxo;
    }
  }
  for (int nl = 0; nl < L.ntiles; ++nl)
    for (int no = 0; no < O.ntiles; ++no) {
      int any = 0;
      for (long long e = 0; e < r.lndxocn.n && !any; ++e) any = (r.lndxocn.t1[e] == nl && r.lndxocn.t2[e] == no);
      if (!any || nlxo >= MAXF) continue;
      snprintf(lxo[nlxo], STRING, same_mosaic ? "lnd_%s_%sXocn_%s_%s.nc" : "%s_%sX%s_%s.nc", L.name, L.tile[nl], O.name, O.tile[no]);
      snprintf(contact, sizeof contact, "%s:%s::%s:%s", L.name, L.tile[nl], O.name, O.tile[no]);
      write_xgrid(lxo[nlxo], contact, interp_order, &r.lndxocn, nl, no, ocn_south_ext, "parent1_cell_indices", "parent2_cell_indices");
      ++nlxo;
    }
  if (lnd_same_as_atm) {                      /* :2968-2974: land on the atmosphere mosaic uses the atm x ocn grids for runoff */
    nlxo = naxo;
    for (int n = 0; n < naxo; ++n) strcpy(lxo[n], axo[n]);
  }
  /* ---- the coupler mosaic file (:3664-3805) */
  {
    nc3_file *f = create_or_die(mosaic_file);
    int dims[2], d_axo = -1, d_axl = -1, d_lxo = -1;
    const int d_string = nc3_def_dim(f, "string", STRING);
    if (naxo > 0) d_axo = nc3_def_dim(f, "nfile_aXo", naxo);
    if (naxl > 0) d_axl = nc3_def_dim(f, "nfile_aXl", naxl);
    if (nlxo > 0) d_lxo = nc3_def_dim(f, "nfile_lXo", nlxo);
    const int v0 = def_text_var(f, "atm_mosaic_dir", 1, &d_string, 1, "standard_name", "directory_storing_atmosphere_mosaic");
    const int v1 = def_text_var(f, "atm_mosaic_file", 1, &d_string, 1, "standard_name", "atmosphere_mosaic_file_name");
    const int v2 = def_text_var(f, "atm_mosaic", 1, &d_string, 1, "standard_name", "atmosphere_mosaic_name");
    const int v3 = def_text_var(f, "lnd_mosaic_dir", 1, &d_string, 1, "standard_name", "directory_storing_land_mosaic");
    const int v4 = def_text_var(f, "lnd_mosaic_file", 1, &d_string, 1, "standard_name", "land_mosaic_file_name");
    const int v5 = def_text_var(f, "lnd_mosaic", 1, &d_string, 1, "standard_name", "land_mosaic_name");
    const int v6 = def_text_var(f, "ocn_mosaic_dir", 1, &d_string, 1, "standard_name", "directory_storing_ocean_mosaic");
    const int v7 = def_text_var(f, "ocn_mosaic_file", 1, &d_string, 1, "standard_name", "ocean_mosaic_file_name");
    const int v8 = def_text_var(f, "ocn_mosaic", 1, &d_string, 1, "standard_name", "ocean_mosaic_name");
    const int v9 = def_text_var(f, "ocn_topog_dir", 1, &d_string, 1, "standard_name", "directory_storing_ocean_topog");
    const int v10 = def_text_var(f, "ocn_topog_file", 1, &d_string, 1, "standard_name", "ocean_topog_file_name");
    int vaxo = -1, vaxl = -1, vlxo = -1;
    dims[1] = d_string;
    if (naxo > 0) { dims[0] = d_axo; vaxo = def_text_var(f, "aXo_file", 2, dims, 1, "standard_name", "atmXocn_exchange_grid_file"); }
    if (naxl > 0) { dims[0] = d_axl; vaxl = def_text_var(f, "aXl_file", 2, dims, 1, "standard_name", "atmXlnd_exchange_grid_file"); }
    if (nlxo > 0) { dims[0] = d_lxo; vlxo = def_text_var(f, "lXo_file", 2, dims, 1, "standard_name", "lndXocn_exchange_grid_file"); }
    if (nc3_enddef(f)) die("%s: %s", mosaic_file, nc3_strerror(f));
    put_text(f, v0, 0, 1, adir); put_text(f, v1, 0, 1, afile); put_text(f, v2, 0, 1, A.name);
    put_text(f, v3, 0, 1, ldir); put_text(f, v4, 0, 1, lfile); put_text(f, v5, 0, 1, L.name);
    put_text(f, v6, 0, 1, odir); put_text(f, v7, 0, 1, ofile); put_text(f, v8, 0, 1, O.name);
    put_text(f, v9, 0, 1, tdir); put_text(f, v10, 0, 1, tfile);
    for (int n = 0; n < naxo; ++n) put_text(f, vaxo, n, 2, axo[n]);
    for (int n = 0; n < naxl; ++n) put_text(f, vaxl, n, 2, axl[n]);
    for (int n = 0; n < nlxo; ++n) put_text(f, vlxo, n, 2, lxo[n]);
    if (nc3_close(f)) die("make_coupler_mosaic: cannot close %s", mosaic_file);
  }
  xgb_coupler_result_free(&r);
  return 0;
}
