/* TEST INFRASTRUCTURE — CPU oracle, never shipped or measured as the product.
 *
 * Great-circle exchange grid of fregrid (--great_circle_algorithm), restated from the reference:
 *   orc_create_xgrid_great_circle   create_xgrid.c:1366-1466
 *   orc_clip_2dx2d_great_circle     create_xgrid.c:1479-1908   (polygon walk over intersection lists)
 *   gc_line_intersect               create_xgrid.c:1919-2081   (line_intersect_2D_3D)
 *   gc_plane_line_param             mosaic_util.c:967-1043     (intersect_tri_with_line / invert_matrix_3x3 / mult,
 *                                                               x87 long double exactly as the reference declares it)
 *   gc_inside_polygon               mosaic_util.c:1487-1530    (insidePolygon)
 *   orc_great_circle_area           mosaic_util.c:763-790
 *   orc_get_grid_great_circle_area  create_xgrid.c:98-137
 * The reference keeps its polygons in singly linked lists carved out of a global 100-node pool
 * (mosaic_util.c:1047-1078); here every list is a small array and "insert after" is an array insertion.  The
 * operations on the lists (duplicate suppression in addEnd :1095, addIntersect :1133, insertIntersect :1291,
 * setInbound :1437, getFirstInbound :1389) are restated one for one, because which vertex ends up where decides
 * the cell lists.  Pinned against the compiled reference by tests/test_gc_cpu.py and tests/golden/gc_*.npz.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "xgrid_oracle.h"

#define GC_RADIUS   6371000.0
#define GC_RANGE    0.05        /* RANGE_CHECK_CRITERIA, mosaic_util.h:26 */
#define GC_EPS8     1.e-8
#define GC_EPS10    1.e-10
#define GC_EPS15    1.e-15
#define GC_EPS30    1.e-30
#define GC_CAP      48          /* nodes per list; the reference's pool holds 100 nodes for all lists together */

typedef struct {
  double x, y, z, u, u_clip;
  int intersect;   /* 0 vertex, 1 intersection, 2 vertex that is also an intersection */
  int inbound;     /* 0 undecided, 1 going out, 2 going in */
  int inside;      /* 1 inside the other polygon, 0 outside, -1 undecided */
  int subj_index, clip_index;
} GcNode;

typedef struct { GcNode n[GC_CAP]; int len; int overflow; } GcList;

double orc_spherical_angle(const double v1[3], const double v2[3], const double v3[3]);   /* xgrid_oracle_grad.c */

static int gc_same_point(double x1, double y1, double z1, double x2, double y2, double z2)   /* mosaic_util.c:1193 */
{
  return !(fabs(x1-x2) > GC_EPS10 || fabs(y1-y2) > GC_EPS10 || fabs(z1-z2) > GC_EPS10);
}

static int gc_same_node(const GcNode *a, const GcNode *b) { return a->x == b->x && a->y == b->y && a->z == b->z; }

static GcNode *gc_grow(GcList *l, int at)
{
  int k;
  if (l->len >= GC_CAP) { l->overflow = 1; return &l->n[GC_CAP-1]; }
  for (k = l->len; k > at; k--) l->n[k] = l->n[k-1];
  l->len++;
  return &l->n[at];
}

/* addEnd: append unless a point within EPSLN10 is already in the list */
static void gc_append_unique(GcList *l, double x, double y, double z, int intersect, double u, int inbound, int inside)
{
  GcNode *q;
  int k;
  for (k = 0; k < l->len; k++) if (gc_same_point(l->n[k].x, l->n[k].y, l->n[k].z, x, y, z)) return;
  q = gc_grow(l, l->len);
  memset(q, 0, sizeof(*q));
  q->x = x; q->y = y; q->z = z; q->u = u; q->intersect = intersect; q->inbound = inbound; q->inside = inside;
}

/* addIntersect: 1 if appended, 0 if an intersection on the same subject edge / clip edge parameter is already there */
static int gc_add_intersect(GcList *l, const double p[3], double u1, double u2, int inbound, int is1, int ie1, int is2, int ie2)
{
  double u1c = u1, u2c = u2;
  int i1c = is1, i2c = is2, k;
  GcNode *q;
  if (u1c == 1) { u1c = 0; i1c = ie1; }
  if (u2c == 1) { u2c = 0; i2c = ie2; }
  for (k = 0; k < l->len; k++) {
    if (l->n[k].u == u1c && l->n[k].subj_index == i1c) return 0;
    if (l->n[k].u_clip == u2c && l->n[k].clip_index == i2c) return 0;
  }
  q = gc_grow(l, l->len);
  q->x = p[0]; q->y = p[1]; q->z = p[2]; q->intersect = 1; q->inbound = inbound; q->inside = 0;
  q->u = u1c; q->subj_index = i1c; q->u_clip = u2c; q->clip_index = i2c;
  return 1;
}

/* insertIntersect: put intersection p on the edge that starts at vertex v (exact coordinates), ordered by u.
 * returns 0, or 1 when v is not in the list (the reference aborts) */
static int gc_insert_intersect(GcList *l, const double p[3], double u1, double u2, int inbound, const double v[3])
{
  int a = -1, b, k;
  double ucur = u1;
  GcNode *q;
  for (k = 0; k < l->len; k++) if (l->n[k].x == v[0] && l->n[k].y == v[1] && l->n[k].z == v[2]) { a = k; break; }
  if (a < 0) return 1;
  if (u1 == 1) { ucur = 0; a = (a + 1 < l->len) ? a + 1 : 0; }
  if (ucur == 0) {                                   /* the vertex itself is the intersection (:1322-1330) */
    l->n[a].intersect = 2; l->n[a].inside = 1; l->n[a].u = ucur;
    l->n[a].x = p[0]; l->n[a].y = p[1]; l->n[a].z = p[2];
    return 0;
  }
  if (u2 != 0 && u2 != 1) {                          /* :1333-1348 */
    if (inbound == 1) {
      b = (a + 1 < l->len) ? a + 1 : 0;
      while (l->n[b].intersect) b = (b + 1 < l->len) ? b + 1 : 0;
      l->n[b].inside = 0;
    } else if (inbound == 2) l->n[a].inside = 0;
  }
  b = a + 1;                                         /* :1350-1361, no wrap-around */
  while (b < l->len) {
    if (l->n[b].intersect == 1) { if (l->n[b].u > ucur) break; }
    else break;
    a = b; b++;
  }
  q = gc_grow(l, a + 1);
  memset(q, 0, sizeof(*q));
  q->x = p[0]; q->y = p[1]; q->z = p[2]; q->u = ucur; q->intersect = 1; q->inbound = inbound; q->inside = 1;
  return 0;
}

double orc_great_circle_area(int n, const double *x, const double *y, const double *z)
{
  double sum = 0.0, p0[3], p1[3], p2[3];
  int i;
  for (i = 0; i < n; i++) {
    p0[0] = x[i]; p0[1] = y[i]; p0[2] = z[i];
    p1[0] = x[(i+1)%n]; p1[1] = y[(i+1)%n]; p1[2] = z[(i+1)%n];
    p2[0] = x[(i+2)%n]; p2[1] = y[(i+2)%n]; p2[2] = z[(i+2)%n];
    sum += orc_spherical_angle(p1, p2, p0);
  }
  return (sum - (n-2.)*M_PI)*GC_RADIUS*GC_RADIUS;
}

static double gc_list_area(const GcList *l)           /* gridArea, mosaic_util.c:1364 */
{
  double x[GC_CAP], y[GC_CAP], z[GC_CAP];
  int k;
  for (k = 0; k < l->len; k++) { x[k] = l->n[k].x; y[k] = l->n[k].y; z[k] = l->n[k].z; }
  return orc_great_circle_area(l->len, x, y, z);
}

static int gc_inside_polygon(const GcNode *q, const GcList *l)
{
  double sum = 0, p0[3], p1[3], p2[3];
  int k;
  p0[0] = q->x; p0[1] = q->y; p0[2] = q->z;
  for (k = 0; k < l->len; k++) {
    const GcNode *a = &l->n[k], *b = &l->n[(k+1 < l->len) ? k+1 : 0];
    p1[0] = a->x; p1[1] = a->y; p1[2] = a->z;
    p2[0] = b->x; p2[1] = b->y; p2[2] = b->z;
    if (gc_same_point(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2])) return 1;
    sum += orc_spherical_angle(p0, p2, p1);
  }
  return fabs(sum - 2*M_PI) < GC_EPS8;
}

/* parameter t of the point where the line l1 + t (l2 - l1) meets the plane through a, b and the origin.
 * intersect_tri_with_line builds M = [l1-l2 | b-a | 0-a], inverts it by cofactors and multiplies by l1-a, all in
 * long double; only row 0 of the inverse reaches the caller.  returns 0 when |det| < EPSLN15. */
static int gc_plane_line_param(const double *a, const double *b, const double *l1, const double *l2, double *t)
{
  long double m[9], v[3], det, deti, r0, r1, r2;
  m[0] = l1[0]-l2[0]; m[1] = b[0]-a[0]; m[2] = 0.0-a[0];
  m[3] = l1[1]-l2[1]; m[4] = b[1]-a[1]; m[5] = 0.0-a[1];
  m[6] = l1[2]-l2[2]; m[7] = b[2]-a[2]; m[8] = 0.0-a[2];
  det = m[0]*(m[4]*m[8] - m[5]*m[7]) - m[1]*(m[3]*m[8] - m[5]*m[6]) + m[2]*(m[3]*m[7] - m[4]*m[6]);
  if (fabsl(det) < GC_EPS15) return 0;
  deti = 1.0/det;
  r0 = (m[4]*m[8] - m[5]*m[7])*deti;
  r1 = (m[2]*m[7] - m[1]*m[8])*deti;
  r2 = (m[1]*m[5] - m[2]*m[4])*deti;
  v[0] = l1[0]-a[0]; v[1] = l1[1]-a[1]; v[2] = l1[2]-a[2];
  *t = r0*v[0] + r1*v[1] + r2*v[2];
  return 1;
}

static void gc_cross(const double *a, const double *b, double *e)
{
  e[0] = a[1]*b[2] - a[2]*b[1];
  e[1] = a[2]*b[0] - a[0]*b[2];
  e[2] = a[0]*b[1] - a[1]*b[0];
}

static int gc_line_intersect(const double *a1, const double *a2, const double *q1, const double *q2, const double *q3,
                             double *p, double *ua, double *uq, int *inbound)
{
  double c1[3], c2[3], c3[3], d[3], v1[3], v2[3], u, norm;
  int k;
  *inbound = 0;
  if (gc_same_point(a1[0], a1[1], a1[2], q1[0], q1[1], q1[2])) { *ua = 0; *uq = 0; p[0] = a1[0]; p[1] = a1[1]; p[2] = a1[2]; return 1; }
  if (gc_same_point(a1[0], a1[1], a1[2], q2[0], q2[1], q2[2])) { *ua = 0; *uq = 1; p[0] = a1[0]; p[1] = a1[1]; p[2] = a1[2]; return 1; }
  if (gc_same_point(a2[0], a2[1], a2[2], q1[0], q1[1], q1[2])) { *ua = 1; *uq = 0; p[0] = a2[0]; p[1] = a2[1]; p[2] = a2[2]; return 1; }
  if (gc_same_point(a2[0], a2[1], a2[2], q2[0], q2[1], q2[2])) { *ua = 1; *uq = 1; p[0] = a2[0]; p[1] = a2[1]; p[2] = a2[2]; return 1; }

  if (!gc_plane_line_param(q1, q2, a1, a2, ua)) return 0;
  if (fabs(*ua) < GC_EPS8) *ua = 0;
  if (fabs(*ua - 1) < GC_EPS8) *ua = 1;
  if (*ua < 0 || *ua > 1) return 0;
  if (!gc_plane_line_param(a1, a2, q1, q2, uq)) return 0;
  if (fabs(*uq) < GC_EPS8) *uq = 0;
  if (fabs(*uq - 1) < GC_EPS8) *uq = 1;
  if (*uq < 0 || *uq > 1) return 0;
  u = *ua;
  gc_cross(a1, a2, c1); gc_cross(q1, q2, c2); gc_cross(c1, c2, c3);
  if (fabs(sqrt(c3[0]*c3[0] + c3[1]*c3[1] + c3[2]*c3[2])) < GC_EPS30) return 0;     /* coincident planes */
  p[0] = a1[0] + u*(a2[0]-a1[0]);
  p[1] = a1[1] + u*(a2[1]-a1[1]);
  p[2] = a1[2] + u*(a2[2]-a1[2]);
  norm = sqrt(p[0]*p[0] + p[1]*p[1] + p[2]*p[2]);
  for (k = 0; k < 3; k++) p[k] /= norm;
  if (*uq != 0 && *uq != 1) {
    for (k = 0; k < 3; k++) { d[k] = a2[k]-a1[k]; v1[k] = q2[k]-q1[k]; v2[k] = q3[k]-q2[k]; }
    gc_cross(v1, v2, c1); gc_cross(v1, d, c2);
    *inbound = (c1[0]*c2[0] + c1[1]*c2[1] + c1[2]*c2[2] > 0) ? 2 : 1;
  }
  return 1;
}

static int gc_find_exact(const GcList *l, const GcNode *q)
{
  int k;
  for (k = 0; k < l->len; k++) if (gc_same_node(&l->n[k], q)) return k;
  return -1;
}

static int gc_first_inbound(const GcList *l, GcNode *out)
{
  int k;
  for (k = 0; k < l->len; k++) if (l->n[k].inbound == 2) { *out = l->n[k]; return 1; }
  return 0;
}

/* setInbound, mosaic_util.c:1437; returns 1 if an intersection is missing from the subject list */
static int gc_set_inbound(GcList *inter, const GcList *subj)
{
  int k, a;
  for (k = 0; k < inter->len; k++) {
    if (inter->n[k].inbound) continue;
    a = gc_find_exact(subj, &inter->n[k]);
    if (a < 0) return 1;
    {
      const GcNode *prev = &subj->n[a > 0 ? a-1 : subj->len-1], *next = &subj->n[a+1 < subj->len ? a+1 : 0];
      inter->n[k].inbound = (prev->inside == 0 && next->inside == 1) ? 2 : 1;
    }
  }
  return 0;
}

/* error codes (negative return): the conditions on which the reference calls error_handler() */
#define GC_ERR_NOT_CONVEX   (-1)
#define GC_ERR_WALK         (-2)
#define GC_ERR_POOL         (-3)

int orc_clip_2dx2d_great_circle(const double x1[], const double y1[], const double z1[], int n1,
                                const double x2[], const double y2[], const double z2[], int n2,
                                double xo[], double yo[], double zo[])
{
  static GcList g1, g2, inter, poly;
  double pt1[GC_CAP][3], pt2[GC_CAP][3], lo, hi, p[3], u1, u2;
  int i1, i2, k, npts1, npts2, nint, n_out = 0, has_inbound = 0, inbound;
  GcNode first, cur;

#define GC_MIN(v, n, r) { int q_; r = v[0]; for (q_ = 1; q_ < n; q_++) if (v[q_] < r) r = v[q_]; }
#define GC_MAX(v, n, r) { int q_; r = v[0]; for (q_ = 1; q_ < n; q_++) if (v[q_] > r) r = v[q_]; }
  GC_MIN(x1, n1, lo); GC_MAX(x2, n2, hi); if (lo >= hi + GC_RANGE) return 0;      /* :1508-1528 */
  GC_MAX(x1, n1, hi); GC_MIN(x2, n2, lo); if (lo >= hi + GC_RANGE) return 0;
  GC_MIN(y1, n1, lo); GC_MAX(y2, n2, hi); if (lo >= hi + GC_RANGE) return 0;
  GC_MAX(y1, n1, hi); GC_MIN(y2, n2, lo); if (lo >= hi + GC_RANGE) return 0;
  GC_MIN(z1, n1, lo); GC_MAX(z2, n2, hi); if (lo >= hi + GC_RANGE) return 0;
  GC_MAX(z1, n1, hi); GC_MIN(z2, n2, lo); if (lo >= hi + GC_RANGE) return 0;

  g1.len = g2.len = inter.len = poly.len = 0;
  g1.overflow = g2.overflow = inter.overflow = poly.overflow = 0;
  for (k = 0; k < n1; k++) gc_append_unique(&g1, x1[k], y1[k], z1[k], 0, 0, 0, -1);
  for (k = 0; k < n2; k++) gc_append_unique(&g2, x2[k], y2[k], z2[k], 0, 0, 0, -1);
  npts1 = g1.len; npts2 = g2.len;
  for (k = 0; k < g1.len; k++) g1.n[k].inside = gc_inside_polygon(&g1.n[k], &g2);   /* :1549-1568 */
  for (k = 0; k < g2.len; k++) g2.n[k].inside = gc_inside_polygon(&g2.n[k], &g1);
  if (gc_list_area(&g1) <= 0 || gc_list_area(&g2) <= 0) return GC_ERR_NOT_CONVEX;   /* :1575-1578 */
  for (k = 0; k < npts1; k++) { pt1[k][0] = g1.n[k].x; pt1[k][1] = g1.n[k].y; pt1[k][2] = g1.n[k].z; }
  for (k = 0; k < npts2; k++) { pt2[k][0] = g2.n[k].x; pt2[k][1] = g2.n[k].y; pt2[k][2] = g2.n[k].z; }

  for (i1 = 0; i1 < npts1; i1++) {                                                   /* :1606-1670 */
    const int i1p = (i1+1)%npts1;
    double *a0 = pt1[i1], *a1 = pt1[i1p];
    for (i2 = 0; i2 < npts2; i2++) {
      const int i2p = (i2+1)%npts2, i2p2 = (i2+2)%npts2;
      double *b0 = pt2[i2], *b1 = pt2[i2p], *b2 = pt2[i2p2];
      if (!gc_line_intersect(a0, a1, b0, b1, b2, p, &u1, &u2, &inbound)) continue;
      if (!gc_add_intersect(&inter, p, u1, u2, inbound, i1, i1p, i2, i2p)) continue;
      if (u1 == 1) { if (gc_insert_intersect(&g1, p, 0.0, u2, inbound, a1)) return GC_ERR_WALK; }
      else         { if (gc_insert_intersect(&g1, p, u1, u2, inbound, a0)) return GC_ERR_WALK; }
      if (u1 == 1)      { a1[0] = p[0]; a1[1] = p[1]; a1[2] = p[2]; }
      else if (u1 == 0) { a0[0] = p[0]; a0[1] = p[1]; a0[2] = p[2]; }
      if (u2 == 1) { if (gc_insert_intersect(&g2, p, 0.0, u1, 0, b1)) return GC_ERR_WALK; }
      else         { if (gc_insert_intersect(&g2, p, u2, u1, 0, b0)) return GC_ERR_WALK; }
      if (u2 == 1)      { b1[0] = p[0]; b1[1] = p[1]; b1[2] = p[2]; }
      else if (u2 == 0) { b0[0] = p[0]; b0[1] = p[1]; b0[2] = p[2]; }
    }
  }
  if (g1.overflow || g2.overflow || inter.overflow) return GC_ERR_POOL;

  nint = inter.len;                                                                  /* :1676-1693 */
  if (nint > 1) has_inbound = gc_first_inbound(&inter, &first);
  if (!has_inbound && nint > 1) {
    if (gc_set_inbound(&inter, &g1)) return GC_ERR_WALK;
    has_inbound = gc_first_inbound(&inter, &first);
  }

  if (has_inbound) {                                                                 /* :1697-1838 */
    const int maxiter1 = nint;
    GcList *curl = &g1;
    int iter1 = 0, found1 = 0, which = 0;
    if (gc_find_exact(&g1, &first) < 0) return GC_ERR_WALK;
    gc_append_unique(&poly, first.x, first.y, first.z, first.intersect, first.u, first.inbound, first.inside);
    nint--;
    cur = first;
    while (iter1 < maxiter1) {
      int a = gc_find_exact(curl, &cur), b, found2 = 0, iter2 = 0;
      const int maxiter2 = curl->len;
      if (a < 0) return GC_ERR_WALK;
      b = (a+1 < curl->len) ? a+1 : 0;
      while (iter2 < maxiter2) {
        int b_is_x = 0;
        const GcNode *nb = &curl->n[b];
        if (nb->intersect) {
          const GcNode *nc = &curl->n[(b+1 < curl->len) ? b+1 : 0];
          if (gc_same_node(nb, &first)) { found1 = 1; break; }
          found2 = 1; b_is_x = 1;
          if (nc->intersect || nc->inside == 1) found2 = 0;
        }
        if (found2) { cur = *nb; break; }
        gc_append_unique(&poly, nb->x, nb->y, nb->z, nb->intersect, nb->u, nb->inbound, nb->inside);
        if (b_is_x) nint--;
        b = (b+1 < curl->len) ? b+1 : 0;
        iter2++;
      }
      if (found1) break;
      if (!found2) return GC_ERR_WALK;                      /* " not found the next intersection " */
      if (gc_same_node(&cur, &first)) { found1 = 1; break; }
      gc_append_unique(&poly, cur.x, cur.y, cur.z, cur.intersect, cur.u, cur.inbound, cur.inside);
      nint--;
      which = !which; curl = which ? &g2 : &g1;
      iter1++;
    }
    if (!found1) return GC_ERR_WALK;                        /* "not return back to the first intersection" */
    if (nint > 0) return GC_ERR_WALK;                       /* "After clipping, nintersect should be 0" */
    if (poly.overflow) return GC_ERR_POOL;
    for (k = 0; k < poly.len; k++) { xo[k] = poly.n[k].x; yo[k] = poly.n[k].y; zo[k] = poly.n[k].z; }
    n_out = poly.len;
    if (n_out < 3) n_out = 0;
  }

  if (n_out == 0) {                                                                  /* grid1 inside grid2, :1841-1871 */
    int c = 0;
    for (k = 0; k < g1.len; k++) if (g1.n[k].intersect != 1 && g1.n[k].inside == 1) c++;
    if (c == npts1) {
      for (k = 0; k < g1.len; k++) { xo[k] = g1.n[k].x; yo[k] = g1.n[k].y; zo[k] = g1.n[k].z; }
      n_out = npts1;
    }
    if (n_out > 0) return n_out;
  }
  if (n_out == 0) {                                                                  /* grid2 inside grid1, :1874-1904 */
    int c = 0;
    for (k = 0; k < g2.len; k++) if (g2.n[k].intersect != 1 && g2.n[k].inside == 1) c++;
    if (c == npts2) {
      for (k = 0; k < g2.len; k++) { xo[k] = g2.n[k].x; yo[k] = g2.n[k].y; zo[k] = g2.n[k].z; }
      n_out = npts2;
    }
  }
  return n_out;
}

static void gc_ll2xyz(long n, const double *lon, const double *lat, double *x, double *y, double *z)
{
  long k;
  for (k = 0; k < n; k++) { x[k] = cos(lat[k])*cos(lon[k]); y[k] = cos(lat[k])*sin(lon[k]); z[k] = sin(lat[k]); }
}

/* the four corners of cell (i, j) in the reference's clockwise order (create_xgrid.c:1421-1427) */
static void gc_cell(const double *x, const double *y, const double *z, int nxp, int i, int j, double cx[4], double cy[4], double cz[4])
{
  const long n0 = (long)j*nxp+i, n1 = (long)(j+1)*nxp+i, n2 = (long)(j+1)*nxp+i+1, n3 = (long)j*nxp+i+1;
  cx[0] = x[n0]; cy[0] = y[n0]; cz[0] = z[n0];
  cx[1] = x[n1]; cy[1] = y[n1]; cz[1] = z[n1];
  cx[2] = x[n2]; cy[2] = y[n2]; cz[2] = z[n2];
  cx[3] = x[n3]; cy[3] = y[n3]; cz[3] = z[n3];
}

void orc_get_grid_great_circle_area(int nlon, int nlat, const double *lon, const double *lat, double *area)
{
  const int nxp = nlon+1, nyp = nlat+1;
  const long nv = (long)nxp*nyp;
  double *x = (double *)malloc(nv*sizeof(double)), *y = (double *)malloc(nv*sizeof(double)), *z = (double *)malloc(nv*sizeof(double));
  static GcList g;
  int i, j, k;
  gc_ll2xyz(nv, lon, lat, x, y, z);
  for (j = 0; j < nlat; j++) for (i = 0; i < nlon; i++) {
    double cx[4], cy[4], cz[4];
    gc_cell(x, y, z, nxp, i, j, cx, cy, cz);
    g.len = 0; g.overflow = 0;
    for (k = 0; k < 4; k++) gc_append_unique(&g, cx[k], cy[k], cz[k], 0, 0, 0, -1);
    area[(long)j*nlon+i] = gc_list_area(&g);
  }
  free(x); free(y); free(z);
}

/* returns the count; -1 capacity exceeded; <= -10: the reference would have aborted (-10 + clip error code) */
long orc_create_xgrid_great_circle(int nlon_in, int nlat_in, int nlon_out, int nlat_out,
                                   const double *lon_in, const double *lat_in,
                                   const double *lon_out, const double *lat_out, const double *mask_in,
                                   long cap, int *i_in, int *j_in, int *i_out, int *j_out,
                                   double *xarea, double *xclon, double *xclat)
{
  const int nx1 = nlon_in, ny1 = nlat_in, nx2 = nlon_out, ny2 = nlat_out, nx1p = nx1+1, nx2p = nx2+1;
  const long nv1 = (long)nx1p*(ny1+1), nv2 = (long)nx2p*(ny2+1);
  double *x1 = (double *)malloc(nv1*sizeof(double)), *y1 = (double *)malloc(nv1*sizeof(double)), *z1 = (double *)malloc(nv1*sizeof(double));
  double *x2 = (double *)malloc(nv2*sizeof(double)), *y2 = (double *)malloc(nv2*sizeof(double)), *z2 = (double *)malloc(nv2*sizeof(double));
  double *area1 = (double *)malloc((size_t)nx1*ny1*sizeof(double)), *area2 = (double *)malloc((size_t)nx2*ny2*sizeof(double));
  /* per destination cell coordinate ranges, so the six range rejections of clip_2dx2d_great_circle (:1508-1528) cost
   * six comparisons per pair instead of forty; the outcome of each test is unchanged */
  double *bb = (double *)malloc((size_t)nx2*ny2*6*sizeof(double));
  long nxgrid = 0;
  int i1, j1, i2, j2, k;

  gc_ll2xyz(nv1, lon_in, lat_in, x1, y1, z1);
  gc_ll2xyz(nv2, lon_out, lat_out, x2, y2, z2);
  orc_get_grid_great_circle_area(nlon_in, nlat_in, lon_in, lat_in, area1);
  orc_get_grid_great_circle_area(nlon_out, nlat_out, lon_out, lat_out, area2);
  for (j2 = 0; j2 < ny2; j2++) for (i2 = 0; i2 < nx2; i2++) {
    double cx[4], cy[4], cz[4], *b = bb + ((long)j2*nx2+i2)*6;
    gc_cell(x2, y2, z2, nx2p, i2, j2, cx, cy, cz);
    b[0] = b[1] = cx[0]; b[2] = b[3] = cy[0]; b[4] = b[5] = cz[0];
    for (k = 1; k < 4; k++) {
      if (cx[k] < b[0]) b[0] = cx[k]; if (cx[k] > b[1]) b[1] = cx[k];
      if (cy[k] < b[2]) b[2] = cy[k]; if (cy[k] > b[3]) b[3] = cy[k];
      if (cz[k] < b[4]) b[4] = cz[k]; if (cz[k] > b[5]) b[5] = cz[k];
    }
  }
  for (j1 = 0; j1 < ny1; j1++) for (i1 = 0; i1 < nx1; i1++) {
    double ax[4], ay[4], az[4], lo[3], hi[3];
    const double m = mask_in ? mask_in[(long)j1*nx1+i1] : 1.0;
    if (!(m > 0.5)) continue;
    gc_cell(x1, y1, z1, nx1p, i1, j1, ax, ay, az);
    lo[0] = hi[0] = ax[0]; lo[1] = hi[1] = ay[0]; lo[2] = hi[2] = az[0];
    for (k = 1; k < 4; k++) {
      if (ax[k] < lo[0]) lo[0] = ax[k]; if (ax[k] > hi[0]) hi[0] = ax[k];
      if (ay[k] < lo[1]) lo[1] = ay[k]; if (ay[k] > hi[1]) hi[1] = ay[k];
      if (az[k] < lo[2]) lo[2] = az[k]; if (az[k] > hi[2]) hi[2] = az[k];
    }
    for (j2 = 0; j2 < ny2; j2++) for (i2 = 0; i2 < nx2; i2++) {
      const double *b = bb + ((long)j2*nx2+i2)*6;
      double bx[4], by[4], bz[4], ox[GC_CAP], oy[GC_CAP], oz[GC_CAP], a, amin;
      int n_out;
      if (lo[0] >= b[1] + GC_RANGE || b[0] >= hi[0] + GC_RANGE || lo[1] >= b[3] + GC_RANGE || b[2] >= hi[1] + GC_RANGE ||
          lo[2] >= b[5] + GC_RANGE || b[4] >= hi[2] + GC_RANGE) continue;
      gc_cell(x2, y2, z2, nx2p, i2, j2, bx, by, bz);
      n_out = orc_clip_2dx2d_great_circle(ax, ay, az, 4, bx, by, bz, 4, ox, oy, oz);
      if (n_out < 0) { nxgrid = -10 + n_out; goto done; }
      if (n_out == 0) continue;
      a = orc_great_circle_area(n_out, ox, oy, oz)*m;
      amin = area1[(long)j1*nx1+i1] < area2[(long)j2*nx2+i2] ? area1[(long)j1*nx1+i1] : area2[(long)j2*nx2+i2];
      if (a/amin > 1.e-6) {
        if (nxgrid >= cap) { nxgrid = -1; goto done; }
        xarea[nxgrid] = a;
        if (xclon) xclon[nxgrid] = 0;
        if (xclat) xclat[nxgrid] = 0;
        i_in[nxgrid] = i1; j_in[nxgrid] = j1; i_out[nxgrid] = i2; j_out[nxgrid] = j2;
        nxgrid++;
      }
    }
  }
done:
  free(x1); free(y1); free(z1); free(x2); free(y2); free(z2); free(area1); free(area2); free(bb);
  return nxgrid;
}
