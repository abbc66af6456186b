/* Classic netCDF (CDF-1 / CDF-2 / CDF-5) reader and writer — see nc3.h.
 *
 * On-disk layout (netCDF classic format specification; what libnetcdf's nc__create(NC_CLOBBER | NC_CLASSIC_MODEL) and
 * NC_64BIT_OFFSET produce for the reference, mpp_io.c:163-175):
 *   magic "CDF" + version byte, numrecs, dim_list, gatt_list, var_list; all integers big-endian, names and attribute
 *   values padded with zeros to a multiple of 4; every variable carries (type, vsize, begin); fixed-size variables
 *   follow the header in definition order, record variables are interleaved record by record after them.
 */
#define _FILE_OFFSET_BITS 64
#define _GNU_SOURCE
#include "nc3.h"

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include "h5r.h"

enum { TAG_DIM = 0x0A, TAG_VAR = 0x0B, TAG_ATT = 0x0C };
#define NC3_MAX_LIST 65536u            /* dimensions / variables / attributes per list: a larger count is a corrupt header */

typedef struct { char *name; long long len; } Dim;
typedef struct { char *name; int type; long long n; unsigned char *raw; /* big-endian, as on disk, unpadded */ } Att;
typedef struct {
  char *name; int type; int ndims; int dimids[NC3_MAX_DIMS];
  int natts; Att *atts;
  long long vsize, begin; int isrec;
} Var;

struct nc3_file {
  struct h5r_file *h5; int *h5ds;                       /* a netCDF-4 file: the HDF5 reader and each variable's dataset */
  int fd, fmt, writing, defmode;
  long long numrecs;
  int ndims; Dim *dims;
  int ngatts; Att *gatts;
  int nvars; Var *vars;
  int unlim;
  long long recsize;
  char err[320];
};

static nc3_file *open_h5(const char *path, char *err, size_t errlen);
static int h5_slab(nc3_file *f, int varid, const size_t *start, const size_t *count, void *host, int kind);

static int tsize(int type)
{
  switch (type) {
    case NC3_BYTE: case NC3_CHAR: return 1;
    case NC3_SHORT: return 2;
    case NC3_INT: case NC3_FLOAT: return 4;
    case NC3_DOUBLE: return 8;
  }
  return 0;
}

static int fail(nc3_file *f, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(f->err, sizeof f->err, fmt, ap);
  va_end(ap);
  return -1;
}

const char *nc3_strerror(const nc3_file *f) { return f ? f->err : "nc3: no file"; }

/* ------------------------------------------------------------------------------------------------------------- */
/* big-endian primitives */
static void be_store(unsigned char *p, unsigned long long v, int nbytes)
{
  int i;
  for (i = nbytes - 1; i >= 0; --i) { p[i] = (unsigned char)(v & 0xff); v >>= 8; }
}
static unsigned long long be_load(const unsigned char *p, int nbytes)
{
  unsigned long long v = 0;
  int i;
  for (i = 0; i < nbytes; ++i) v = (v << 8) | p[i];
  return v;
}
static void swap_copy(void *dst, const void *src, size_t n, int ts)   /* n elements of ts bytes, byte order reversed */
{
  size_t i;
  const unsigned char *s = (const unsigned char *)src;
  unsigned char *d = (unsigned char *)dst;
  if (ts == 1) { memcpy(dst, src, n); return; }
  if (ts == 2) for (i = 0; i < n; ++i) { uint16_t v; memcpy(&v, s + 2 * i, 2); v = __builtin_bswap16(v); memcpy(d + 2 * i, &v, 2); }
  else if (ts == 4) for (i = 0; i < n; ++i) { uint32_t v; memcpy(&v, s + 4 * i, 4); v = __builtin_bswap32(v); memcpy(d + 4 * i, &v, 4); }
  else for (i = 0; i < n; ++i) { uint64_t v; memcpy(&v, s + 8 * i, 8); v = __builtin_bswap64(v); memcpy(d + 8 * i, &v, 8); }
}

/* one on-disk element (big-endian) -> double, and back */
static double disk_to_double(const unsigned char *p, int type)
{
  switch (type) {
    case NC3_BYTE: return (double)(signed char)p[0];
    case NC3_CHAR: return (double)p[0];
    case NC3_SHORT: return (double)(int16_t)be_load(p, 2);
    case NC3_INT: return (double)(int32_t)be_load(p, 4);
    case NC3_FLOAT: { uint32_t u = (uint32_t)be_load(p, 4); float v; memcpy(&v, &u, 4); return (double)v; }
    default: { uint64_t u = be_load(p, 8); double v; memcpy(&v, &u, 8); return v; }
  }
}
static void double_to_disk(unsigned char *p, int type, double v)
{
  switch (type) {
    case NC3_BYTE: case NC3_CHAR: p[0] = (unsigned char)(signed char)v; break;
    case NC3_SHORT: be_store(p, (unsigned long long)(uint16_t)(int16_t)v, 2); break;
    case NC3_INT: be_store(p, (unsigned long long)(uint32_t)(int32_t)v, 4); break;
    case NC3_FLOAT: { float x = (float)v; uint32_t u; memcpy(&u, &x, 4); be_store(p, u, 4); break; }
    default: { uint64_t u; memcpy(&u, &v, 8); be_store(p, u, 8); }
  }
}

/* ------------------------------------------------------------------------------------------------------------- */
/* header parsing */
typedef struct { const unsigned char *p; size_t n, at; int fmt; int short_read; } Cur;

static unsigned long long rd_u32(Cur *c)
{
  unsigned long long v;
  if (c->at + 4 > c->n) { c->short_read = 1; return 0; }
  v = be_load(c->p + c->at, 4); c->at += 4;
  return v;
}
static unsigned long long rd_nonneg(Cur *c)      /* NON_NEG: 4 bytes, 8 in CDF-5 */
{
  unsigned long long v;
  const int w = (c->fmt == 5) ? 8 : 4;
  if (c->at + w > c->n) { c->short_read = 1; return 0; }
  v = be_load(c->p + c->at, w); c->at += w;
  return v;
}
static char *rd_name(Cur *c)
{
  unsigned long long len = rd_nonneg(c);
  size_t padded = (size_t)((len + 3) & ~3ULL);
  char *s;
  if (c->short_read || len > (1u << 20) || c->at + padded > c->n) { c->short_read = 1; return NULL; }
  s = (char *)malloc(len + 1);
  memcpy(s, c->p + c->at, len); s[len] = 0;
  c->at += padded;
  return s;
}
static int rd_atts(Cur *c, int *natts, Att **atts)
{
  unsigned long long tag = rd_u32(c), n = rd_nonneg(c), i;
  *natts = 0; *atts = NULL;
  if (c->short_read) return 0;
  if (tag == 0 && n == 0) return 0;
  if (tag != TAG_ATT) return -1;
  if (n > NC3_MAX_LIST) return -5;
  *atts = (Att *)calloc(n ? n : 1, sizeof(Att));
  for (i = 0; i < n; ++i) {
    Att *a = &(*atts)[i];
    size_t bytes, padded;
    a->name = rd_name(c);
    a->type = (int)rd_u32(c);
    a->n = (long long)rd_nonneg(c);
    if (c->short_read) { *natts = (int)i; return 0; }
    if (!tsize(a->type)) return -2;
    /* a CDF-5 count is 64 bits: refuse what cannot be an attribute before the multiplication can wrap */
    if (a->n < 0 || a->n > (1LL << 31) / 8) return -5;
    bytes = (size_t)a->n * tsize(a->type);
    padded = (bytes + 3) & ~(size_t)3;
    if (c->at + padded > c->n) { c->short_read = 1; *natts = (int)i; return 0; }
    a->raw = (unsigned char *)malloc(bytes ? bytes : 1);
    memcpy(a->raw, c->p + c->at, bytes);
    c->at += padded;
    *natts = (int)i + 1;
  }
  return 0;
}

static void free_atts(int n, Att *a)
{
  int i;
  for (i = 0; i < n; ++i) { free(a[i].name); free(a[i].raw); }
  free(a);
}
static void free_meta(nc3_file *f)
{
  int i;
  for (i = 0; i < f->ndims; ++i) free(f->dims[i].name);
  free(f->dims); f->dims = NULL; f->ndims = 0;
  free_atts(f->ngatts, f->gatts); f->gatts = NULL; f->ngatts = 0;
  for (i = 0; i < f->nvars; ++i) { free(f->vars[i].name); free_atts(f->vars[i].natts, f->vars[i].atts); }
  free(f->vars); f->vars = NULL; f->nvars = 0;
}

/* returns 0 ok, 1 need more bytes, <0 malformed */
static int parse_header(nc3_file *f, const unsigned char *buf, size_t n)
{
  Cur c;
  unsigned long long tag, cnt, i;
  int k, rc;
  if (n < 8) return 1;
  c.p = buf; c.n = n; c.at = 4; c.fmt = buf[3]; c.short_read = 0;
  f->fmt = buf[3];
  f->numrecs = (long long)rd_nonneg(&c);
  if (f->fmt != 5 && f->numrecs == 0xFFFFFFFFLL) f->numrecs = -1;       /* STREAMING: derived from the file size below */
  tag = rd_u32(&c); cnt = rd_nonneg(&c);
  if (c.short_read) return 1;
  if (!(tag == 0 && cnt == 0)) {
    if (tag != TAG_DIM) return -1;
    if (cnt > NC3_MAX_LIST) return -5;
    f->dims = (Dim *)calloc(cnt ? cnt : 1, sizeof(Dim));
    for (i = 0; i < cnt; ++i) {
      f->dims[i].name = rd_name(&c);
      f->dims[i].len = (long long)rd_nonneg(&c);
      if (c.short_read) { f->ndims = (int)i; return 1; }
      f->ndims = (int)i + 1;
      if (f->dims[i].len == 0) f->unlim = (int)i;
    }
  }
  rc = rd_atts(&c, &f->ngatts, &f->gatts);
  if (rc < 0) return rc;
  if (c.short_read) return 1;
  tag = rd_u32(&c); cnt = rd_nonneg(&c);
  if (c.short_read) return 1;
  if (!(tag == 0 && cnt == 0)) {
    if (tag != TAG_VAR) return -1;
    if (cnt > NC3_MAX_LIST) return -5;
    f->vars = (Var *)calloc(cnt ? cnt : 1, sizeof(Var));
    for (i = 0; i < cnt; ++i) {
      Var *v = &f->vars[i];
      f->nvars = (int)i + 1;
      v->name = rd_name(&c);
      v->ndims = (int)rd_nonneg(&c);
      if (c.short_read) return 1;
      if (v->ndims > NC3_MAX_DIMS) return -3;
      for (k = 0; k < v->ndims; ++k) v->dimids[k] = (int)rd_nonneg(&c);
      rc = rd_atts(&c, &v->natts, &v->atts);
      if (rc < 0) return rc;
      v->type = (int)rd_u32(&c);
      v->vsize = (long long)rd_nonneg(&c);
      if (f->fmt == 1) v->begin = (long long)rd_u32(&c);
      else { if (c.at + 8 > c.n) c.short_read = 1; else { v->begin = (long long)be_load(c.p + c.at, 8); c.at += 8; } }
      if (c.short_read) return 1;
      if (!tsize(v->type)) return -2;
      for (k = 0; k < v->ndims; ++k) if (v->dimids[k] < 0 || v->dimids[k] >= f->ndims) return -4;
      v->isrec = (v->ndims > 0 && v->dimids[0] == f->unlim);
    }
  }
  return 0;
}

static long long var_fixed_elems(const nc3_file *f, const Var *v)   /* elements per record (record vars) or in total */
{
  long long n = 1;
  int k;
  for (k = v->isrec ? 1 : 0; k < v->ndims; ++k) {
    const long long len = f->dims[v->dimids[k]].len;
    if (len < 0 || (len > 0 && n > (long long)(0x7fffffffffffffffLL / 16) / len)) return -1;   /* product would overflow */
    n *= len;
  }
  return n;
}

static void compute_recsize(nc3_file *f)
{
  int i, nrec = 0;
  long long sz = 0;
  const Var *only = NULL;
  for (i = 0; i < f->nvars; ++i)
    if (f->vars[i].isrec) {
      long long b = var_fixed_elems(f, &f->vars[i]) * tsize(f->vars[i].type);
      sz += (b + 3) & ~3LL;
      only = &f->vars[i]; ++nrec;
    }
  /* a single record variable is not padded between records (classic format special case) */
  if (nrec == 1) sz = var_fixed_elems(f, only) * tsize(only->type);
  f->recsize = sz;
}

nc3_file *nc3_open(const char *path, char *err, size_t errlen)
{
  nc3_file *f = (nc3_file *)calloc(1, sizeof(nc3_file));
  size_t cap = 1 << 16;
  unsigned char *buf = NULL;
  struct stat st;
  int rc = 1;
  f->unlim = -1;
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) { if (err) snprintf(err, errlen, "nc3_open: cannot open %s: %s", path, strerror(errno)); free(f); return NULL; }
  fstat(f->fd, &st);
  while (rc == 1) {
    ssize_t got;
    buf = (unsigned char *)realloc(buf, cap);
    got = pread(f->fd, buf, cap, 0);
    if (got < 4) { if (err) snprintf(err, errlen, "nc3_open: %s is not a netCDF file (too short)", path); goto bad; }
    if (!memcmp(buf, "\211HDF", 4)) {                  /* netCDF-4: the reference's default format (mpp_io.c:52) */
      free(buf); close(f->fd); free(f);
      return open_h5(path, err, errlen);
    }
    if (memcmp(buf, "CDF", 3) || !(buf[3] == 1 || buf[3] == 2 || buf[3] == 5)) {
      unsigned char sig[8];
      long long off;
      for (off = 512; off + 8 <= (long long)st.st_size; off *= 2)     /* an HDF5 superblock behind a user block */
        if (pread(f->fd, sig, 8, off) == 8 && !memcmp(sig, "\211HDF\r\n\032\n", 8)) {
          free(buf); close(f->fd); free(f);
          return open_h5(path, err, errlen);
        }
      if (err) snprintf(err, errlen, "nc3_open: %s is not a classic netCDF file", path);
      goto bad;
    }
    free_meta(f); f->unlim = -1;
    rc = parse_header(f, buf, (size_t)got);
    if (rc == 1) {
      if ((size_t)got < cap) { if (err) snprintf(err, errlen, "nc3_open: %s: truncated header", path); goto bad; }
      cap *= 4;
    }
  }
  if (rc < 0) { if (err) snprintf(err, errlen, "nc3_open: %s: malformed header (%d)", path, rc); goto bad; }
  free(buf); buf = NULL;
  { int i; for (i = 0; i < f->nvars; ++i) if (var_fixed_elems(f, &f->vars[i]) < 0) {
      if (err) snprintf(err, errlen, "nc3_open: %s: variable %s: the product of its dimensions overflows", path, f->vars[i].name);
      goto bad; } }
  compute_recsize(f);
  if (f->numrecs < 0) {                                    /* streaming numrecs: infer from the file size */
    long long first = -1; int i;
    for (i = 0; i < f->nvars; ++i) if (f->vars[i].isrec && (first < 0 || f->vars[i].begin < first)) first = f->vars[i].begin;
    f->numrecs = (first >= 0 && f->recsize > 0) ? ((long long)st.st_size - first) / f->recsize : 0;
  }
  if (f->unlim >= 0) f->dims[f->unlim].len = 0;
  return f;
bad:
  free(buf);
  free_meta(f);
  close(f->fd);
  free(f);
  return NULL;
}

int nc3_format(const nc3_file *f) { return f->fmt; }
int nc3_ndims(const nc3_file *f) { return f->ndims; }
int nc3_nvars(const nc3_file *f) { return f->nvars; }
int nc3_unlimdim(const nc3_file *f) { return f->unlim; }
int nc3_dim_id(const nc3_file *f, const char *name)
{
  int i;
  for (i = 0; i < f->ndims; ++i) if (!strcmp(f->dims[i].name, name)) return i;
  return -1;
}
long long nc3_dim_len(const nc3_file *f, int d)
{
  if (d < 0 || d >= f->ndims) return -1;
  return (d == f->unlim) ? f->numrecs : f->dims[d].len;
}
const char *nc3_dim_name(const nc3_file *f, int d) { return (d < 0 || d >= f->ndims) ? NULL : f->dims[d].name; }
int nc3_var_id(const nc3_file *f, const char *name)
{
  int i;
  for (i = 0; i < f->nvars; ++i) if (!strcmp(f->vars[i].name, name)) return i;
  return -1;
}
const char *nc3_var_name(const nc3_file *f, int v) { return (v < 0 || v >= f->nvars) ? NULL : f->vars[v].name; }
int nc3_var_type(const nc3_file *f, int v) { return (v < 0 || v >= f->nvars) ? -1 : f->vars[v].type; }
int nc3_var_ndims(const nc3_file *f, int v) { return (v < 0 || v >= f->nvars) ? -1 : f->vars[v].ndims; }
const int *nc3_var_dimids(const nc3_file *f, int v) { return (v < 0 || v >= f->nvars) ? NULL : f->vars[v].dimids; }

static int att_list(const nc3_file *f, int varid, Att **a)
{
  if (varid == NC3_GLOBAL) { *a = f->gatts; return f->ngatts; }
  if (varid < 0 || varid >= f->nvars) { *a = NULL; return 0; }
  *a = f->vars[varid].atts; return f->vars[varid].natts;
}
int nc3_var_natts(const nc3_file *f, int varid) { Att *a; return att_list(f, varid, &a); }
const char *nc3_att_name(const nc3_file *f, int varid, int k)
{
  Att *a; int n = att_list(f, varid, &a);
  return (k < 0 || k >= n) ? NULL : a[k].name;
}
static const Att *find_att(const nc3_file *f, int varid, const char *name)
{
  Att *a; int n = att_list(f, varid, &a), i;
  for (i = 0; i < n; ++i) if (!strcmp(a[i].name, name)) return &a[i];
  return NULL;
}
int nc3_att_inq(const nc3_file *f, int varid, const char *name, int *type, long long *len)
{
  const Att *a = find_att(f, varid, name);
  if (!a) return -1;
  if (type) *type = a->type;
  if (len) *len = a->n;
  return 0;
}
int nc3_get_att_text(const nc3_file *f, int varid, const char *name, char *out, size_t outlen)
{
  const Att *a = find_att(f, varid, name);
  size_t n;
  if (!a || a->type != NC3_CHAR || outlen == 0) return -1;
  n = (size_t)a->n < outlen - 1 ? (size_t)a->n : outlen - 1;
  memcpy(out, a->raw, n); out[n] = 0;
  return 0;
}
int nc3_get_att_double(const nc3_file *f, int varid, const char *name, double *out, int maxn)
{
  const Att *a = find_att(f, varid, name);
  int i, n;
  if (!a || a->type == NC3_CHAR) return -1;
  n = a->n < maxn ? (int)a->n : maxn;
  for (i = 0; i < n; ++i) out[i] = disk_to_double(a->raw + (size_t)i * tsize(a->type), a->type);
  return n;
}

/* ------------------------------------------------------------------------------------------------------------- */
/* hyperslab access.  The slab is walked as rows of the innermost dimension; rows of the last-but-one dimension that
 * are close together on disk are moved through one window per system call. */
enum { K_DOUBLE, K_INT, K_TEXT };

static void conv_in(const unsigned char *disk, int type, size_t n, void *out, int kind)     /* disk -> host */
{
  size_t i;
  const int ts = tsize(type);
  if (kind == K_TEXT) { memcpy(out, disk, n); return; }
  if (kind == K_DOUBLE && type == NC3_DOUBLE) { swap_copy(out, disk, n, 8); return; }
  if (kind == K_INT && type == NC3_INT) { swap_copy(out, disk, n, 4); return; }
  for (i = 0; i < n; ++i) {
    const double v = disk_to_double(disk + i * ts, type);
    if (kind == K_DOUBLE) ((double *)out)[i] = v; else ((int *)out)[i] = (int)v;
  }
}
static void conv_out(unsigned char *disk, int type, size_t n, const void *in, int kind)     /* host -> disk */
{
  size_t i;
  const int ts = tsize(type);
  if (kind == K_TEXT) { memcpy(disk, in, n); return; }
  if (kind == K_DOUBLE && type == NC3_DOUBLE) { swap_copy(disk, in, n, 8); return; }
  if (kind == K_INT && type == NC3_INT) { swap_copy(disk, in, n, 4); return; }
  for (i = 0; i < n; ++i)
    double_to_disk(disk + i * ts, type, kind == K_DOUBLE ? ((const double *)in)[i] : (double)((const int *)in)[i]);
}

static int full_read(nc3_file *f, void *buf, size_t n, long long off, int allow_short)
{
  size_t done = 0;
  while (done < n) {
    ssize_t r = pread(f->fd, (char *)buf + done, n - done, off + (long long)done);
    if (r < 0) { if (errno == EINTR) continue; return fail(f, "nc3: read error: %s", strerror(errno)); }
    if (r == 0) {
      if (!allow_short) return fail(f, "nc3: unexpected end of file at offset %lld", off + (long long)done);
      memset((char *)buf + done, 0, n - done);
      break;
    }
    done += (size_t)r;
  }
  return 0;
}
static int full_write(nc3_file *f, const void *buf, size_t n, long long off)
{
  size_t done = 0;
  while (done < n) {
    ssize_t r = pwrite(f->fd, (const char *)buf + done, n - done, off + (long long)done);
    if (r < 0) { if (errno == EINTR) continue; return fail(f, "nc3: write error: %s", strerror(errno)); }
    done += (size_t)r;
  }
  return 0;
}

static int slab_io(nc3_file *f, int varid, const size_t *start, const size_t *count, void *host, int kind, int writing)
{
  const Var *v;
  int nd, k, ts, hs, appended = 0;
  size_t st[NC3_MAX_DIMS + 2], ct[NC3_MAX_DIMS + 2], idx[NC3_MAX_DIMS + 2];
  long long dl[NC3_MAX_DIMS + 2], stride[NC3_MAX_DIMS + 2];
  long long nouter, o;
  size_t rowlen, nrows, hpos = 0;
  const size_t WIN = (size_t)8 << 20;
  unsigned char *win;
  if (varid < 0 || varid >= f->nvars) return fail(f, "nc3: bad variable id %d", varid);
  if (f->h5) return writing ? fail(f, "nc3: file is open read-only") : h5_slab(f, varid, start, count, host, kind);
  if (f->writing && f->defmode) return fail(f, "nc3: still in define mode");
  if (writing && !f->writing) return fail(f, "nc3: file is open read-only");
  v = &f->vars[varid];
  ts = tsize(v->type);
  if ((kind == K_TEXT) != (v->type == NC3_CHAR)) return fail(f, "nc3: %s: text/numeric access mismatch", v->name);
  hs = (kind == K_DOUBLE) ? 8 : (kind == K_INT) ? 4 : 1;
  /* pad to at least two dimensions so that the (rows, row) logic below is general */
  nd = v->ndims;
  for (k = 0; k < nd; ++k) {
    dl[k] = (k == 0 && v->isrec) ? (writing ? (long long)1 << 62 : f->numrecs) : f->dims[v->dimids[k]].len;
    st[k] = start ? start[k] : 0;
    ct[k] = count ? count[k] : (size_t)dl[k];
    if ((long long)(st[k] + ct[k]) > dl[k]) return fail(f, "nc3: %s: start+count exceeds dimension %s", v->name, f->dims[v->dimids[k]].name);
  }
  if (nd == 1 && v->isrec) { dl[1] = 1; st[1] = 0; ct[1] = 1; nd = 2; appended = 1; }   /* (records, 1) */
  while (nd < 2) {                     /* prepend unit dimensions */
    for (k = nd; k > 0; --k) { dl[k] = dl[k - 1]; st[k] = st[k - 1]; ct[k] = ct[k - 1]; }
    dl[0] = 1; st[0] = 0; ct[0] = 1; ++nd;
  }
  {
    const int pad = nd - v->ndims - appended;     /* dimension `pad` is the variable's dimension 0 */
    long long s = ts;
    for (k = nd - 1; k >= 0; --k) { stride[k] = s; s *= (dl[k] > 0 ? dl[k] : 1); }
    if (v->isrec) stride[pad] = f->recsize;              /* records are recsize apart, whatever follows in the record */
    for (k = 0; k < pad; ++k) stride[k] = 0;
  }
  for (k = 0; k < nd; ++k) if (ct[k] == 0) return 0;
  rowlen = ct[nd - 1]; nrows = ct[nd - 2];
  nouter = 1;
  for (k = 0; k < nd - 2; ++k) { nouter *= (long long)ct[k]; idx[k] = 0; }
  win = (unsigned char *)malloc(WIN);
  if (!win) return fail(f, "nc3: out of memory");
  for (o = 0; o < nouter; ++o) {
    long long base = v->begin;
    size_t r = 0;
    for (k = 0; k < nd - 2; ++k) base += (long long)(st[k] + idx[k]) * stride[k];
    base += (long long)st[nd - 2] * stride[nd - 2] + (long long)st[nd - 1] * ts;
    while (r < nrows) {
      /* rows r .. r+m-1 through one window when they are near each other, else one row (in pieces) at a time */
      const long long rs = stride[nd - 2];
      size_t m = 1, j;
      if (rs > 0 && rowlen * ts <= WIN && (size_t)rs <= WIN) {
        m = (size_t)((WIN - rowlen * ts) / (size_t)rs) + 1;
        if (m > nrows - r) m = nrows - r;
      }
      if (rowlen * ts > WIN) {                            /* very long row: stream it */
        size_t e = 0;
        while (e < rowlen) {
          size_t ne = WIN / ts; if (ne > rowlen - e) ne = rowlen - e;
          if (writing) { conv_out(win, v->type, ne, (const char *)host + hpos * hs, kind);
                         if (full_write(f, win, ne * ts, base + (long long)r * rs + (long long)e * ts)) { free(win); return -1; } }
          else { if (full_read(f, win, ne * ts, base + (long long)r * rs + (long long)e * ts, 0)) { free(win); return -1; }
                 conv_in(win, v->type, ne, (char *)host + hpos * hs, kind); }
          hpos += ne; e += ne;
        }
      } else {
        const size_t span = (size_t)((long long)(m - 1) * rs) + rowlen * ts;
        const long long off = base + (long long)r * rs;
        const int dense = ((size_t)rs == rowlen * ts) || m == 1;
        if (writing) {
          if (!dense && full_read(f, win, span, off, 1)) { free(win); return -1; }    /* keep the bytes between rows */
          for (j = 0; j < m; ++j) conv_out(win + j * (size_t)rs, v->type, rowlen, (const char *)host + (hpos + j * rowlen) * hs, kind);
          if (full_write(f, win, span, off)) { free(win); return -1; }
        } else {
          if (full_read(f, win, span, off, 0)) { free(win); return -1; }
          for (j = 0; j < m; ++j) conv_in(win + j * (size_t)rs, v->type, rowlen, (char *)host + (hpos + j * rowlen) * hs, kind);
        }
        hpos += m * rowlen;
      }
      r += m;
    }
    for (k = nd - 3; k >= 0; --k) { if (++idx[k] < ct[k]) break; idx[k] = 0; }
  }
  free(win);
  if (writing && v->isrec) {
    const long long last = (long long)((start ? start[0] : 0) + (count ? count[0] : 0));
    if (last > f->numrecs) f->numrecs = last;
  }
  return 0;
}

int nc3_get_vara_double(nc3_file *f, int v, const size_t *s, const size_t *c, double *out) { return slab_io(f, v, s, c, out, K_DOUBLE, 0); }
int nc3_get_vara_int(nc3_file *f, int v, const size_t *s, const size_t *c, int *out) { return slab_io(f, v, s, c, out, K_INT, 0); }
int nc3_get_vara_text(nc3_file *f, int v, const size_t *s, const size_t *c, char *out) { return slab_io(f, v, s, c, out, K_TEXT, 0); }
int nc3_get_var_double(nc3_file *f, int v, double *out) { return slab_io(f, v, NULL, NULL, out, K_DOUBLE, 0); }
int nc3_get_var_int(nc3_file *f, int v, int *out) { return slab_io(f, v, NULL, NULL, out, K_INT, 0); }
int nc3_put_vara_double(nc3_file *f, int v, const size_t *s, const size_t *c, const double *in) { return slab_io(f, v, s, c, (void *)in, K_DOUBLE, 1); }
int nc3_put_vara_int(nc3_file *f, int v, const size_t *s, const size_t *c, const int *in) { return slab_io(f, v, s, c, (void *)in, K_INT, 1); }
int nc3_put_vara_text(nc3_file *f, int v, const size_t *s, const size_t *c, const char *in) { return slab_io(f, v, s, c, (void *)in, K_TEXT, 1); }
int nc3_put_var_double(nc3_file *f, int v, const double *in) { return slab_io(f, v, NULL, NULL, (void *)in, K_DOUBLE, 1); }
int nc3_put_var_int(nc3_file *f, int v, const int *in) { return slab_io(f, v, NULL, NULL, (void *)in, K_INT, 1); }

/* ------------------------------------------------------------------------------------------------------------- */
/* writing */
nc3_file *nc3_create(const char *path, int format, char *err, size_t errlen)
{
  nc3_file *f;
  if (!(format == 1 || format == 2 || format == 5)) { if (err) snprintf(err, errlen, "nc3_create: format must be 1, 2 or 5"); return NULL; }
  f = (nc3_file *)calloc(1, sizeof(nc3_file));
  f->fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (f->fd < 0) { if (err) snprintf(err, errlen, "nc3_create: cannot create %s: %s", path, strerror(errno)); free(f); return NULL; }
  f->fmt = format; f->writing = 1; f->defmode = 1; f->unlim = -1;
  return f;
}

int nc3_def_dim(nc3_file *f, const char *name, long long len)
{
  if (!f->writing || !f->defmode) return fail(f, "nc3_def_dim: not in define mode");
  if (len == 0 && f->unlim >= 0) return fail(f, "nc3_def_dim: only one unlimited dimension is allowed");
  if (nc3_dim_id(f, name) >= 0) return fail(f, "nc3_def_dim: %s is already defined", name);
  f->dims = (Dim *)realloc(f->dims, (size_t)(f->ndims + 1) * sizeof(Dim));
  f->dims[f->ndims].name = strdup(name);
  f->dims[f->ndims].len = len;
  if (len == 0) f->unlim = f->ndims;
  return f->ndims++;
}

int nc3_def_var(nc3_file *f, const char *name, int type, int ndims, const int *dimids)
{
  Var *v;
  int k;
  if (!f->writing || !f->defmode) return fail(f, "nc3_def_var: not in define mode");
  if (!tsize(type) || ndims < 0 || ndims > NC3_MAX_DIMS) return fail(f, "nc3_def_var: %s: bad type or rank", name);
  for (k = 0; k < ndims; ++k) {
    if (dimids[k] < 0 || dimids[k] >= f->ndims) return fail(f, "nc3_def_var: %s: bad dimension id", name);
    if (k > 0 && dimids[k] == f->unlim) return fail(f, "nc3_def_var: %s: the unlimited dimension must come first", name);
  }
  f->vars = (Var *)realloc(f->vars, (size_t)(f->nvars + 1) * sizeof(Var));
  v = &f->vars[f->nvars];
  memset(v, 0, sizeof *v);
  v->name = strdup(name); v->type = type; v->ndims = ndims;
  for (k = 0; k < ndims; ++k) v->dimids[k] = dimids[k];
  v->isrec = (ndims > 0 && dimids[0] == f->unlim);
  return f->nvars++;
}

static Att *new_att(nc3_file *f, int varid, const char *name)
{
  Att **list; int *n; int i;
  if (varid == NC3_GLOBAL) { list = &f->gatts; n = &f->ngatts; }
  else if (varid >= 0 && varid < f->nvars) { list = &f->vars[varid].atts; n = &f->vars[varid].natts; }
  else return NULL;
  for (i = 0; i < *n; ++i) if (!strcmp((*list)[i].name, name)) { free((*list)[i].raw); (*list)[i].raw = NULL; return &(*list)[i]; }
  *list = (Att *)realloc(*list, (size_t)(*n + 1) * sizeof(Att));
  memset(&(*list)[*n], 0, sizeof(Att));
  (*list)[*n].name = strdup(name);
  return &(*list)[(*n)++];
}

int nc3_put_att_text(nc3_file *f, int varid, const char *name, const char *text)
{
  Att *a;
  if (!f->writing || !f->defmode) return fail(f, "nc3_put_att_text: not in define mode");
  a = new_att(f, varid, name);
  if (!a) return fail(f, "nc3_put_att_text: bad variable id");
  a->type = NC3_CHAR; a->n = (long long)strlen(text);
  a->raw = (unsigned char *)malloc((size_t)a->n + 1);
  memcpy(a->raw, text, (size_t)a->n);
  return 0;
}

int nc3_put_att_double(nc3_file *f, int varid, const char *name, int type, int n, const double *vals)
{
  Att *a;
  int i;
  if (!f->writing || !f->defmode) return fail(f, "nc3_put_att_double: not in define mode");
  if (!tsize(type) || type == NC3_CHAR) return fail(f, "nc3_put_att_double: bad type");
  a = new_att(f, varid, name);
  if (!a) return fail(f, "nc3_put_att_double: bad variable id");
  a->type = type; a->n = n;
  a->raw = (unsigned char *)malloc((size_t)n * tsize(type) + 1);
  for (i = 0; i < n; ++i) double_to_disk(a->raw + (size_t)i * tsize(type), type, vals[i]);
  return 0;
}

int nc3_copy_att(const nc3_file *fin, int varid_in, const char *name, nc3_file *fout, int varid_out)
{
  const Att *a = find_att(fin, varid_in, name);
  Att *b;
  if (!a) return fail(fout, "nc3_copy_att: no attribute %s", name);
  if (!fout->writing || !fout->defmode) return fail(fout, "nc3_copy_att: not in define mode");
  b = new_att(fout, varid_out, name);
  if (!b) return fail(fout, "nc3_copy_att: bad variable id");
  b->type = a->type; b->n = a->n;
  b->raw = (unsigned char *)malloc((size_t)a->n * tsize(a->type) + 1);
  memcpy(b->raw, a->raw, (size_t)a->n * tsize(a->type));
  return 0;
}

int nc3_copy_atts(const nc3_file *fin, int varid_in, nc3_file *fout, int varid_out)
{
  Att *a;
  const int n = att_list(fin, varid_in, &a);
  int i;
  for (i = 0; i < n; ++i) if (nc3_copy_att(fin, varid_in, a[i].name, fout, varid_out)) return -1;
  return 0;
}

/* header serialisation */
typedef struct { unsigned char *p; size_t n, cap; int fmt; } Out;
static void o_need(Out *o, size_t k) { if (o->n + k > o->cap) { o->cap = (o->n + k) * 2 + 256; o->p = (unsigned char *)realloc(o->p, o->cap); } }
static void o_u32(Out *o, unsigned long long v) { o_need(o, 4); be_store(o->p + o->n, v, 4); o->n += 4; }
static void o_nonneg(Out *o, unsigned long long v) { const int w = (o->fmt == 5) ? 8 : 4; o_need(o, 8); be_store(o->p + o->n, v, w); o->n += w; }
static void o_bytes(Out *o, const void *b, size_t len)
{
  const size_t padded = (len + 3) & ~(size_t)3;
  o_need(o, padded);
  memcpy(o->p + o->n, b, len);
  memset(o->p + o->n + len, 0, padded - len);
  o->n += padded;
}
static void o_name(Out *o, const char *s) { o_nonneg(o, strlen(s)); o_bytes(o, s, strlen(s)); }
static void o_atts(Out *o, int n, const Att *a)
{
  int i;
  if (n == 0) { o_u32(o, 0); o_nonneg(o, 0); return; }
  o_u32(o, TAG_ATT); o_nonneg(o, (unsigned long long)n);
  for (i = 0; i < n; ++i) {
    o_name(o, a[i].name);
    o_u32(o, (unsigned long long)a[i].type);
    o_nonneg(o, (unsigned long long)a[i].n);
    o_bytes(o, a[i].raw, (size_t)a[i].n * tsize(a[i].type));
  }
}
static void serialise(const nc3_file *f, Out *o)
{
  int i, k;
  o->n = 0;
  o_need(o, 4);
  memcpy(o->p, "CDF", 3); o->p[3] = (unsigned char)f->fmt; o->n = 4;
  o_nonneg(o, (unsigned long long)f->numrecs);
  if (f->ndims == 0) { o_u32(o, 0); o_nonneg(o, 0); }
  else {
    o_u32(o, TAG_DIM); o_nonneg(o, (unsigned long long)f->ndims);
    for (i = 0; i < f->ndims; ++i) { o_name(o, f->dims[i].name); o_nonneg(o, (unsigned long long)f->dims[i].len); }
  }
  o_atts(o, f->ngatts, f->gatts);
  if (f->nvars == 0) { o_u32(o, 0); o_nonneg(o, 0); }
  else {
    o_u32(o, TAG_VAR); o_nonneg(o, (unsigned long long)f->nvars);
    for (i = 0; i < f->nvars; ++i) {
      const Var *v = &f->vars[i];
      unsigned long long vs = (unsigned long long)v->vsize;
      o_name(o, v->name);
      o_nonneg(o, (unsigned long long)v->ndims);
      for (k = 0; k < v->ndims; ++k) o_nonneg(o, (unsigned long long)v->dimids[k]);
      o_atts(o, v->natts, v->atts);
      o_u32(o, (unsigned long long)v->type);
      if (f->fmt != 5 && vs > 0xFFFFFFFCULL) vs = 0xFFFFFFFFULL;      /* "vsize is not used when it does not fit" */
      o_nonneg(o, vs);
      if (f->fmt == 1) o_u32(o, (unsigned long long)v->begin);
      else { o_need(o, 8); be_store(o->p + o->n, (unsigned long long)v->begin, 8); o->n += 8; }
    }
  }
}

int nc3_enddef(nc3_file *f)
{
  Out o = {NULL, 0, 0, 0};
  int i;
  long long at;
  if (!f->writing || !f->defmode) return fail(f, "nc3_enddef: not in define mode");
  o.fmt = f->fmt;
  for (i = 0; i < f->nvars; ++i) {
    Var *v = &f->vars[i];
    const long long b = var_fixed_elems(f, v) * tsize(v->type);
    v->vsize = (b + 3) & ~3LL;
    v->begin = 0;
  }
  compute_recsize(f);
  {  /* a lone record variable keeps its unpadded size as vsize too (libnetcdf does the same) */
    int nrec = 0; Var *only = NULL;
    for (i = 0; i < f->nvars; ++i) if (f->vars[i].isrec) { ++nrec; only = &f->vars[i]; }
    if (nrec == 1) only->vsize = var_fixed_elems(f, only) * tsize(only->type);
  }
  serialise(f, &o);                                      /* the header size does not depend on the offsets */
  at = (long long)o.n;
  for (i = 0; i < f->nvars; ++i) if (!f->vars[i].isrec) { f->vars[i].begin = at; at += f->vars[i].vsize; }
  for (i = 0; i < f->nvars; ++i) if (f->vars[i].isrec) { f->vars[i].begin = at; at += (f->vars[i].vsize + 3) & ~3LL; }
  if (f->fmt == 1)
    for (i = 0; i < f->nvars; ++i)
      if (f->vars[i].begin > 0x7FFFFFFFLL) { free(o.p); return fail(f, "nc3_enddef: %s starts beyond 2 GiB; use format 2 or 5", f->vars[i].name); }
  serialise(f, &o);
  f->defmode = 0;
  i = full_write(f, o.p, o.n, 0);
  free(o.p);
  if (i) return -1;
  {  /* size the fixed part so that short writes later leave a well-formed file */
    long long fixed_end = (long long)o.n;
    int j;
    for (j = 0; j < f->nvars; ++j) if (!f->vars[j].isrec && f->vars[j].begin + f->vars[j].vsize > fixed_end) fixed_end = f->vars[j].begin + f->vars[j].vsize;
    if (ftruncate(f->fd, fixed_end)) return fail(f, "nc3_enddef: cannot size the file: %s", strerror(errno));
  }
  return 0;
}

int nc3_close(nc3_file *f)
{
  int rc = 0;
  if (!f) return 0;
  if (f->writing) {
    if (f->defmode) rc = nc3_enddef(f);
    if (!rc && f->unlim >= 0) {                          /* numrecs lives right after the magic */
      unsigned char b[8];
      const int w = (f->fmt == 5) ? 8 : 4;
      be_store(b, (unsigned long long)f->numrecs, w);
      rc = full_write(f, b, (size_t)w, 4);
      if (!rc && f->recsize > 0) {                       /* records that were never written read back as zeros */
        long long first = -1; int i; struct stat st;
        for (i = 0; i < f->nvars; ++i) if (f->vars[i].isrec && (first < 0 || f->vars[i].begin < first)) first = f->vars[i].begin;
        fstat(f->fd, &st);
        if (first >= 0 && (long long)st.st_size < first + f->numrecs * f->recsize)
          if (ftruncate(f->fd, first + f->numrecs * f->recsize)) rc = -1;
      }
    }
  }
  if (f->h5) { h5r_close(f->h5); free(f->h5ds); }
  else if (close(f->fd)) rc = -1;
  free_meta(f);
  free(f);
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------- */
/* netCDF-4 files.  The reference opens every file through libnetcdf and writes NC_FORMAT_NETCDF4_CLASSIC by default
 * (tools/libfrencutils/mpp_io.c:52, :109-140, :163-169); h5r.c reads the HDF5 container, this part lays the netCDF-4
 * conventions over it (netcdf-c libhdf5/hdf5open.c, nc4hdf.c): a dimension is a dimension-scale dataset (attribute CLASS =
 * "DIMENSION_SCALE"; NAME starts with "This is a netCDF dimension but not a netCDF variable." when no coordinate variable
 * goes with it; _Netcdf4Dimid holds its id), a variable's dimensions are the object references of DIMENSION_LIST, a variable
 * that shares a dimension's name without being its coordinate variable is stored as "_nc4_non_coord_<name>", and the
 * bookkeeping attributes (CLASS, NAME, DIMENSION_LIST, REFERENCE_LIST, _Netcdf4Dimid, _Netcdf4Coordinates, _nc3_strict,
 * _NCProperties) are not netCDF attributes. */
static const char NOT_A_VAR[] = "This is a netCDF dimension but not a netCDF variable.";
static const char NON_COORD[] = "_nc4_non_coord_";

static const h5r_att *h5_find_att(const h5r_dset *d, const char *name)
{
  int k;
  for (k = 0; k < d->natts; ++k) if (!strcmp(d->atts[k].name, name)) return &d->atts[k];
  return NULL;
}
static int h5_hidden_att(const char *n)
{
  static const char *const hidden[] = { "CLASS", "NAME", "DIMENSION_LIST", "REFERENCE_LIST", "_Netcdf4Dimid", "_Netcdf4Coordinates",
                                        "_nc3_strict", "_NCProperties", "_Netcdf4Inq" };
  size_t k;
  for (k = 0; k < sizeof hidden / sizeof hidden[0]; ++k) if (!strcmp(n, hidden[k])) return 1;
  return 0;
}
/* the classic type that holds every value of an HDF5 atomic type (int64 and the wide unsigned types: double) */
static int h5_nc_type(const h5r_type *t)
{
  if (t->cls == H5R_STRING || t->cls == H5R_VLEN_STRING) return NC3_CHAR;
  if (t->cls == H5R_FLOAT) return t->size == 4 ? NC3_FLOAT : t->size == 8 ? NC3_DOUBLE : 0;
  if (t->cls == H5R_INT) {
    if (t->size == 1) return t->is_signed ? NC3_BYTE : NC3_SHORT;
    if (t->size == 2) return t->is_signed ? NC3_SHORT : NC3_INT;
    if (t->size == 4) return t->is_signed ? NC3_INT : NC3_DOUBLE;
    if (t->size == 8) return NC3_DOUBLE;
  }
  return 0;
}
/* one element in host byte order -> double */
static double h5_elem(const unsigned char *p, const h5r_type *t)
{
  if (t->cls == H5R_FLOAT) {
    if (t->size == 4) { float v; memcpy(&v, p, 4); return (double)v; }
    { double v; memcpy(&v, p, 8); return v; }
  }
  switch (t->size) {
    case 1: return t->is_signed ? (double)(signed char)p[0] : (double)p[0];
    case 2: { uint16_t v; memcpy(&v, p, 2); return t->is_signed ? (double)(int16_t)v : (double)v; }
    case 4: { uint32_t v; memcpy(&v, p, 4); return t->is_signed ? (double)(int32_t)v : (double)v; }
    default: { uint64_t v; memcpy(&v, p, 8); return t->is_signed ? (double)(int64_t)v : (double)v; }
  }
}
static void h5_swap_elems(unsigned char *p, size_t n, int es)   /* attribute bytes come as stored */
{
  size_t i; int b;
  for (i = 0; i < n; ++i) for (b = 0; b < es / 2; ++b) { unsigned char t = p[i * es + b]; p[i * es + b] = p[i * es + es - 1 - b]; p[i * es + es - 1 - b] = t; }
}

static int h5_add_att(int *natts, Att **atts, const h5r_att *a)
{
  Att *o;
  const int type = h5_nc_type(&a->type);
  if (!type) return 0;                                   /* compound, enum, references ...: not a classic attribute */
  *atts = (Att *)realloc(*atts, (size_t)(*natts + 1) * sizeof(Att));
  if (!*atts) return -1;
  o = &(*atts)[*natts];
  o->name = strdup(a->name); o->type = type;
  if (a->type.cls == H5R_STRING) {                       /* a text attribute is ONE fixed-length string (nc4hdf.c put_att_grpa) */
    const size_t n = a->data_len;
    o->n = (long long)n; o->raw = (unsigned char *)malloc(n + 1); memcpy(o->raw, a->data, n); o->raw[n] = 0;
  } else if (a->type.cls == H5R_VLEN_STRING) {           /* NC_STRING: the first string, as text */
    const size_t n = a->data_len ? strlen((const char *)a->data) : 0;
    o->n = (long long)n; o->raw = (unsigned char *)malloc(n + 1); memcpy(o->raw, a->data, n); o->raw[n] = 0;
  } else {
    long long i;
    const int ts = tsize(type), es = a->type.size;
    unsigned char *tmp = (unsigned char *)malloc(a->data_len ? a->data_len : 1);
    memcpy(tmp, a->data, a->data_len);
    if (a->type.big_endian) h5_swap_elems(tmp, (size_t)a->nelem, es);      /* -> host order */
    o->n = a->nelem; o->raw = (unsigned char *)malloc((size_t)(a->nelem > 0 ? a->nelem : 1) * (size_t)ts);
    for (i = 0; i < a->nelem; ++i) double_to_disk(o->raw + i * ts, type, h5_elem(tmp + i * es, &a->type));
    free(tmp);
  }
  ++*natts;
  return 0;
}

static nc3_file *open_h5(const char *path, char *err, size_t errlen)
{
  nc3_file *f = (nc3_file *)calloc(1, sizeof(nc3_file));
  int nds, i, k, *dim_of_ds = NULL, *is_var = NULL, ok = 0;
  if (!f) { if (err) snprintf(err, errlen, "nc3_open: out of memory"); return NULL; }
  f->fd = -1; f->unlim = -1; f->fmt = 4;
  f->h5 = h5r_open(path, err, errlen);
  if (!f->h5) { free(f); return NULL; }
  nds = h5r_ndsets(f->h5);
  dim_of_ds = (int *)malloc((size_t)(nds + 1) * sizeof(int));
  is_var = (int *)calloc((size_t)(nds + 1), sizeof(int));
  f->dims = (Dim *)calloc((size_t)(nds + 1) * (NC3_MAX_DIMS + 1), sizeof(Dim));     /* scales + phony dimensions */
  f->vars = (Var *)calloc((size_t)(nds + 1), sizeof(Var));
  f->h5ds = (int *)malloc((size_t)(nds + 1) * sizeof(int));
  /* dimensions: the dimension scales, numbered by _Netcdf4Dimid where the file carries it, else in creation order */
  {
    int nscale = 0, have_ids = 1;
    for (i = 0; i < nds; ++i) {
      const h5r_dset *d = h5r_dset_at(f->h5, i);
      const h5r_att *cls = h5_find_att(d, "CLASS");
      dim_of_ds[i] = -1;
      if (cls && cls->type.cls == H5R_STRING && cls->data_len >= 15 && !strncmp((const char *)cls->data, "DIMENSION_SCALE", 15) && d->rank >= 1) {
        const h5r_att *id = h5_find_att(d, "_Netcdf4Dimid");
        dim_of_ds[i] = nscale++;
        if (!id || id->type.cls != H5R_INT || id->nelem < 1) have_ids = 0;
      }
    }
    if (have_ids)
      for (i = 0; i < nds; ++i) if (dim_of_ds[i] >= 0) {
        const h5r_att *id = h5_find_att(h5r_dset_at(f->h5, i), "_Netcdf4Dimid");
        unsigned char tmp[8] = {0};
        memcpy(tmp, id->data, (size_t)(id->type.size <= 8 ? id->type.size : 8));
        if (id->type.big_endian) h5_swap_elems(tmp, 1, id->type.size);
        dim_of_ds[i] = (int)h5_elem(tmp, &id->type);
        if (dim_of_ds[i] < 0 || dim_of_ds[i] >= nscale) { if (err) snprintf(err, errlen, "nc3_open: %s: _Netcdf4Dimid %d of %s is out of range", path, dim_of_ds[i], h5r_dset_at(f->h5, i)->name); goto done; }
      }
    f->ndims = nscale;
    for (i = 0; i < nds; ++i) if (dim_of_ds[i] >= 0) {
      const h5r_dset *d = h5r_dset_at(f->h5, i);
      const h5r_att *nm = h5_find_att(d, "NAME");
      Dim *dm = &f->dims[dim_of_ds[i]];
      const char *name = d->name;
      if (dm->name) { if (err) snprintf(err, errlen, "nc3_open: %s: two dimensions carry the id %d", path, dim_of_ds[i]); goto done; }
      if (!strncmp(name, NON_COORD, sizeof NON_COORD - 1)) name += sizeof NON_COORD - 1;
      dm->name = strdup(name); dm->len = d->dims[0];
      if (d->maxdims[0] < 0 && f->unlim < 0) f->unlim = dim_of_ds[i];
      is_var[i] = !(nm && nm->type.cls == H5R_STRING && nm->data_len >= sizeof NOT_A_VAR - 1 &&
                    !strncmp((const char *)nm->data, NOT_A_VAR, sizeof NOT_A_VAR - 1));
    } else is_var[i] = 1;
  }
  /* variables, in creation order (h5r sorts the links by it) */
  for (i = 0; i < nds; ++i) {
    const h5r_dset *d = h5r_dset_at(f->h5, i);
    const h5r_att *dl = h5_find_att(d, "DIMENSION_LIST");
    Var *v;
    const char *name = d->name;
    if (!is_var[i]) continue;
    v = &f->vars[f->nvars];
    v->type = h5_nc_type(&d->type);
    if (!v->type || d->rank > NC3_MAX_DIMS) continue;      /* not expressible in the classic model: invisible, like a group */
    if (!strncmp(name, NON_COORD, sizeof NON_COORD - 1)) name += sizeof NON_COORD - 1;
    v->name = strdup(name); v->ndims = d->rank;
    for (k = 0; k < d->rank; ++k) {
      int dimid = -1, j;
      if (dim_of_ds[i] >= 0 && d->rank == 1) dimid = dim_of_ds[i];                   /* a coordinate variable is its own scale */
      else if (dl && k < dl->ndimrefs)
        for (j = 0; j < nds; ++j) if (dim_of_ds[j] >= 0 && h5r_dset_at(f->h5, j)->addr == dl->dimrefs[k]) { dimid = dim_of_ds[j]; break; }
      if (dimid < 0) {                                     /* no scale attached (a plain HDF5 file): a dimension per length */
        char nm[48];
        for (j = 0; j < f->ndims; ++j) if (!strncmp(f->dims[j].name, "phony_dim_", 10) && f->dims[j].len == d->dims[k]) { dimid = j; break; }
        if (dimid < 0) { dimid = f->ndims++; snprintf(nm, sizeof nm, "phony_dim_%d", dimid); f->dims[dimid].name = strdup(nm); f->dims[dimid].len = d->dims[k]; }
      }
      v->dimids[k] = dimid;
      if (dimid == f->unlim && d->dims[k] > f->dims[dimid].len) f->dims[dimid].len = d->dims[k];
    }
    v->isrec = (v->ndims > 0 && v->dimids[0] == f->unlim);
    for (k = 0; k < d->natts; ++k) if (!h5_hidden_att(d->atts[k].name) && h5_add_att(&v->natts, &v->atts, &d->atts[k])) goto done;
    f->h5ds[f->nvars++] = i;
  }
  for (k = 0; k < h5r_ngatts(f->h5); ++k) {
    const h5r_att *a = h5r_gatt_at(f->h5, k);
    if (!h5_hidden_att(a->name) && h5_add_att(&f->ngatts, &f->gatts, a)) goto done;
  }
  if (f->unlim >= 0) { f->numrecs = f->dims[f->unlim].len; f->dims[f->unlim].len = 0; }
  ok = 1;
done:
  free(dim_of_ds); free(is_var);
  if (!ok) { h5r_close(f->h5); free(f->h5ds); free_meta(f); free(f); return NULL; }
  return f;
}

static int h5_slab(nc3_file *f, int varid, const size_t *start, const size_t *count, void *host, int kind)
{
  const Var *v = &f->vars[varid];
  const h5r_dset *d = h5r_dset_at(f->h5, f->h5ds[varid]);
  size_t total = 1, i, st[NC3_MAX_DIMS], ct[NC3_MAX_DIMS];
  int k;
  unsigned char *tmp;
  if ((kind == K_TEXT) != (v->type == NC3_CHAR)) return fail(f, "nc3: %s: text/numeric access mismatch", v->name);
  /* the request in terms of the netCDF dimensions (what the caller sized its buffer by); h5r_read checks it against the
   * dataset's own extent */
  for (k = 0; k < v->ndims; ++k) {
    const long long len = (v->dimids[k] == f->unlim) ? f->numrecs : f->dims[v->dimids[k]].len;
    st[k] = start ? start[k] : 0;
    ct[k] = count ? count[k] : (size_t)len;
    if ((long long)(st[k] + ct[k]) > len) return fail(f, "nc3: %s: start+count exceeds dimension %s", v->name, f->dims[v->dimids[k]].name);
    total *= ct[k];
  }
  start = st; count = ct;
  if (total == 0) return 0;
  if (kind == K_TEXT) {
    if (d->type.cls != H5R_STRING || d->type.size != 1) return fail(f, "nc3: %s: text stored as strings of %d bytes is not a classic char variable", v->name, d->type.size);
    if (h5r_read(f->h5, f->h5ds[varid], start, count, host)) return fail(f, "%s", h5r_strerror(f->h5));
    return 0;
  }
  if (kind == K_DOUBLE && d->type.cls == H5R_FLOAT && d->type.size == 8) {              /* the common case, in place */
    if (h5r_read(f->h5, f->h5ds[varid], start, count, host)) return fail(f, "%s", h5r_strerror(f->h5));
    return 0;
  }
  if (kind == K_INT && d->type.cls == H5R_INT && d->type.size == 4 && d->type.is_signed) {
    if (h5r_read(f->h5, f->h5ds[varid], start, count, host)) return fail(f, "%s", h5r_strerror(f->h5));
    return 0;
  }
  tmp = (unsigned char *)malloc(total * (size_t)d->type.size);
  if (!tmp) return fail(f, "nc3: out of memory");
  if (h5r_read(f->h5, f->h5ds[varid], start, count, tmp)) { free(tmp); return fail(f, "%s", h5r_strerror(f->h5)); }
  for (i = 0; i < total; ++i) {
    const double x = h5_elem(tmp + i * (size_t)d->type.size, &d->type);
    if (kind == K_DOUBLE) ((double *)host)[i] = x; else ((int *)host)[i] = (int)x;
  }
  free(tmp);
  return 0;
}
