/* Read-only HDF5 subset for netCDF-4 files — see h5r.h for scope.  Written from the HDF5 File Format Specification 3.0;
 * section names in the comments are that document's ("III.A. Disk Format: Level 1A1 - Version 1 B-trees", ...).
 * Every structure is bounds-checked against the bytes actually read: a corrupt file gives an error, not a wild read. */
#define _FILE_OFFSET_BITS 64
#define _GNU_SOURCE
#include "h5r.h"

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
#include <zlib.h>

typedef unsigned long long u64;
#define UNDEF (~0ULL)
#define MAX_FILTERS 8
#define MAX_OBJ_BLOCKS 256

typedef struct { u64 off[H5R_MAX_DIMS]; u64 addr; unsigned size, mask; } Chunk;
typedef struct {
  int layout;                                   /* 0 compact, 1 contiguous, 2 chunked */
  u64 addr, size;                               /* contiguous */
  unsigned char *compact; size_t compact_len;
  int cnd; u64 btree; u64 cdims[H5R_MAX_DIMS + 1];   /* chunk shape (rank entries) then the element size */
  int nfilters; struct { int id; int ncd; unsigned cd[8]; } filt[MAX_FILTERS];
  unsigned char *fill; int fill_len;
  int indexed; long nchunks; Chunk *chunks;
  unsigned char *cache; long cache_chunk;       /* last decoded chunk */
} DsPriv;

struct h5r_file {
  int fd;
  u64 base, eof, fsize;
  int so, sl;                                   /* size of offsets / lengths */
  int ndsets; h5r_dset *dsets;
  int ngatts; h5r_att *gatts;
  unsigned char *gcol; u64 gcol_addr; size_t gcol_len;   /* last global heap collection */
  long budget;                                  /* index nodes / heap blocks one walk may still visit (a damaged file can
                                                   point a tree back at itself) */
  char err[256];
};

typedef struct { int type; unsigned flags; size_t size; const unsigned char *data; } Msg;
typedef struct { int n, cap; Msg *m; int nblk; unsigned char *blk[MAX_OBJ_BLOCKS]; } MsgList;
typedef struct { char *name; u64 addr; long long corder; } Link;
typedef struct { int n, cap; Link *l; } LinkList;

const char *h5r_strerror(const h5r_file *f) { return f ? f->err : "h5r: no file"; }

static int fail(h5r_file *f, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(f->err, sizeof f->err, fmt, ap);
  va_end(ap);
  return -1;
}

static u64 le(const unsigned char *p, int n)
{
  u64 v = 0;
  int i;
  for (i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
  if (n < 8 && n > 0) {                          /* an all-ones field of any width is "undefined" */
    const u64 ones = (1ULL << (8 * n)) - 1;
    if (v == ones) return UNDEF;
  }
  return v;
}
static u64 le_raw(const unsigned char *p, int n)   /* plain little-endian (sizes, counts) */
{
  u64 v = 0;
  int i;
  for (i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}

/* n bytes at file address `addr` (relative to the base address) */
static int rd(h5r_file *f, u64 addr, void *buf, size_t n)
{
  size_t done = 0;
  if (addr == UNDEF || addr + f->base + n > f->fsize || addr + f->base + n < addr)
    return fail(f, "h5r: read of %zu bytes at address %llu is outside the file", n, addr);
  while (done < n) {
    ssize_t r = pread(f->fd, (char *)buf + done, n - done, (off_t)(f->base + addr + done));
    if (r < 0) { if (errno == EINTR) continue; return fail(f, "h5r: read error: %s", strerror(errno)); }
    if (r == 0) return fail(f, "h5r: unexpected end of file at address %llu", addr + done);
    done += (size_t)r;
  }
  return 0;
}
/* up to n bytes (structures whose length is learnt from their first bytes); returns bytes read */
static size_t rd_upto(h5r_file *f, u64 addr, void *buf, size_t n)
{
  if (addr == UNDEF || addr + f->base >= f->fsize) return 0;
  if (addr + f->base + n > f->fsize) n = (size_t)(f->fsize - addr - f->base);
  return rd(f, addr, buf, n) ? 0 : n;
}

static int ilog2(u64 v) { int b = -1; while (v) { v >>= 1; ++b; } return b; }
#define WALK_BUDGET 2000000L
static int spend(h5r_file *f) { return (--f->budget < 0) ? fail(f, "h5r: index structure too large or cyclic") : 0; }

/* ---------------------------------------------------------------------------------------------------------------- */
/* IV.A. object headers: version 1 (IV.A.1.a) and version 2 (IV.A.1.b), continuation blocks (IV.A.2.q) */
static void msgs_free(MsgList *L)
{
  int i;
  for (i = 0; i < L->nblk; ++i) free(L->blk[i]);
  free(L->m);
  memset(L, 0, sizeof *L);
}
static int msgs_add(MsgList *L, int type, unsigned flags, size_t size, const unsigned char *data)
{
  if (L->n == L->cap) {
    L->cap = L->cap ? 2 * L->cap : 32;
    L->m = (Msg *)realloc(L->m, (size_t)L->cap * sizeof(Msg));
    if (!L->m) return -1;
  }
  L->m[L->n].type = type; L->m[L->n].flags = flags; L->m[L->n].size = size; L->m[L->n].data = data;
  ++L->n;
  return 0;
}

static int ohdr_block(h5r_file *f, MsgList *L, const unsigned char *p, size_t size, int v2, int corder,
                      u64 *cq_off, u64 *cq_len, int *ncq)
{
  size_t pos = 0;
  const size_t hdr = v2 ? (size_t)(4 + (corder ? 2 : 0)) : 8;
  while (pos + hdr <= size) {
    int type; size_t sz; unsigned flags;
    if (v2) { type = p[pos]; sz = (size_t)le_raw(p + pos + 1, 2); flags = p[pos + 3]; }
    else    { type = (int)le_raw(p + pos, 2); sz = (size_t)le_raw(p + pos + 2, 2); flags = p[pos + 4]; }
    if (pos + hdr + sz > size) {
      if (v2) break;                             /* a gap too small for a message ends a version 2 chunk */
      return fail(f, "h5r: object header message runs past its block");
    }
    if (type == 0x10) {                          /* continuation */
      if (sz < (size_t)(f->so + f->sl)) return fail(f, "h5r: short continuation message");
      if (*ncq >= MAX_OBJ_BLOCKS) return fail(f, "h5r: too many object header continuation blocks");
      cq_off[*ncq] = le(p + pos + hdr, f->so); cq_len[*ncq] = le_raw(p + pos + hdr + f->so, f->sl); ++*ncq;
    } else if (type != 0) {
      if (msgs_add(L, type, flags, sz, p + pos + hdr)) return fail(f, "h5r: out of memory");
    }
    pos += hdr + sz;
  }
  return 0;
}

static int ohdr_read(h5r_file *f, u64 addr, MsgList *L)
{
  unsigned char h[64];
  u64 cq_off[MAX_OBJ_BLOCKS], cq_len[MAX_OBJ_BLOCKS];
  int ncq = 0, icq = 0, v2, corder = 0;
  size_t got, pos, size;
  unsigned char *blk;
  memset(L, 0, sizeof *L);
  got = rd_upto(f, addr, h, sizeof h);
  if (got < 16) return fail(f, "h5r: object header at %llu is outside the file", addr);
  v2 = !memcmp(h, "OHDR", 4);
  if (v2) {
    const unsigned flags = h[5];
    if (h[4] != 2) return fail(f, "h5r: object header version %d", h[4]);
    pos = 6;
    if (flags & 0x20) pos += 16;                 /* access, modification, change, birth times */
    if (flags & 0x10) pos += 4;                  /* attribute storage phase change values */
    size = (size_t)le_raw(h + pos, 1 << (flags & 3));
    pos += (size_t)1 << (flags & 3);
    corder = (flags & 4) != 0;
  } else {
    if (h[0] != 1) return fail(f, "h5r: no object header at %llu (version byte %d)", addr, h[0]);
    size = (size_t)le_raw(h + 8, 4);
    pos = 16;                                    /* the 12-byte prefix is padded to the 8-byte message alignment */
  }
  if (size > ((size_t)1 << 28)) return fail(f, "h5r: object header at %llu claims %zu bytes", addr, size);
  blk = (unsigned char *)malloc(size ? size : 1);
  if (!blk) return fail(f, "h5r: out of memory");
  L->blk[L->nblk++] = blk;
  if (size && rd(f, addr + pos, blk, size)) { msgs_free(L); return -1; }
  if (ohdr_block(f, L, blk, size, v2, corder, cq_off, cq_len, &ncq)) { msgs_free(L); return -1; }
  while (icq < ncq) {
    const u64 off = cq_off[icq], len = cq_len[icq];
    ++icq;
    if (len > ((u64)1 << 28) || (v2 && len < 8)) { msgs_free(L); return fail(f, "h5r: bad continuation block length %llu", len); }
    if (L->nblk >= MAX_OBJ_BLOCKS) { msgs_free(L); return fail(f, "h5r: too many object header blocks"); }
    blk = (unsigned char *)malloc((size_t)len ? (size_t)len : 1);
    if (!blk) { msgs_free(L); return fail(f, "h5r: out of memory"); }
    L->blk[L->nblk++] = blk;
    if (rd(f, off, blk, (size_t)len)) { msgs_free(L); return -1; }
    if (v2) {
      if (memcmp(blk, "OCHK", 4)) { msgs_free(L); return fail(f, "h5r: continuation block at %llu lacks the OCHK signature", off); }
      if (ohdr_block(f, L, blk + 4, (size_t)len - 8, 1, corder, cq_off, cq_len, &ncq)) { msgs_free(L); return -1; }
    } else if (ohdr_block(f, L, blk, (size_t)len, 0, 0, cq_off, cq_len, &ncq)) { msgs_free(L); return -1; }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* III.E. global heap (variable-length data of attributes) */
static int gheap_get(h5r_file *f, u64 addr, unsigned index, const unsigned char **out, size_t *len)
{
  size_t pos;
  if (f->gcol_addr != addr || !f->gcol) {
    unsigned char h[32];
    u64 size;
    if (rd(f, addr, h, (size_t)(8 + f->sl))) return -1;
    if (memcmp(h, "GCOL", 4)) return fail(f, "h5r: no global heap collection at %llu", addr);
    size = le_raw(h + 8, f->sl);
    if (size < (u64)(8 + f->sl) || size > ((u64)1 << 30)) return fail(f, "h5r: bad global heap collection size");
    free(f->gcol);
    f->gcol = (unsigned char *)malloc((size_t)size);
    if (!f->gcol) return fail(f, "h5r: out of memory");
    f->gcol_addr = UNDEF;
    if (rd(f, addr, f->gcol, (size_t)size)) return -1;
    f->gcol_addr = addr; f->gcol_len = (size_t)size;
  }
  pos = (size_t)(8 + f->sl);
  while (pos + 8 + (size_t)f->sl <= f->gcol_len) {
    const unsigned idx = (unsigned)le_raw(f->gcol + pos, 2);
    const u64 osz = le_raw(f->gcol + pos + 8, f->sl);
    const size_t dpos = pos + 8 + (size_t)f->sl;
    if (idx == 0) break;                          /* free space: the rest of the collection */
    if (osz > f->gcol_len || dpos + osz > f->gcol_len) break;
    if (idx == index) { *out = f->gcol + dpos; *len = (size_t)osz; return 0; }
    pos = dpos + (size_t)((osz + 7) & ~7ULL);
  }
  return fail(f, "h5r: global heap object %u not found in the collection at %llu", index, addr);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* IV.A.2.b dataspace, IV.A.2.d datatype */
static int parse_dataspace(h5r_file *f, const unsigned char *p, size_t n, int *rank, long long *dims, long long *maxdims,
                           long long *nelem)
{
  size_t pos;
  int k, r, flags;
  if (n < 4) return fail(f, "h5r: short dataspace message");
  r = p[1]; flags = p[2];
  if (r > H5R_MAX_DIMS) return fail(f, "h5r: dataspace of rank %d", r);
  if (p[0] == 1) pos = 8;
  else if (p[0] == 2) { pos = 4; if (p[3] == 2) { *rank = 0; *nelem = 0; return 0; } }   /* null dataspace */
  else return fail(f, "h5r: dataspace message version %d", p[0]);
  if (pos + (size_t)r * f->sl * ((flags & 1) ? 2 : 1) > n) return fail(f, "h5r: short dataspace message");
  *rank = r; *nelem = 1;
  for (k = 0; k < r; ++k) {
    dims[k] = (long long)le_raw(p + pos + (size_t)k * f->sl, f->sl);
    if (dims[k] < 0 || (dims[k] > 0 && *nelem > (long long)(0x7fffffffffffffffLL / 16) / dims[k]))
      return fail(f, "h5r: dataspace extent overflows");
    *nelem *= dims[k];
    if (maxdims) maxdims[k] = dims[k];
  }
  if ((flags & 1) && maxdims)
    for (k = 0; k < r; ++k) {
      const u64 m = le(p + pos + (size_t)(r + k) * f->sl, f->sl);
      maxdims[k] = (m == UNDEF) ? -1 : (long long)m;
    }
  return 0;
}

enum { T_VLEN_REF = -9 };   /* internal: variable-length sequence of object references (DIMENSION_LIST) */
static int parse_datatype(const unsigned char *p, size_t n, h5r_type *t)
{
  int cls;
  if (n < 8) return -1;
  cls = p[0] & 0x0f;
  t->size = (int)le_raw(p + 4, 4);
  t->is_signed = 0; t->big_endian = 0; t->cls = H5R_OTHER;
  if (t->size <= 0) return -1;
  if (cls == 0) { t->cls = H5R_INT; t->big_endian = p[1] & 1; t->is_signed = (p[1] >> 3) & 1; }
  else if (cls == 1) { t->cls = H5R_FLOAT; t->big_endian = p[1] & 1; }
  else if (cls == 3) t->cls = H5R_STRING;
  else if (cls == 9) {
    if ((p[1] & 0x0f) == 1) t->cls = H5R_VLEN_STRING;
    else if (n >= 16 && (p[8] & 0x0f) == 7) t->cls = T_VLEN_REF;
  }
  return 0;
}

/* IV.A.2.m attribute message, versions 1 - 3.  Returns the bytes it occupies (heap scans step by it), -1 on error,
 * -2 if it cannot be sized (shared datatype / dataspace). */
static long parse_attribute(h5r_file *f, const unsigned char *p, size_t n, h5r_att *a)
{
  size_t pos, nsz, tsz, ssz, dbytes;
  int v, k;
  memset(a, 0, sizeof *a);
  if (n < 8) return fail(f, "h5r: short attribute message");
  v = p[0];
  if (v < 1 || v > 3) return fail(f, "h5r: attribute message version %d", v);
  if (v >= 2 && (p[1] & 3)) return -2;
  nsz = (size_t)le_raw(p + 2, 2); tsz = (size_t)le_raw(p + 4, 2); ssz = (size_t)le_raw(p + 6, 2);
  pos = (v == 3) ? 9 : 8;
#define STEP(x) ((v == 1) ? (((x) + 7) & ~(size_t)7) : (x))
  if (pos + STEP(nsz) + STEP(tsz) + STEP(ssz) > n || nsz == 0) return fail(f, "h5r: attribute message runs past its block");
  a->name = (char *)malloc(nsz + 1);
  if (!a->name) return fail(f, "h5r: out of memory");
  memcpy(a->name, p + pos, nsz); a->name[nsz] = 0;
  pos += STEP(nsz);
  if (parse_datatype(p + pos, tsz, &a->type)) { free(a->name); a->name = NULL; return fail(f, "h5r: bad datatype in an attribute"); }
  pos += STEP(tsz);
  {
    long long md[H5R_MAX_DIMS];
    if (parse_dataspace(f, p + pos, ssz, &a->rank, a->dims, md, &a->nelem)) { free(a->name); a->name = NULL; return -1; }
  }
  pos += STEP(ssz);
#undef STEP
  if (a->nelem > (1LL << 28) || (u64)a->nelem * (u64)a->type.size > ((u64)1 << 30)) { free(a->name); a->name = NULL; return fail(f, "h5r: oversized attribute"); }
  dbytes = (size_t)a->nelem * (size_t)a->type.size;
  if (pos + dbytes > n) { free(a->name); a->name = NULL; return fail(f, "h5r: attribute data runs past its block"); }
  if (a->type.cls == H5R_VLEN_STRING || a->type.cls == T_VLEN_REF) {
    /* IV.B: a variable-length element is (length, global heap collection address, object index) */
    const size_t es = (size_t)(4 + f->so + 4);
    size_t cap = 64, used = 0;
    if ((size_t)a->type.size != es) { free(a->name); a->name = NULL; return fail(f, "h5r: unexpected variable-length element size"); }
    a->data = (unsigned char *)malloc(cap);
    for (k = 0; k < a->nelem; ++k) {
      const unsigned char *e = p + pos + (size_t)k * es, *obj = NULL;
      const u64 len = le_raw(e, 4), haddr = le(e + 4, f->so);
      const unsigned idx = (unsigned)le_raw(e + 4 + f->so, 4);
      size_t olen = 0;
      if (len > 0 && haddr != UNDEF && haddr != 0) {
        if (gheap_get(f, haddr, idx, &obj, &olen)) { free(a->name); free(a->data); a->name = NULL; a->data = NULL; return -1; }
      }
      if (a->type.cls == T_VLEN_REF) {
        if (k < H5R_MAX_DIMS) { a->dimrefs[k] = (obj && olen >= 8) ? le(obj, 8) : UNDEF; a->ndimrefs = k + 1; }
      } else {
        const size_t sl = (len < olen) ? (size_t)len : olen;
        if (used + sl + 1 > cap) { while (used + sl + 1 > cap) cap *= 2; a->data = (unsigned char *)realloc(a->data, cap); }
        if (sl) memcpy(a->data + used, obj, sl);
        a->data[used + sl] = 0;
        used += sl + 1;
      }
    }
    a->data_len = used;
    if (a->type.cls == T_VLEN_REF) a->type.cls = H5R_OTHER;
  } else {
    a->data = (unsigned char *)malloc(dbytes ? dbytes : 1);
    if (!a->data) { free(a->name); a->name = NULL; return fail(f, "h5r: out of memory"); }
    memcpy(a->data, p + pos, dbytes);
    a->data_len = dbytes;
  }
  return (long)(pos + dbytes);
}

static void att_free(h5r_att *a) { free(a->name); free(a->data); }

typedef struct { int n, cap; h5r_att *a; } AttList;
static int atts_add(AttList *A, const h5r_att *a)
{
  if (A->n == A->cap) { A->cap = A->cap ? 2 * A->cap : 16; A->a = (h5r_att *)realloc(A->a, (size_t)A->cap * sizeof(h5r_att)); if (!A->a) return -1; }
  A->a[A->n++] = *a;
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* IV.A.2.g link message */
static long parse_link(h5r_file *f, const unsigned char *p, size_t n, Link *l)
{
  size_t pos = 2, nlen;
  unsigned flags;
  int type = 0, lw;
  l->name = NULL; l->addr = UNDEF; l->corder = -1;
  if (n < 4 || p[0] != 1) return fail(f, "h5r: link message version %d", n ? p[0] : -1);
  flags = p[1];
  if (flags & 8) type = p[pos++];
  if (flags & 4) { if (pos + 8 > n) return fail(f, "h5r: short link message"); l->corder = (long long)le_raw(p + pos, 8); pos += 8; }
  if (flags & 0x10) pos++;
  lw = 1 << (flags & 3);
  if (pos + (size_t)lw > n) return fail(f, "h5r: short link message");
  nlen = (size_t)le_raw(p + pos, lw); pos += (size_t)lw;
  if (pos + nlen > n || nlen == 0) return fail(f, "h5r: short link message");
  l->name = (char *)malloc(nlen + 1);
  if (!l->name) return fail(f, "h5r: out of memory");
  memcpy(l->name, p + pos, nlen); l->name[nlen] = 0;
  pos += nlen;
  if (type == 0) {
    if (pos + (size_t)f->so > n) { free(l->name); l->name = NULL; return fail(f, "h5r: short link message"); }
    l->addr = le(p + pos, f->so); pos += (size_t)f->so;
  } else {                                       /* soft, external, user-defined: (length, value); not followed */
    size_t vlen;
    if (pos + 2 > n) { free(l->name); l->name = NULL; return fail(f, "h5r: short link message"); }
    vlen = (size_t)le_raw(p + pos, 2); pos += 2 + vlen;
    if (pos > n) { free(l->name); l->name = NULL; return fail(f, "h5r: short link message"); }
  }
  return (long)pos;
}
static int links_add(LinkList *L, const Link *l)
{
  if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 32; L->l = (Link *)realloc(L->l, (size_t)L->cap * sizeof(Link)); if (!L->l) return -1; }
  L->l[L->n++] = *l;
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* III.G. fractal heap: every managed object handed to `cb` in address order.  The name index (a version 2 B-tree over
 * hashed names) is not needed for listing: the objects of a heap that was only ever appended to — all netCDF-4 does
 * when it defines a file — lie back to back from the start of each direct block, and a zero byte (no link or attribute
 * message starts with one) ends a block's used part.  The header's object count checks the walk. */
typedef long (*heap_cb)(h5r_file *f, const unsigned char *p, size_t n, void *ctx);
typedef struct {
  unsigned flags; u64 nman, nhuge, ntiny, start, maxdirect; unsigned width, maxbits, boff; u64 haddr; u64 found;
} FrHeap;

static int frheap_dblock(h5r_file *f, FrHeap *H, u64 addr, u64 size, heap_cb cb, void *ctx)
{
  const size_t hdr = (size_t)(5 + f->so + H->boff + ((H->flags & 2) ? 4 : 0));
  unsigned char *blk;
  size_t pos = hdr;
  if (spend(f)) return -1;
  if (size < hdr || size > ((u64)1 << 30)) return fail(f, "h5r: bad fractal heap direct block size %llu", size);
  blk = (unsigned char *)malloc((size_t)size);
  if (!blk) return fail(f, "h5r: out of memory");
  if (rd(f, addr, blk, (size_t)size)) { free(blk); return -1; }
  if (memcmp(blk, "FHDB", 4)) { free(blk); return fail(f, "h5r: no fractal heap direct block at %llu", addr); }
  while (pos < (size_t)size && blk[pos] != 0) {
    const long used = cb(f, blk + pos, (size_t)size - pos, ctx);
    if (used <= 0) { free(blk); return used == -2 ? fail(f, "h5r: a densely stored object uses a shared datatype (not supported)") : -1; }
    pos += (size_t)used;
    ++H->found;
  }
  free(blk);
  return 0;
}

static int frheap_iblock(h5r_file *f, FrHeap *H, u64 addr, unsigned nrows, heap_cb cb, void *ctx, int depth)
{
  const unsigned maxdrows = (unsigned)(ilog2(H->maxdirect) - ilog2(H->start)) + 2;
  const unsigned drows = nrows < maxdrows ? nrows : maxdrows;
  const size_t hdr = (size_t)(5 + f->so + H->boff);
  const size_t nent = (size_t)nrows * H->width;
  const size_t len = hdr + nent * (size_t)f->so + 4;
  unsigned char *blk;
  unsigned r, c;
  int rc = 0;
  if (depth > 8 || nrows > 64) return fail(f, "h5r: fractal heap nesting too deep");
  if (spend(f)) return -1;
  blk = (unsigned char *)malloc(len);
  if (!blk) return fail(f, "h5r: out of memory");
  if (rd(f, addr, blk, len - 4)) { free(blk); return -1; }
  if (memcmp(blk, "FHIB", 4)) { free(blk); return fail(f, "h5r: no fractal heap indirect block at %llu", addr); }
  for (r = 0; r < nrows && !rc; ++r) {
    const u64 bsize = (r < 2) ? H->start : H->start << (r - 1);
    for (c = 0; c < H->width && !rc; ++c) {
      const u64 child = le(blk + hdr + ((size_t)r * H->width + c) * (size_t)f->so, f->so);
      if (child == UNDEF) continue;
      if (r < drows) rc = frheap_dblock(f, H, child, bsize, cb, ctx);
      else {                                     /* a child indirect block spanning bsize bytes of heap space */
        const unsigned crow = (unsigned)(ilog2(bsize) - ilog2(H->start * H->width)) + 1;
        rc = frheap_iblock(f, H, child, crow, cb, ctx, depth + 1);
      }
    }
  }
  free(blk);
  return rc;
}

static int frheap_walk(h5r_file *f, u64 addr, heap_cb cb, void *ctx)
{
  unsigned char h[256];
  FrHeap H;
  size_t pos, got;
  unsigned iolen, currows;
  u64 root;
  memset(&H, 0, sizeof H);
  got = rd_upto(f, addr, h, sizeof h);
  if (got < (size_t)(22 + 12 * f->sl + 3 * f->so + 8)) return fail(f, "h5r: fractal heap header at %llu is outside the file", addr);
  if (memcmp(h, "FRHP", 4) || h[4] != 0) return fail(f, "h5r: no fractal heap header at %llu", addr);
  iolen = (unsigned)le_raw(h + 7, 2);
  H.flags = h[9];
  pos = 14;                                      /* signature, version, heap ID length, filter length, flags, max managed size */
  pos += (size_t)f->sl;                          /* next huge object ID */
  pos += (size_t)f->so;                          /* B-tree of huge objects */
  pos += (size_t)f->sl;                          /* free space in managed blocks */
  pos += (size_t)f->so;                          /* free space manager */
  pos += 3 * (size_t)f->sl;                      /* managed space, allocated managed space, allocation iterator offset */
  H.nman = le_raw(h + pos, f->sl); pos += (size_t)f->sl;
  pos += (size_t)f->sl;                          /* size of huge objects */
  H.nhuge = le_raw(h + pos, f->sl); pos += (size_t)f->sl;
  pos += (size_t)f->sl;                          /* size of tiny objects */
  H.ntiny = le_raw(h + pos, f->sl); pos += (size_t)f->sl;
  H.width = (unsigned)le_raw(h + pos, 2); pos += 2;
  H.start = le_raw(h + pos, f->sl); pos += (size_t)f->sl;
  H.maxdirect = le_raw(h + pos, f->sl); pos += (size_t)f->sl;
  H.maxbits = (unsigned)le_raw(h + pos, 2); pos += 2;
  pos += 2;                                      /* starting rows of the root indirect block */
  root = le(h + pos, f->so); pos += (size_t)f->so;
  currows = (unsigned)le_raw(h + pos, 2);
  H.boff = (H.maxbits + 7) / 8;
  H.haddr = addr;
  if (iolen) return fail(f, "h5r: filtered fractal heaps are not supported");
  if (H.nhuge || H.ntiny) return fail(f, "h5r: fractal heap with huge or tiny objects is not supported");
  if (!H.width || !H.start || H.start > H.maxdirect || (H.start & (H.start - 1)) || H.boff > 8)
    return fail(f, "h5r: inconsistent fractal heap header at %llu", addr);
  if (root != UNDEF) {
    if (currows == 0) { if (frheap_dblock(f, &H, root, H.start, cb, ctx)) return -1; }
    else if (frheap_iblock(f, &H, root, currows, cb, ctx, 0)) return -1;
  }
  if (H.found != H.nman)
    return fail(f, "h5r: fractal heap at %llu: walked %llu of %llu objects (the heap has holes; not supported)", addr, H.found, H.nman);
  return 0;
}

static long link_heap_cb(h5r_file *f, const unsigned char *p, size_t n, void *ctx)
{
  Link l;
  const long used = parse_link(f, p, n, &l);
  if (used <= 0) return -1;
  if (l.addr == UNDEF) { free(l.name); return used; }
  if (links_add((LinkList *)ctx, &l)) return fail(f, "h5r: out of memory");
  return used;
}
static long attr_heap_cb(h5r_file *f, const unsigned char *p, size_t n, void *ctx)
{
  h5r_att a;
  const long used = parse_attribute(f, p, n, &a);
  if (used <= 0) return used;
  if (atts_add((AttList *)ctx, &a)) return fail(f, "h5r: out of memory");
  return used;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* III.A. version 1 B-trees + III.C symbol table nodes + III.D local heaps: old-style groups */
static int group_btree(h5r_file *f, u64 addr, const unsigned char *heap, size_t heaplen, LinkList *out, int depth, int want_level)
{
  unsigned char h[8 + 2 * 8];
  unsigned char *node;
  unsigned level, nused, i;
  size_t hdr = (size_t)(8 + 2 * f->so), len;
  if (depth > 32) return fail(f, "h5r: group B-tree too deep");
  if (spend(f)) return -1;
  if (rd(f, addr, h, hdr)) return -1;
  if (memcmp(h, "TREE", 4) || h[4] != 0) return fail(f, "h5r: no group B-tree node at %llu", addr);
  level = h[5]; nused = (unsigned)le_raw(h + 6, 2);
  if (want_level >= 0 && (int)level != want_level) return fail(f, "h5r: group B-tree node at %llu has level %u, %d expected", addr, level, want_level);
  len = (size_t)nused * (size_t)(f->sl + f->so) + (size_t)f->sl;
  node = (unsigned char *)malloc(len);
  if (!node) return fail(f, "h5r: out of memory");
  if (rd(f, addr + hdr, node, len)) { free(node); return -1; }
  for (i = 0; i < nused; ++i) {
    const u64 child = le(node + (size_t)f->sl + (size_t)i * (size_t)(f->sl + f->so), f->so);
    if (level > 0) { if (group_btree(f, child, heap, heaplen, out, depth + 1, (int)level - 1)) { free(node); return -1; } }
    else {
      unsigned char sh[8];
      unsigned nsym, s;
      const size_t esz = (size_t)(2 * f->so + 24);
      unsigned char *ent;
      if (rd(f, child, sh, 8)) { free(node); return -1; }
      if (spend(f)) { free(node); return -1; }
      if (memcmp(sh, "SNOD", 4)) { free(node); return fail(f, "h5r: no symbol table node at %llu", child); }
      nsym = (unsigned)le_raw(sh + 6, 2);
      ent = (unsigned char *)malloc(nsym * esz + 1);
      if (!ent) { free(node); return fail(f, "h5r: out of memory"); }
      if (nsym && rd(f, child + 8, ent, nsym * esz)) { free(ent); free(node); return -1; }
      for (s = 0; s < nsym; ++s) {
        const u64 noff = le_raw(ent + s * esz, f->so);
        Link l;
        size_t nl;
        if (noff >= heaplen) { free(ent); free(node); return fail(f, "h5r: symbol name outside the local heap"); }
        nl = strnlen((const char *)heap + noff, heaplen - (size_t)noff);
        l.name = (char *)malloc(nl + 1);
        memcpy(l.name, heap + noff, nl); l.name[nl] = 0;
        l.addr = le(ent + s * esz + (size_t)f->so, f->so);
        l.corder = -1;
        if (links_add(out, &l)) { free(ent); free(node); return fail(f, "h5r: out of memory"); }
      }
      free(ent);
    }
  }
  free(node);
  return 0;
}

static int group_symtab(h5r_file *f, u64 btree, u64 heapaddr, LinkList *out)
{
  unsigned char h[8 + 3 * 8];
  unsigned char *heap;
  u64 dsize, daddr;
  int rc;
  if (rd(f, heapaddr, h, (size_t)(8 + 2 * f->sl + f->so))) return -1;
  if (memcmp(h, "HEAP", 4)) return fail(f, "h5r: no local heap at %llu", heapaddr);
  dsize = le_raw(h + 8, f->sl);
  daddr = le(h + 8 + 2 * f->sl, f->so);
  if (dsize > ((u64)1 << 30)) return fail(f, "h5r: bad local heap size");
  heap = (unsigned char *)malloc((size_t)dsize + 1);
  if (!heap) return fail(f, "h5r: out of memory");
  if (dsize && rd(f, daddr, heap, (size_t)dsize)) { free(heap); return -1; }
  rc = group_btree(f, btree, heap, (size_t)dsize, out, 0, -1);
  free(heap);
  return rc;
}

static int group_links(h5r_file *f, const MsgList *L, LinkList *out)
{
  int i;
  for (i = 0; i < L->n; ++i) {
    const Msg *m = &L->m[i];
    if (m->type == 0x11) {                       /* symbol table message */
      if (m->size < (size_t)(2 * f->so)) return fail(f, "h5r: short symbol table message");
      if (group_symtab(f, le(m->data, f->so), le(m->data + f->so, f->so), out)) return -1;
    } else if (m->type == 0x06) {                /* link message (compact storage) */
      Link l;
      if (parse_link(f, m->data, m->size, &l) <= 0) return -1;
      if (l.addr == UNDEF) free(l.name);
      else if (links_add(out, &l)) return fail(f, "h5r: out of memory");
    } else if (m->type == 0x02) {                /* link info: dense storage in a fractal heap */
      size_t pos = 2;
      u64 heap;
      if (m->size < 2) return fail(f, "h5r: short link info message");
      if (m->data[1] & 1) pos += 8;
      if (pos + (size_t)f->so > m->size) return fail(f, "h5r: short link info message");
      heap = le(m->data + pos, f->so);
      if (heap != UNDEF && frheap_walk(f, heap, link_heap_cb, out)) return -1;
    }
  }
  return 0;
}

static int object_atts(h5r_file *f, const MsgList *L, AttList *A)
{
  int i;
  for (i = 0; i < L->n; ++i) {
    const Msg *m = &L->m[i];
    if (m->type == 0x0C) {
      h5r_att a;
      long used;
      if (m->flags & 2) continue;                /* shared message */
      used = parse_attribute(f, m->data, m->size, &a);
      if (used == -2) continue;
      if (used <= 0) return -1;
      if (atts_add(A, &a)) return fail(f, "h5r: out of memory");
    } else if (m->type == 0x15) {                /* attribute info: dense storage */
      size_t pos = 2;
      u64 heap;
      if (m->size < 2) return fail(f, "h5r: short attribute info message");
      if (m->data[1] & 1) pos += 2;
      if (pos + (size_t)f->so > m->size) return fail(f, "h5r: short attribute info message");
      heap = le(m->data + pos, f->so);
      if (heap != UNDEF && frheap_walk(f, heap, attr_heap_cb, A)) return -1;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* datasets: IV.A.2.i layout, IV.A.2.l filter pipeline, IV.A.2.f / e fill value */
static void unsupported(h5r_dset *d, const char *fmt, ...)
{
  va_list ap;
  if (!d->supported) return;
  d->supported = 0;
  va_start(ap, fmt);
  vsnprintf(d->why, sizeof d->why, fmt, ap);
  va_end(ap);
}

static int parse_dataset(h5r_file *f, const MsgList *L, h5r_dset *d)
{
  DsPriv *P = (DsPriv *)calloc(1, sizeof(DsPriv));
  int i, have_space = 0, have_type = 0, have_layout = 0;
  if (!P) return fail(f, "h5r: out of memory");
  d->priv = P; d->supported = 1; P->cache_chunk = -1;
  for (i = 0; i < L->n; ++i) {
    const Msg *m = &L->m[i];
    const unsigned char *p = m->data;
    if (m->type == 0x01) {
      long long ne;
      if (m->flags & 2) { unsupported(d, "shared dataspace"); continue; }
      if (parse_dataspace(f, p, m->size, &d->rank, d->dims, d->maxdims, &ne)) return -1;
      have_space = 1;
    } else if (m->type == 0x03) {
      if (m->flags & 2) { unsupported(d, "committed (shared) datatype"); continue; }
      if (parse_datatype(p, m->size, &d->type)) return fail(f, "h5r: bad datatype message");
      if (d->type.cls == T_VLEN_REF) d->type.cls = H5R_OTHER;
      have_type = 1;
    } else if (m->type == 0x08) {
      if (m->size < 2) return fail(f, "h5r: short layout message");
      have_layout = 1;
      if (p[0] == 3) {
        P->layout = p[1];
        if (p[1] == 0) {
          size_t n;
          if (m->size < 4) return fail(f, "h5r: short layout message");
          n = (size_t)le_raw(p + 2, 2);
          if (4 + n > m->size) return fail(f, "h5r: compact data runs past the layout message");
          P->compact = (unsigned char *)malloc(n ? n : 1); memcpy(P->compact, p + 4, n); P->compact_len = n;
        } else if (p[1] == 1) {
          if (m->size < (size_t)(2 + f->so + f->sl)) return fail(f, "h5r: short layout message");
          P->addr = le(p + 2, f->so); P->size = le_raw(p + 2 + f->so, f->sl);
        } else if (p[1] == 2) {
          int k;
          if (m->size < 3 || p[2] < 1 || p[2] > H5R_MAX_DIMS + 1 || m->size < (size_t)(3 + f->so + 4 * p[2])) return fail(f, "h5r: bad chunked layout message");
          P->cnd = p[2];
          P->btree = le(p + 3, f->so);
          for (k = 0; k < P->cnd; ++k) P->cdims[k] = le_raw(p + 3 + f->so + 4 * k, 4);
        } else unsupported(d, "layout class %d", p[1]);
      } else if (p[0] == 1 || p[0] == 2) {       /* HDF5 1.6 and earlier */
        const int nd = p[1], cls = p[2];
        size_t pos = 8;
        int k;
        if (nd > H5R_MAX_DIMS + 1) return fail(f, "h5r: bad layout message");
        P->layout = cls;
        if (cls != 0) { if (pos + (size_t)f->so > m->size) return fail(f, "h5r: short layout message"); P->addr = P->btree = le(p + pos, f->so); pos += (size_t)f->so; }
        if (pos + 4 * (size_t)nd > m->size) return fail(f, "h5r: short layout message");
        for (k = 0; k < nd; ++k) P->cdims[k] = le_raw(p + pos + 4 * k, 4);
        pos += 4 * (size_t)nd;
        if (cls == 2) P->cnd = nd;
        else if (cls == 0) {
          size_t n;
          if (pos + 4 > m->size) return fail(f, "h5r: short layout message");
          n = (size_t)le_raw(p + pos, 4); pos += 4;
          if (pos + n > m->size) return fail(f, "h5r: compact data runs past the layout message");
          P->compact = (unsigned char *)malloc(n ? n : 1); memcpy(P->compact, p + pos, n); P->compact_len = n;
        }
      } else unsupported(d, "data layout message version %d (HDF5 1.10+ chunk indices)", p[0]);
    } else if (m->type == 0x0B) {
      size_t pos;
      int k, nf, c;
      if (m->size < 2) return fail(f, "h5r: short filter pipeline message");
      nf = p[1];
      if (nf > MAX_FILTERS) { unsupported(d, "%d filters", nf); continue; }
      pos = (p[0] == 1) ? 8 : 2;
      for (k = 0; k < nf; ++k) {
        size_t nlen = 0;
        int id, ncd;
        if (pos + 2 > m->size) return fail(f, "h5r: short filter pipeline message");
        id = (int)le_raw(p + pos, 2); pos += 2;
        if (p[0] == 1 || id >= 256) { if (pos + 2 > m->size) return fail(f, "h5r: short filter pipeline message"); nlen = (size_t)le_raw(p + pos, 2); pos += 2; }
        if (pos + 4 > m->size) return fail(f, "h5r: short filter pipeline message");
        pos += 2;                                /* flags */
        ncd = (int)le_raw(p + pos, 2); pos += 2;
        if (p[0] == 1) nlen = (nlen + 7) & ~(size_t)7;
        pos += nlen;
        if (pos + 4 * (size_t)ncd > m->size) return fail(f, "h5r: short filter pipeline message");
        P->filt[k].id = id; P->filt[k].ncd = ncd < 8 ? ncd : 8;
        for (c = 0; c < P->filt[k].ncd; ++c) P->filt[k].cd[c] = (unsigned)le_raw(p + pos + 4 * c, 4);
        pos += 4 * (size_t)ncd;
        if (p[0] == 1 && (ncd & 1)) pos += 4;
        if (id != 1 && id != 2 && id != 3) unsupported(d, "filter %d (only deflate, shuffle, fletcher32)", id);
      }
      P->nfilters = nf;
    } else if (m->type == 0x05) {                /* fill value */
      if (m->size >= 2 && p[0] == 3) {
        if ((p[1] & 0x20) && m->size >= 6) {
          const size_t n = (size_t)le_raw(p + 2, 4);
          if (6 + n <= m->size && n > 0 && n <= 64) { P->fill = (unsigned char *)malloc(n); memcpy(P->fill, p + 6, n); P->fill_len = (int)n; }
        }
      } else if (m->size >= 4 && (p[0] == 1 || p[0] == 2)) {
        if ((p[0] == 1 || p[3]) && m->size >= 8) {
          const size_t n = (size_t)le_raw(p + 4, 4);
          if (8 + n <= m->size && n > 0 && n <= 64) { P->fill = (unsigned char *)malloc(n); memcpy(P->fill, p + 8, n); P->fill_len = (int)n; }
        }
      }
    }
  }
  if (!have_space || !have_type || !have_layout) unsupported(d, "no dataspace, datatype or layout message");
  if (d->supported && d->type.cls == H5R_OTHER) unsupported(d, "datatype class outside integer / float / string");
  if (d->supported && d->type.cls == H5R_VLEN_STRING) unsupported(d, "variable-length string data");
  if (d->supported && P->layout == 2) {
    if (d->rank == 0) unsupported(d, "chunked scalar");
    else if (P->cnd != d->rank + 1) unsupported(d, "chunk rank %d for a rank-%d dataset", P->cnd - 1, d->rank);
    else {
      int k;
      u64 bytes = (u64)d->type.size;
      for (k = 0; k < d->rank; ++k) { if (P->cdims[k] == 0 || bytes > ((u64)1 << 32) / P->cdims[k]) { unsupported(d, "chunk too large"); break; } bytes *= P->cdims[k]; }
    }
  }
  if (P->fill && P->fill_len != d->type.size) { free(P->fill); P->fill = NULL; P->fill_len = 0; }
  {
    AttList A = {0, 0, NULL};
    if (object_atts(f, L, &A)) { int k; for (k = 0; k < A.n; ++k) att_free(&A.a[k]); free(A.a); return -1; }
    d->natts = A.n; d->atts = A.a;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------- */
static int cmp_links(const void *a, const void *b)
{
  const Link *x = (const Link *)a, *y = (const Link *)b;
  if (x->corder != y->corder) return x->corder < y->corder ? -1 : 1;
  return strcmp(x->name, y->name);
}

h5r_file *h5r_open(const char *path, char *err, size_t errlen)
{
  h5r_file *f = (h5r_file *)calloc(1, sizeof(h5r_file));
  struct stat st;
  unsigned char sb[128];
  u64 off, root = UNDEF;
  int found = 0, i;
  MsgList L;
  LinkList links = {0, 0, NULL};
  if (!f) { if (err) snprintf(err, errlen, "h5r_open: out of memory"); return NULL; }
  f->gcol_addr = UNDEF; f->budget = WALK_BUDGET;
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) { if (err) snprintf(err, errlen, "h5r_open: cannot open %s: %s", path, strerror(errno)); free(f); return NULL; }
  fstat(f->fd, &st);
  f->fsize = (u64)st.st_size;
  /* II.A: the superblock sits at 0, 512, 1024, 2048, ... (a user block may precede it) */
  for (off = 0; off + 16 < f->fsize; off = off ? off * 2 : 512) {
    if (pread(f->fd, sb, 8, (off_t)off) == 8 && !memcmp(sb, "\211HDF\r\n\032\n", 8)) { found = 1; break; }
    if (off > ((u64)1 << 32)) break;
  }
  if (!found) { snprintf(f->err, sizeof f->err, "h5r_open: %s has no HDF5 superblock", path); goto bad; }
  {
    ssize_t got = pread(f->fd, sb, sizeof sb, (off_t)off);
    if (got < 48) { snprintf(f->err, sizeof f->err, "h5r_open: %s: truncated superblock", path); goto bad; }
    memset(sb + got, 0, sizeof sb - (size_t)got);
  }
  if (sb[8] == 0 || sb[8] == 1) {
    size_t pos = (sb[8] == 1) ? 28 : 24;
    f->so = sb[13]; f->sl = sb[14];
    if ((f->so != 4 && f->so != 8) || (f->sl != 4 && f->sl != 8)) { snprintf(f->err, sizeof f->err, "h5r_open: %s: offsets of %d / lengths of %d bytes", path, f->so, f->sl); goto bad; }
    f->base = le(sb + pos, f->so); pos += (size_t)f->so;
    pos += (size_t)f->so;                        /* free-space info */
    f->eof = le(sb + pos, f->so); pos += (size_t)f->so;
    pos += (size_t)f->so;                        /* driver information block */
    root = le(sb + pos + (size_t)f->so, f->so);  /* root group symbol table entry: name offset, OBJECT HEADER ADDRESS */
  } else if (sb[8] == 2 || sb[8] == 3) {
    f->so = sb[9]; f->sl = sb[10];
    if ((f->so != 4 && f->so != 8) || (f->sl != 4 && f->sl != 8)) { snprintf(f->err, sizeof f->err, "h5r_open: %s: offsets of %d / lengths of %d bytes", path, f->so, f->sl); goto bad; }
    f->base = le(sb + 12, f->so);
    f->eof = le(sb + 12 + 2 * f->so, f->so);
    root = le(sb + 12 + 3 * f->so, f->so);
  } else { snprintf(f->err, sizeof f->err, "h5r_open: %s: superblock version %d", path, sb[8]); goto bad; }
  if (f->base == UNDEF) f->base = 0;
  if (f->base == 0 && off != 0) f->base = off;   /* written behind a user block without recording it */
  if (ohdr_read(f, root, &L)) goto bad;
  {
    AttList A = {0, 0, NULL};
    int rc = group_links(f, &L, &links);
    if (!rc) rc = object_atts(f, &L, &A);
    f->ngatts = A.n; f->gatts = A.a;
    msgs_free(&L);
    if (rc) goto bad;
  }
  if (links.n) qsort(links.l, (size_t)links.n, sizeof(Link), cmp_links);
  f->dsets = (h5r_dset *)calloc(links.n ? (size_t)links.n : 1, sizeof(h5r_dset));
  for (i = 0; i < links.n; ++i) {
    int k, is_dataset = 0;
    if (ohdr_read(f, links.l[i].addr, &L)) goto bad;
    for (k = 0; k < L.n; ++k) if (L.m[k].type == 0x08) is_dataset = 1;
    if (is_dataset) {
      h5r_dset *d = &f->dsets[f->ndsets];
      d->name = links.l[i].name; links.l[i].name = NULL;
      d->addr = links.l[i].addr; d->order = links.l[i].corder;
      ++f->ndsets;
      if (parse_dataset(f, &L, d)) { msgs_free(&L); goto bad; }
    }
    msgs_free(&L);
  }
  for (i = 0; i < links.n; ++i) free(links.l[i].name);
  free(links.l);
  return f;
bad:
  if (err) snprintf(err, errlen, "%s", f->err);
  for (i = 0; i < links.n; ++i) free(links.l[i].name);
  free(links.l);
  h5r_close(f);
  return NULL;
}

void h5r_close(h5r_file *f)
{
  int i, k;
  if (!f) return;
  for (i = 0; i < f->ndsets; ++i) {
    h5r_dset *d = &f->dsets[i];
    DsPriv *P = (DsPriv *)d->priv;
    free(d->name);
    for (k = 0; k < d->natts; ++k) att_free(&d->atts[k]);
    free(d->atts);
    if (P) { free(P->compact); free(P->fill); free(P->chunks); free(P->cache); free(P); }
  }
  free(f->dsets);
  for (k = 0; k < f->ngatts; ++k) att_free(&f->gatts[k]);
  free(f->gatts);
  free(f->gcol);
  if (f->fd >= 0) close(f->fd);
  free(f);
}

int h5r_ndsets(const h5r_file *f) { return f->ndsets; }
const h5r_dset *h5r_dset_at(const h5r_file *f, int i) { return (i < 0 || i >= f->ndsets) ? NULL : &f->dsets[i]; }
int h5r_ngatts(const h5r_file *f) { return f->ngatts; }
const h5r_att *h5r_gatt_at(const h5r_file *f, int i) { return (i < 0 || i >= f->ngatts) ? NULL : &f->gatts[i]; }

/* ---------------------------------------------------------------------------------------------------------------- */
/* reading */
static int g_rank_for_cmp;
static int cmp_chunks(const void *a, const void *b)
{
  const Chunk *x = (const Chunk *)a, *y = (const Chunk *)b;
  int k;
  for (k = 0; k < g_rank_for_cmp; ++k) if (x->off[k] != y->off[k]) return x->off[k] < y->off[k] ? -1 : 1;
  return 0;
}

/* III.A.1 B-tree nodes of type 1: keys are (chunk size, filter mask, offset of the chunk in every dimension + 0) */
static int chunk_btree(h5r_file *f, DsPriv *P, int rank, u64 addr, int depth, long *cap, int want_level)
{
  unsigned char h[8 + 2 * 8];
  unsigned char *node;
  unsigned level, nused, i;
  const size_t hdr = (size_t)(8 + 2 * f->so), ksz = (size_t)(8 + 8 * (rank + 1));
  size_t len;
  int k;
  if (depth > 32) return fail(f, "h5r: chunk B-tree too deep");
  if (spend(f)) return -1;
  if (rd(f, addr, h, hdr)) return -1;
  if (memcmp(h, "TREE", 4) || h[4] != 1) return fail(f, "h5r: no chunk B-tree node at %llu", addr);
  level = h[5]; nused = (unsigned)le_raw(h + 6, 2);
  if (want_level >= 0 && (int)level != want_level) return fail(f, "h5r: chunk B-tree node at %llu has level %u, %d expected", addr, level, want_level);
  len = (size_t)nused * (ksz + (size_t)f->so) + ksz;
  node = (unsigned char *)malloc(len);
  if (!node) return fail(f, "h5r: out of memory");
  if (rd(f, addr + hdr, node, len)) { free(node); return -1; }
  for (i = 0; i < nused; ++i) {
    const unsigned char *key = node + (size_t)i * (ksz + (size_t)f->so);
    const u64 child = le(key + ksz, f->so);
    if (level > 0) { if (chunk_btree(f, P, rank, child, depth + 1, cap, (int)level - 1)) { free(node); return -1; } }
    else {
      Chunk *c;
      if (P->nchunks == *cap) { *cap = *cap ? 2 * *cap : 256; P->chunks = (Chunk *)realloc(P->chunks, (size_t)*cap * sizeof(Chunk)); if (!P->chunks) { free(node); return fail(f, "h5r: out of memory"); } }
      c = &P->chunks[P->nchunks++];
      c->size = (unsigned)le_raw(key, 4); c->mask = (unsigned)le_raw(key + 4, 4);
      for (k = 0; k < rank; ++k) c->off[k] = le_raw(key + 8 + 8 * k, 8);
      for (k = rank; k < H5R_MAX_DIMS; ++k) c->off[k] = 0;
      c->addr = child;
    }
  }
  free(node);
  return 0;
}

static int decode_chunk(h5r_file *f, const h5r_dset *d, DsPriv *P, const Chunk *c, unsigned char *out, size_t nbytes)
{
  unsigned char *cur = (unsigned char *)malloc(c->size ? c->size : 1), *tmp;
  size_t curlen = c->size;
  int k;
  if (!cur) return fail(f, "h5r: out of memory");
  if (rd(f, c->addr, cur, c->size)) { free(cur); return -1; }
  for (k = P->nfilters - 1; k >= 0; --k) {
    if (c->mask & (1u << k)) continue;           /* this filter was skipped for this chunk */
    if (P->filt[k].id == 3) { if (curlen < 4) { free(cur); return fail(f, "h5r: %s: chunk shorter than its checksum", d->name); } curlen -= 4; }
    else if (P->filt[k].id == 1) {
      uLongf dlen = (uLongf)(nbytes + 4);
      int zr;
      tmp = (unsigned char *)malloc(dlen);
      if (!tmp) { free(cur); return fail(f, "h5r: out of memory"); }
      zr = uncompress(tmp, &dlen, cur, (uLong)curlen);
      if (zr != Z_OK) { free(tmp); free(cur); return fail(f, "h5r: %s: inflate failed (%d)", d->name, zr); }
      free(cur); cur = tmp; curlen = (size_t)dlen;
    } else if (P->filt[k].id == 2) {
      const size_t es = P->filt[k].ncd > 0 ? P->filt[k].cd[0] : (size_t)d->type.size;
      if (es > 1 && curlen >= es) {
        const size_t ne = curlen / es, tail = curlen - ne * es;
        size_t e, b;
        tmp = (unsigned char *)malloc(curlen);
        if (!tmp) { free(cur); return fail(f, "h5r: out of memory"); }
        for (b = 0; b < es; ++b) for (e = 0; e < ne; ++e) tmp[e * es + b] = cur[b * ne + e];
        memcpy(tmp + ne * es, cur + ne * es, tail);
        free(cur); cur = tmp;
      }
    }
  }
  if (curlen < nbytes) { free(cur); return fail(f, "h5r: %s: chunk decodes to %zu bytes, %zu expected", d->name, curlen, nbytes); }
  memcpy(out, cur, nbytes);
  free(cur);
  return 0;
}

static void fill_elems(unsigned char *out, size_t n, const DsPriv *P, int es)
{
  size_t i;
  if (!P->fill) { memset(out, 0, n * (size_t)es); return; }
  for (i = 0; i < n; ++i) memcpy(out + i * (size_t)es, P->fill, (size_t)es);
}

int h5r_read(h5r_file *f, int idx, const size_t *start, const size_t *count, void *outv)
{
  h5r_dset *d;
  DsPriv *P;
  unsigned char *out = (unsigned char *)outv;
  size_t st[H5R_MAX_DIMS + 1], ct[H5R_MAX_DIMS + 1], dl[H5R_MAX_DIMS + 1], total = 1;
  int rank, k, es;
  if (idx < 0 || idx >= f->ndsets) return fail(f, "h5r: bad dataset index %d", idx);
  d = &f->dsets[idx]; P = (DsPriv *)d->priv;
  if (!d->supported) return fail(f, "h5r: dataset %s cannot be read: %s", d->name, d->why);
  es = d->type.size;
  rank = d->rank;
  for (k = 0; k < rank; ++k) {
    dl[k] = (size_t)d->dims[k];
    st[k] = start ? start[k] : 0; ct[k] = count ? count[k] : dl[k];
    if (st[k] + ct[k] > dl[k]) return fail(f, "h5r: %s: start+count exceeds dimension %d", d->name, k);
    total *= ct[k];
  }
  if (rank == 0) { rank = 1; dl[0] = 1; st[0] = 0; ct[0] = 1; }
  if (total == 0) return 0;

  if (P->layout == 0 || P->layout == 1) {
    /* rows of the innermost dimension, merged over the trailing dimensions the slab covers in full */
    size_t idxv[H5R_MAX_DIMS + 1] = {0}, run = ct[rank - 1], opos = 0;
    int top = rank - 1;                          /* dimensions > top are covered in full and folded into the run */
    u64 need = (u64)es;
    for (k = 0; k < rank; ++k) need *= dl[k];
    while (top > 0 && ct[top] == dl[top]) { --top; run *= ct[top]; }
    if (P->layout == 0 && P->compact_len < need) return fail(f, "h5r: %s: compact data shorter than the dataspace", d->name);
    for (;;) {
      u64 eoff = 0;
      for (k = 0; k < rank; ++k) eoff = eoff * dl[k] + (k <= top ? st[k] + (k < top ? idxv[k] : 0) : 0);
      if (P->layout == 0) memcpy(out + opos, P->compact + eoff * (u64)es, run * (size_t)es);
      else if (P->addr == UNDEF) fill_elems(out + opos, run, P, es);
      else if (rd(f, P->addr + eoff * (u64)es, out + opos, run * (size_t)es)) return -1;
      opos += run * (size_t)es;
      for (k = top - 1; k >= 0; --k) { if (++idxv[k] < ct[k]) break; idxv[k] = 0; }
      if (k < 0) break;
    }
  } else {
    size_t cshape[H5R_MAX_DIMS], c0[H5R_MAX_DIMS], c1[H5R_MAX_DIMS], ci[H5R_MAX_DIMS], cbytes = (size_t)es;
    size_t ostride[H5R_MAX_DIMS];
    if (!P->indexed) {
      long cap = 0;
      f->budget = WALK_BUDGET;
      if (P->btree != UNDEF && chunk_btree(f, P, d->rank, P->btree, 0, &cap, -1)) return -1;
      g_rank_for_cmp = d->rank;
      if (P->nchunks) qsort(P->chunks, (size_t)P->nchunks, sizeof(Chunk), cmp_chunks);
      P->indexed = 1;
    }
    for (k = 0; k < rank; ++k) {
      cshape[k] = (size_t)P->cdims[k]; cbytes *= cshape[k];
      c0[k] = st[k] / cshape[k]; c1[k] = (st[k] + ct[k] - 1) / cshape[k]; ci[k] = c0[k];
    }
    ostride[rank - 1] = 1;
    for (k = rank - 2; k >= 0; --k) ostride[k] = ostride[k + 1] * ct[k + 1];
    if (!P->cache) { P->cache = (unsigned char *)malloc(cbytes); if (!P->cache) return fail(f, "h5r: out of memory"); }
    for (;;) {
      Chunk key, *hit;
      size_t lo[H5R_MAX_DIMS], n[H5R_MAX_DIMS], ii[H5R_MAX_DIMS] = {0};
      memset(&key, 0, sizeof key);
      for (k = 0; k < rank; ++k) {
        const size_t cb = ci[k] * cshape[k], ce = cb + cshape[k];
        const size_t a = st[k] > cb ? st[k] : cb, b = (st[k] + ct[k]) < ce ? (st[k] + ct[k]) : ce;
        key.off[k] = cb; lo[k] = a; n[k] = b - a;
      }
      g_rank_for_cmp = rank;
      hit = P->nchunks ? (Chunk *)bsearch(&key, P->chunks, (size_t)P->nchunks, sizeof(Chunk), cmp_chunks) : NULL;
      if (hit && (long)(hit - P->chunks) != P->cache_chunk) {
        P->cache_chunk = -1;
        if (decode_chunk(f, d, P, hit, P->cache, cbytes)) return -1;
        P->cache_chunk = (long)(hit - P->chunks);
      }
      for (;;) {                                 /* rows of the intersection */
        size_t opos = 0, cpos = 0;
        for (k = 0; k < rank; ++k) {
          const size_t g = lo[k] + (k < rank - 1 ? ii[k] : 0);
          opos += (g - st[k]) * ostride[k];
          cpos = cpos * cshape[k] + (g - key.off[k]);
        }
        if (hit) memcpy(out + opos * (size_t)es, P->cache + cpos * (size_t)es, n[rank - 1] * (size_t)es);
        else fill_elems(out + opos * (size_t)es, n[rank - 1], P, es);
        for (k = rank - 2; k >= 0; --k) { if (++ii[k] < n[k]) break; ii[k] = 0; }
        if (k < 0) break;
      }
      for (k = rank - 1; k >= 0; --k) { if (++ci[k] <= c1[k]) break; ci[k] = c0[k]; }
      if (k < 0) break;
    }
  }
  if (d->type.big_endian && es > 1 && (d->type.cls == H5R_INT || d->type.cls == H5R_FLOAT)) {
    size_t i;
    int b;
    for (i = 0; i < total; ++i) {
      unsigned char *e = out + i * (size_t)es;
      for (b = 0; b < es / 2; ++b) { const unsigned char t = e[b]; e[b] = e[es - 1 - b]; e[es - 1 - b] = t; }
    }
  }
  return 0;
}
