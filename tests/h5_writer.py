"""A small netCDF-4 (HDF5) file WRITER for the tests of csrc/h5r.c — test infrastructure, not product code.

This image has no libhdf5 / libnetcdf / h5py, and the reference tree holds no HDF5 sample file, so the reader cannot be
pinned against bytes the real library produced (DESIGN.md §2 says so).  This writer is the independent second
implementation of the same published document (HDF5 File Format Specification 3.0) and of netcdf-c's conventions
(libhdf5/nc4hdf.c: dimension scales, DIMENSION_LIST / REFERENCE_LIST, _Netcdf4Dimid, _NCProperties, creation-order
tracking on groups and attributes).  It lays files out the way libnetcdf does with its two library-version bounds:

  style "v18"      (netcdf-c >= 4.7: H5F_LIBVER_V18 .. LATEST): superblock 2, object headers 2 with times and attribute
                   creation order, dataspace 2, fill value 3, attribute 3, filter pipeline 2, link messages or — above
                   eight links / attributes — dense storage in a fractal heap, layout 3 with version 1 chunk B-trees
  style "earliest" (netcdf-c <= 4.6: H5F_LIBVER_EARLIEST .. LATEST): superblock 0, the same version 2 object headers
                   (attribute creation order forces them), dataspace 1, fill value 2, attribute 1, filter pipeline 1
  style "plain"    a plain HDF5 1.6-style file (no netCDF conventions): symbol-table root group (B-tree 1, local heap,
                   SNOD), version 1 object headers; variables come out with phony dimensions

Metadata checksums are Jenkins lookup3 as the specification names it (unverified here: no reader in this image checks
them).  The writer runs twice so that object references (which need addresses) do not change the layout.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
M32 = 0xFFFFFFFF
NOT_A_VAR = "This is a netCDF dimension but not a netCDF variable."


def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & M32


def lookup3(key, initval=0):
    a = b = c = (0xDEADBEEF + len(key) + initval) & M32
    k = bytes(key)
    n = len(k)
    p = 0
    while n > 12:
        a = (a + int.from_bytes(k[p:p + 4], "little")) & M32
        b = (b + int.from_bytes(k[p + 4:p + 8], "little")) & M32
        c = (c + int.from_bytes(k[p + 8:p + 12], "little")) & M32
        a = (a - c) & M32; a ^= _rot(c, 4); c = (c + b) & M32
        b = (b - a) & M32; b ^= _rot(a, 6); a = (a + c) & M32
        c = (c - b) & M32; c ^= _rot(b, 8); b = (b + a) & M32
        a = (a - c) & M32; a ^= _rot(c, 16); c = (c + b) & M32
        b = (b - a) & M32; b ^= _rot(a, 19); a = (a + c) & M32
        c = (c - b) & M32; c ^= _rot(b, 4); b = (b + a) & M32
        p += 12
        n -= 12
    if n == 0:
        return c
    tail = k[p:] + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & M32
    b = (b + int.from_bytes(tail[4:8], "little")) & M32
    c = (c + int.from_bytes(tail[8:12], "little")) & M32
    c ^= b; c = (c - _rot(b, 14)) & M32
    a ^= c; a = (a - _rot(c, 11)) & M32
    b ^= a; b = (b - _rot(a, 25)) & M32
    c ^= b; c = (c - _rot(b, 16)) & M32
    a ^= c; a = (a - _rot(c, 4)) & M32
    b ^= a; b = (b - _rot(a, 14)) & M32
    c ^= b; c = (c - _rot(b, 24)) & M32
    return c


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class Var:
    """One netCDF variable (or, with data=None and is_dim_only, a dimension without a coordinate variable)."""

    def __init__(self, name, dims, data, atts=None, chunks=None, deflate=0, shuffle=False, fletcher=False, fill=None,
                 big_endian=False, layout=None):
        self.name, self.dims, self.atts = name, tuple(dims), dict(atts or {})
        self.data = data
        self.chunks, self.deflate, self.shuffle, self.fletcher, self.fill = chunks, deflate, shuffle, fletcher, fill
        self.big_endian = big_endian
        self.layout = layout            # None: contiguous unless chunks are given; "compact"


class H5Writer:
    def __init__(self, style="v18", leaf=8, split_header=True, user_block=0):
        assert style in ("v18", "earliest", "plain")
        self.style, self.leaf, self.split_header, self.user_block = style, leaf, split_header, user_block
        self.new = style != "plain"
        self.latest = style == "v18"

    # ------------------------------------------------------------------ allocation
    def alloc(self, data, align=8):
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ------------------------------------------------------------------ messages
    def dataspace(self, shape, maxshape=None, null=False):
        rank = 0 if shape is None else len(shape)
        flags = 1 if maxshape is not None else 0
        body = b"".join(struct.pack("<Q", d) for d in (shape or ()))
        if maxshape is not None:
            body += b"".join(struct.pack("<Q", UNDEF if d is None else d) for d in maxshape)
        if self.latest:
            kind = 2 if null else (1 if rank else 0)
            return struct.pack("<BBBB", 2, rank, flags, kind) + body
        return struct.pack("<BBBBI", 1, rank, flags, 0, 0) + body

    @staticmethod
    def datatype(dt, strlen=None):
        if strlen is not None:            # fixed-length string, null terminated, ASCII
            return struct.pack("<BBBBI", 0x13, 0, 0, 0, strlen)
        dt = np.dtype(dt)
        be = 1 if dt.byteorder == ">" else 0
        if dt.kind in "iu":
            bits0 = be | (8 if dt.kind == "i" else 0)
            return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
        if dt.kind == "f":
            if dt.itemsize == 4:
                sign, eloc, esz, mloc, msz, bias = 31, 23, 8, 0, 23, 127
            else:
                sign, eloc, esz, mloc, msz, bias = 63, 52, 11, 0, 52, 1023
            return (struct.pack("<BBBBI", 0x11, be | 0x20, sign, 0, dt.itemsize) +
                    struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esz, mloc, msz, bias))
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
        raise ValueError(dt)

    @staticmethod
    def dtype_ref():
        return struct.pack("<BBBBI", 0x17, 0, 0, 0, 8)

    def dtype_vlen_ref(self):
        return struct.pack("<BBBBI", 0x19, 0, 0, 0, 16) + self.dtype_ref()

    def dtype_reflist(self):
        """REFERENCE_LIST's compound {dataset: object reference, dimension: uint32} (H5DS)."""
        ref, u32 = self.dtype_ref(), self.datatype("<u4")
        if self.latest:                    # compound version 3: names unpadded, member offset in as few bytes as the size needs
            body = b"dataset\0" + struct.pack("<B", 0) + ref + b"dimension\0" + struct.pack("<B", 8) + u32
            return struct.pack("<BBBBI", 0x36, 2, 0, 0, 12) + body
        def member(name, off, t):
            return (_pad8(name + b"\0") + struct.pack("<IBBBBII", off, 0, 0, 0, 0, 0, 0) + struct.pack("<IIII", 0, 0, 0, 0) + t)
        return struct.pack("<BBBBI", 0x16, 2, 0, 0, 12) + member(b"dataset", 0, ref) + member(b"dimension", 8, u32)

    def attribute(self, name, tmsg, smsg, data):
        nm = name.encode() + b"\0"
        if self.latest:
            return struct.pack("<BBHHHB", 3, 0, len(nm), len(tmsg), len(smsg), 0) + nm + tmsg + smsg + data
        return struct.pack("<BBHHH", 1, 0, len(nm), len(tmsg), len(smsg)) + _pad8(nm) + _pad8(tmsg) + _pad8(smsg) + data

    def att_from_value(self, name, value):
        if isinstance(value, (str, bytes)):
            b = value.encode() if isinstance(value, str) else value
            if len(b) == 0:
                return self.attribute(name, self.datatype(None, strlen=1), self.dataspace(None, null=True) if self.latest
                                      else self.dataspace((0,)), b"")
            return self.attribute(name, self.datatype(None, strlen=len(b)), self.dataspace(None), b)
        a = np.atleast_1d(np.asarray(value))
        return self.attribute(name, self.datatype(a.dtype), self.dataspace(a.shape), a.tobytes())

    # ------------------------------------------------------------------ object headers
    def object_header(self, msgs, force_v1=False):
        """msgs: list of (type, flags, body).  Version 2 (creation-order tracked) unless the style is plain."""
        if self.new and not force_v1:
            def enc(ms):
                out = b""
                for i, (t, fl, body) in enumerate(ms):
                    out += struct.pack("<BHBH", t, len(body), fl, i) + body
                return out
            first, rest = msgs, []
            if self.split_header and len(msgs) > 3:
                first, rest = msgs[:3], msgs[3:]
            cont_addr = None
            if rest:
                blk = b"OCHK" + enc(rest)
                blk += struct.pack("<I", lookup3(blk))
                cont_addr = self.alloc(blk)
                first = first + [(0x10, 0, struct.pack("<QQ", cont_addr, len(blk)))]
            body = enc(first)
            body += b"\0" * 3                       # a gap too small for a message header, as the library leaves them
            hdr = b"OHDR" + struct.pack("<BB", 2, 0x2E) + struct.pack("<IIII", 1600000000, 1600000000, 1600000000, 1600000000)
            hdr += struct.pack("<I", len(body)) + body
            hdr += struct.pack("<I", lookup3(hdr))
            return self.alloc(hdr)
        body = b""
        for (t, fl, b) in msgs:
            b = _pad8(b)
            body += struct.pack("<HHBBBB", t, len(b), fl, 0, 0, 0) + b
        hdr = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body
        return self.alloc(hdr)

    # ------------------------------------------------------------------ heaps
    def global_heap(self, objects):
        body = b""
        for i, o in enumerate(objects):
            body += struct.pack("<HHIQ", i + 1, 1, 0, len(o)) + _pad8(o)
        size = max(4096, 16 + len(body) + 16)
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\0" * (free - 16)
        return self.alloc(b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + body)

    def fractal_heap(self, objects, start=512, width=4, maxdirect=65536, maxbits=32):
        """Managed objects back to back in direct blocks; a root indirect block once one block is not enough."""
        boff = (maxbits + 7) // 8
        dhdr = 5 + 8 + boff + 4
        haddr_pos = len(self.buf) + (-len(self.buf) % 8)
        # header size: 14 + 8*? -- computed below; allocate the header FIRST so that blocks can point back at it
        hsize = 14 + 8 + 8 + 8 + 8 + 3 * 8 + 8 + 8 + 8 + 8 + 8 + 2 + 8 + 8 + 2 + 2 + 8 + 2 + 4
        haddr = self.alloc(b"\0" * hsize)
        assert haddr == haddr_pos
        # distribute objects over blocks of the doubling table
        blocks, cur, row_sizes = [], b"", []
        def bsize(i):
            row = i // width
            return start if row < 2 else start << (row - 1)
        for o in objects:
            while len(cur) + len(o) > bsize(len(blocks)) - dhdr:
                blocks.append(cur); cur = b""
                assert bsize(len(blocks)) <= maxdirect
            cur += o
        blocks.append(cur)
        offs, addrs, off = [], [], 0
        for i, b in enumerate(blocks):
            offs.append(off); off += bsize(i)
        alloc_space = off
        if len(blocks) == 1:
            blk = b"FHDB" + struct.pack("<BQ", 0, haddr) + (0).to_bytes(boff, "little") + b"\0\0\0\0" + blocks[0]
            blk += b"\0" * (start - len(blk))
            root = self.alloc(blk); currows = 0
        else:
            for i, b in enumerate(blocks):
                blk = b"FHDB" + struct.pack("<BQ", 0, haddr) + offs[i].to_bytes(boff, "little") + b"\0\0\0\0" + b
                blk += b"\0" * (bsize(i) - len(blk))
                addrs.append(self.alloc(blk))
            currows = (len(blocks) + width - 1) // width
            ent = b"".join(struct.pack("<Q", addrs[i] if i < len(addrs) else UNDEF) for i in range(currows * width))
            ib = b"FHIB" + struct.pack("<BQ", 0, haddr) + (0).to_bytes(boff, "little") + ent
            ib += struct.pack("<I", lookup3(ib))
            root = self.alloc(ib)
        h = b"FRHP" + struct.pack("<BHHBI", 0, 7, 0, 2, 4096)
        h += struct.pack("<QQQQ", 0, UNDEF, 0, UNDEF)
        h += struct.pack("<QQQ", alloc_space, alloc_space, alloc_space)
        h += struct.pack("<QQQQQ", len(objects), 0, 0, 0, 0)
        h += struct.pack("<HQQHH", width, start, maxdirect, maxbits, 1 if currows else 0)
        h += struct.pack("<QH", root, currows)
        h += struct.pack("<I", lookup3(h))
        assert len(h) == hsize, (len(h), hsize)
        self.buf[haddr:haddr + hsize] = h
        return haddr

    # ------------------------------------------------------------------ raw data
    def chunk_btree(self, entries, rank):
        """entries: (offsets, addr, nbytes, mask) sorted; returns the root node address."""
        def key(e):
            return struct.pack("<II", e[2], e[3]) + b"".join(struct.pack("<Q", o) for o in e[0]) + struct.pack("<Q", 0)
        def node(level, ents, lastkey):
            b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(ents), UNDEF, UNDEF)
            for k, child in ents:
                b += k + struct.pack("<Q", child)
            return self.alloc(b + lastkey)
        lastkey_all = struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in self._end_key) + struct.pack("<Q", 0)
        level, items = 0, [(key(e), e[1]) for e in entries]
        while True:
            groups = [items[i:i + self.leaf] for i in range(0, len(items), self.leaf)] or [[]]
            nxt = []
            for gi, g in enumerate(groups):
                last = groups[gi + 1][0][0] if gi + 1 < len(groups) else lastkey_all
                nxt.append((g[0][0] if g else lastkey_all, node(level, g, last)))
            if len(nxt) == 1:
                return nxt[0][1]
            items, level = nxt, level + 1

    def raw_data(self, v, arr):
        """Returns (layout message, filter message or None)."""
        es = arr.dtype.itemsize
        if v.layout == "compact":
            raw = arr.tobytes()
            return struct.pack("<BBH", 3, 0, len(raw)) + raw, None
        if v.chunks is None:
            if arr.size == 0:
                return struct.pack("<BBQQ", 3, 1, UNDEF, 0), None
            raw = arr.tobytes()
            return struct.pack("<BBQQ", 3, 1, self.alloc(raw), len(raw)), None
        ch = tuple(v.chunks)
        rank = arr.ndim
        filters = []
        if v.shuffle:
            filters.append((2, [es]))
        if v.deflate:
            filters.append((1, [v.deflate]))
        if v.fletcher:
            filters.append((3, []))
        entries = []
        nch = [(arr.shape[k] + ch[k] - 1) // ch[k] for k in range(rank)]
        skip = getattr(v, "skip_chunks", ())
        for ci in np.ndindex(*nch):
            if ci in skip:
                continue                     # an unwritten chunk: reads as the fill value
            off = tuple(ci[k] * ch[k] for k in range(rank))
            block = np.zeros(ch, arr.dtype)
            if v.fill is not None:
                block[...] = v.fill
            sl = tuple(slice(off[k], min(off[k] + ch[k], arr.shape[k])) for k in range(rank))
            sub = arr[sl]
            block[tuple(slice(0, s) for s in sub.shape)] = sub
            raw = block.tobytes()
            for fid, cd in filters:
                if fid == 2 and es > 1:
                    n = len(raw) // es
                    raw = np.frombuffer(raw, np.uint8).reshape(n, es).T.tobytes()
                elif fid == 1:
                    raw = zlib.compress(raw, cd[0])
                elif fid == 3:
                    raw = raw + struct.pack("<I", _fletcher32(raw))
            entries.append((off, self.alloc(raw, align=1), len(raw), 0))
        self._end_key = tuple(nch[k] * ch[k] for k in range(rank))
        root = self.chunk_btree(entries, rank) if entries else UNDEF
        lay = struct.pack("<BBBQ", 3, 2, rank + 1, root) + b"".join(struct.pack("<I", c) for c in ch) + struct.pack("<I", es)
        fm = None
        if filters:
            if self.latest:
                fm = struct.pack("<BB", 2, len(filters))
                for fid, cd in filters:
                    fm += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", c) for c in cd)
            else:
                fm = struct.pack("<BBHI", 1, len(filters), 0, 0)
                names = {1: b"deflate\0", 2: b"shuffle\0", 3: b"fletcher32\0"}
                for fid, cd in filters:
                    nm = _pad8(names[fid])
                    fm += struct.pack("<HHHH", fid, len(nm), 1, len(cd)) + nm + b"".join(struct.pack("<I", c) for c in cd)
                    if len(cd) & 1:
                        fm += b"\0" * 4
        return lay, fm

    def fill_message(self, arr, fill):
        if fill is None:
            if self.latest:
                return struct.pack("<BB", 3, 0x09)              # allocate late, write if set, no value defined
            return struct.pack("<BBBB", 2, 2, 2, 0)
        fb = np.asarray(fill, arr.dtype).tobytes()
        if self.latest:
            return struct.pack("<BBI", 3, 0x29, len(fb)) + fb
        return struct.pack("<BBBBI", 2, 2, 2, 1, len(fb)) + fb

    # ------------------------------------------------------------------ the file
    def write(self, path, dims, variables, gatts=None):
        """dims: ordered {name: length} (length None: unlimited, current extent taken from the variables)."""
        addrs = {}
        for _ in range(2):
            self.buf = bytearray()
            addrs = self._emit(dims, variables, gatts or {}, addrs)
        with open(path, "wb") as f:
            f.write(b"\0" * self.user_block)
            f.write(self.buf)

    def _emit(self, dims, variables, gatts, known):
        sb_size = 48 if self.latest else 96
        self.alloc(b"\0" * sb_size)
        addrs = {}
        dimnames = list(dims)
        byname = {v.name: v for v in variables}
        unlimited = {d for d in dimnames if dims[d] is None}
        extent = {}
        for d in dimnames:
            n = dims[d]
            if n is None:
                n = 0
                for v in variables:
                    if d in v.dims and v.data is not None:
                        n = max(n, np.asarray(v.data).shape[v.dims.index(d)])
            extent[d] = n
        links = []                                   # (name, address) in creation order
        netcdf = self.style != "plain"

        def dataset(name, arr, v, atts_msgs, shape, maxshape):
            lay, fm = self.raw_data(v, arr)
            msgs = [(0x01, 0, self.dataspace(shape, maxshape)), (0x03, 1, self.datatype(arr.dtype)),
                    (0x05, 1, self.fill_message(arr, v.fill))]
            if fm:
                msgs.append((0x0B, 1, fm))
            msgs.append((0x08, 0, lay))
            if self.new and len(atts_msgs) > 8:      # dense attribute storage
                heap = self.fractal_heap(atts_msgs)
                msgs.append((0x15, 0, struct.pack("<BBHQQQ", 0, 3, len(atts_msgs), heap, UNDEF, UNDEF)))
            else:
                msgs += [(0x0C, 0, a) for a in atts_msgs]
            addr = self.object_header(msgs)
            addrs[name] = addr
            links.append((name, addr))

        # the references each scale collects (REFERENCE_LIST) and each variable needs (DIMENSION_LIST)
        users = {d: [] for d in dimnames}
        for v in variables:
            if not (len(v.dims) == 1 and v.name == v.dims[0]):
                for k, d in enumerate(v.dims):
                    users[d].append((v.name, k))

        def stored_name(v):
            if v.name in dims and not (len(v.dims) == 1 and v.dims[0] == v.name):
                return "_nc4_non_coord_" + v.name
            return v.name

        def scale_atts(d, coord):
            out = [self.att_from_value("CLASS", b"DIMENSION_SCALE\0")]
            out.append(self.att_from_value("NAME", (d if coord else NOT_A_VAR + "%10d" % extent[d]).encode() + b"\0"))
            out.append(self.att_from_value("_Netcdf4Dimid", np.int32(dimnames.index(d))))
            if users[d]:
                data = b"".join(struct.pack("<QI", known.get(stored_name(byname[n]), 0), k) for n, k in users[d])
                out.append(self.attribute("REFERENCE_LIST", self.dtype_reflist(), self.dataspace((len(users[d]),)), data))
            return out

        order = []                                   # netCDF defines dimensions first, then variables
        if netcdf:
            for d in dimnames:
                if d not in byname or not (len(byname[d].dims) == 1 and byname[d].dims[0] == d):
                    order.append(("dim", d))
        order += [("var", v.name) for v in variables]
        for kind, name in order:
            if kind == "dim":
                dt = np.dtype(">f4")                 # nc4hdf.c writes dimension-only scales as big-endian floats
                arr = np.zeros((extent[name],), dt)
                v = Var(name, (name,), None, chunks=(max(extent[name], 1),) if name in unlimited else None)
                if name in unlimited:
                    v.skip_chunks = {(0,)}
                dataset(name, arr, v, scale_atts(name, False), (extent[name],), (None,) if name in unlimited else (extent[name],))
                continue
            v = byname[name]
            arr = np.asarray(v.data, order="C")
            if arr.dtype.kind == "S" and arr.dtype.itemsize != 1:
                raise ValueError("char variables are arrays of S1")
            if v.big_endian:
                arr = arr.astype(arr.dtype.newbyteorder(">"))
            atts = [self.att_from_value(k, val) for k, val in v.atts.items()]
            shape = arr.shape
            maxshape = tuple(None if (netcdf and d in unlimited) else s for d, s in zip(v.dims, shape)) if netcdf else shape
            if netcdf:
                coord = len(v.dims) == 1 and v.dims[0] == v.name
                if coord:
                    atts = scale_atts(v.name, True) + atts
                elif v.dims:
                    refs = [struct.pack("<Q", known.get(d, 0)) for d in v.dims]
                    gh = self.global_heap(refs)
                    data = b"".join(struct.pack("<IQI", 1, gh, i + 1) for i in range(len(refs)))
                    atts.append(self.attribute("DIMENSION_LIST", self.dtype_vlen_ref(), self.dataspace((len(refs),)), data))
                if any(m is None for m in maxshape) and v.chunks is None:
                    v.chunks = tuple(max(1, s) for s in shape)
            dataset(stored_name(v) if netcdf else v.name, arr, v, atts, shape, maxshape)

        # root group
        gmsgs_atts = [self.att_from_value(k, val) for k, val in gatts.items()]
        if netcdf:
            gmsgs_atts.insert(0, self.att_from_value("_NCProperties", b"version=2,netcdf=4.9.2,hdf5=1.12.2"))
        if self.new:
            lmsgs = []
            for i, (name, addr) in enumerate(links):
                nm = name.encode()
                lmsgs.append(struct.pack("<BBQB", 1, 0x04, i, len(nm)) + nm + struct.pack("<Q", addr))
            msgs = []
            if len(lmsgs) > 8:
                heap = self.fractal_heap(lmsgs)
                msgs.append((0x02, 0, struct.pack("<BBQQQQ", 0, 3, len(lmsgs), heap, UNDEF, UNDEF)))
            else:
                msgs.append((0x02, 0, struct.pack("<BBQQQQ", 0, 3, len(lmsgs), UNDEF, UNDEF, UNDEF)))
            msgs.append((0x0A, 1, struct.pack("<BB", 0, 0)))
            if len(lmsgs) <= 8:
                msgs += [(0x06, 0, m) for m in lmsgs]
            if len(gmsgs_atts) > 8:
                heap = self.fractal_heap(gmsgs_atts)
                msgs.append((0x15, 0, struct.pack("<BBHQQQ", 0, 3, len(gmsgs_atts), heap, UNDEF, UNDEF)))
            else:
                msgs += [(0x0C, 0, a) for a in gmsgs_atts]
            root = self.object_header(msgs)
            btree = heapaddr = None
        else:                                        # symbol table: names in a local heap, SNODs under a version 1 B-tree
            names = sorted(links)
            heap = bytearray(b"\0" * 8)
            offs = {}
            for name, _ in names:
                offs[name] = len(heap)
                heap += _pad8(name.encode() + b"\0")
            heap += b"\0" * 32
            daddr = self.alloc(bytes(heap))
            heapaddr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), UNDEF, daddr))
            per = 4
            snods = []
            for i in range(0, len(names), per):
                grp = names[i:i + per]
                b = b"SNOD" + struct.pack("<BBH", 1, 0, len(grp))
                for name, addr in grp:
                    b += struct.pack("<QQII", offs[name], addr, 0, 0) + b"\0" * 16
                b += b"\0" * (40 * (2 * per - len(grp)))
                snods.append((offs[grp[-1][0]], self.alloc(b)))
            b = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
            for lastoff, a in snods:
                b += struct.pack("<QQ", a, lastoff)
            btree = self.alloc(b)
            root = self.object_header([(0x11, 0, struct.pack("<QQ", btree, heapaddr))] + [(0x0C, 0, a) for a in gmsgs_atts])
        eof = len(self.buf)
        if self.latest:
            sb = b"\211HDF\r\n\032\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, root)
            sb += struct.pack("<I", lookup3(sb))
        else:
            sb = b"\211HDF\r\n\032\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
            if btree is None:
                sb += struct.pack("<QQII", 0, root, 0, 0) + b"\0" * 16
            else:
                sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heapaddr)
        assert len(sb) == sb_size
        self.buf[0:sb_size] = sb
        return addrs


def _fletcher32(data):
    """HDF5's Fletcher32 (H5_checksum_fletcher32): 16-bit big-endian words, an odd trailing byte in the high half."""
    s1 = s2 = 0
    n = len(data) // 2
    a = np.frombuffer(data[:2 * n], ">u2").astype(np.uint64)
    for w in a:
        s1 = (s1 + int(w)) % 65535
        s2 = (s2 + s1) % 65535
    if len(data) & 1:
        s1 = (s1 + (data[-1] << 8)) % 65535
        s2 = (s2 + s1) % 65535
    return (s2 << 16) | s1


def write_netcdf4(path, dims, variables, gatts=None, style="v18", **kw):
    H5Writer(style, **kw).write(path, dims, variables, gatts)


def from_classic(src, dst, style="v18", chunk=None, deflate=0, shuffle=False, **kw):
    """Re-express a classic netCDF file (scipy reads it) as netCDF-4: same dimensions, variables, attributes, values."""
    from scipy.io import netcdf_file
    with netcdf_file(src, "r", mmap=False) as f:
        dims = {d: f.dimensions[d] for d in f.dimensions}
        variables = []
        for name, var in f.variables.items():
            data = np.array(var[...]) if var.shape != () else np.array(var.getValue())
            if data.dtype.kind == "S":
                data = data.view("S1").reshape(var.shape)
            else:
                data = data.astype(data.dtype.newbyteorder("<"))
            atts = {k: (v if isinstance(v, (bytes, str)) else np.asarray(v)) for k, v in var._attributes.items()}
            ch = None
            if chunk and data.ndim >= 1 and data.dtype.kind != "S" and data.size > 16:
                ch = tuple(max(1, min(s, chunk)) for s in data.shape)
            variables.append(Var(name, var.dimensions, data, atts, chunks=ch, deflate=deflate if ch else 0,
                                 shuffle=shuffle and ch is not None))
        gatts = {k: (v if isinstance(v, (bytes, str)) else np.asarray(v)) for k, v in f._attributes.items()}
    write_netcdf4(dst, dims, variables, gatts, style=style, **kw)
