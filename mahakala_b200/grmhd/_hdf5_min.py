"""Minimal read-only HDF5 reader (pure Python + NumPy) for AthenaK ``.athdf`` dumps.

The reference opens the dump with h5py (``/root/reference/mahakala/grmhd/athenak.py:24, 79-103``: datasets
``x1v x2v x3v x1f x2f x3f uov B LogicalLocations Levels`` plus the root attribute ``VariableNames``).  h5py is
not part of this image, so ``AthenakFluidModel`` falls back to this reader when ``import h5py`` fails.  It covers
the subset of the HDF5 file format that h5py / libhdf5 produce for such files:

* superblock versions 0-3 (with a user block / non-zero base address);
* groups: old-style (symbol-table message -> version-1 B-tree + local heap + SNOD nodes) and new-style with
  compact link messages;
* object headers version 1 and 2 (continuation blocks included);
* datasets: compact, contiguous and chunked (version-1 chunk B-tree) layouts, deflate / shuffle / fletcher32
  filters, fixed-point, floating-point and fixed-length string types, either byte order;
* attributes (message versions 1-3) with the same types plus variable-length strings (global heap).

Anything else (dense link storage, version-4 layouts, compound types, ...) raises ``Hdf5FormatError`` naming the
unsupported feature instead of returning wrong data.
"""
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"


class Hdf5FormatError(ValueError):
    pass


class _Datatype:
    def __init__(self, np_dtype=None, vlen_string=False, size=0):
        self.np_dtype, self.vlen_string, self.size = np_dtype, vlen_string, size


class Hdf5File:
    """``f = Hdf5File(path); f["uov"] -> ndarray; f.attrs["VariableNames"]; f.keys()``."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self._read_superblock()
        self._links, self.attrs = self._read_group(self.root_header)

    # ---- low level -------------------------------------------------------------------------------------
    def _u(self, pos, n):
        return int.from_bytes(self.buf[pos:pos + n], "little")

    def _addr(self, pos):
        """file address stored at pos (relative to the base address) -> absolute position, or None"""
        a = self._u(pos, self.O)
        if a == (1 << (8 * self.O)) - 1:
            return None
        return a + self.base

    def _read_superblock(self):
        b = self.buf
        pos = 0
        while True:                                   # the signature sits at 0, 512, 1024, 2048, ...
            if b[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(b):
                raise Hdf5FormatError("not an HDF5 file (signature not found)")
        self.sb_pos = pos
        ver = b[pos + 8]
        if ver in (0, 1):
            self.O, self.L = b[pos + 13], b[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            self.base = 0
            base = self._u(p, self.O)
            self.base = base
            p += 4 * self.O                           # base, free-space, end-of-file, driver-info addresses
            # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
            self.root_header = self._addr(p + self.O)
        elif ver in (2, 3):
            self.O, self.L = b[pos + 9], b[pos + 10]
            p = pos + 12
            self.base = 0
            self.base = self._u(p, self.O)
            self.root_header = self._addr(p + 3 * self.O)
        else:
            raise Hdf5FormatError(f"unsupported superblock version {ver}")
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise Hdf5FormatError("unsupported size of offsets / lengths")

    # ---- object headers --------------------------------------------------------------------------------
    def _messages(self, pos):
        """[(type, flags, data_pos, size)] of the object header at pos (v1 or v2, continuations followed)."""
        b = self.buf
        out = []
        if b[pos:pos + 4] == b"OHDR":
            if b[pos + 4] != 2:
                raise Hdf5FormatError("unsupported object header version")
            flags = b[pos + 5]
            p = pos + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            n = 1 << (flags & 3)
            size0 = self._u(p, n)
            p += n
            blocks = [(p, size0)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], self._u(p + 1, 2), b[p + 3]
                    p += 4 + (2 if track_order else 0)
                    if mtype == 0x10:
                        cpos, clen = self._addr(p), self._u(p + self.O, self.L)
                        if b[cpos:cpos + 4] != b"OCHK":
                            raise Hdf5FormatError("bad object header continuation block")
                        blocks.append((cpos + 4, clen - 8))          # minus signature and checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        if b[pos] != 1:
            raise Hdf5FormatError(f"unsupported object header version {b[pos]}")
        nmsg = self._u(pos + 2, 2)
        size = self._u(pos + 8, 4)
        blocks = [(pos + 16, size)]
        while blocks and nmsg > 0:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and nmsg > 0:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), b[p + 4]
                p += 8
                nmsg -= 1
                if mtype == 0x10:
                    blocks.append((self._addr(p), self._u(p + self.O, self.L)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # ---- groups ------------------------------------------------------------------------------------------
    def _read_group(self, header_pos):
        links, attrs = {}, {}
        for mtype, mflags, p, size in self._messages(header_pos):
            if mtype == 0x11:                                         # symbol table: B-tree + local heap
                btree, heap = self._addr(p), self._addr(p + self.O)
                if self.buf[heap:heap + 4] != b"HEAP":
                    raise Hdf5FormatError("bad local heap")
                heap_data = self._addr(heap + 8 + 2 * self.L)
                self._walk_group_btree(btree, heap_data, links)
            elif mtype == 0x06:                                       # link message
                name, target = self._link_message(p)
                if target is not None:
                    links[name] = target
            elif mtype == 0x02:                                       # link info: dense storage if a heap exists
                flags = self.buf[p + 1]
                q = p + 2 + (8 if flags & 1 else 0)
                if self._addr(q) is not None:
                    raise Hdf5FormatError("dense link storage (fractal heap) is not supported")
            elif mtype == 0x0C:
                k, v = self._attribute(p)
                attrs[k] = v
        return links, attrs

    def _walk_group_btree(self, pos, heap_data, links):
        b = self.buf
        if b[pos:pos + 4] == b"SNOD":
            n = self._u(pos + 6, 2)
            p = pos + 8
            for _ in range(n):
                name_off = self._u(p, self.O)
                header = self._addr(p + self.O)
                end = b.index(b"\x00", heap_data + name_off)
                links[b[heap_data + name_off:end].decode("utf-8")] = header
                p += 2 * self.O + 24
            return
        if b[pos:pos + 4] != b"TREE" or b[pos + 4] != 0:
            raise Hdf5FormatError("bad group B-tree node")
        n = self._u(pos + 6, 2)
        p = pos + 8 + 2 * self.O + self.L                             # skip siblings and key 0
        for _ in range(n):
            self._walk_group_btree(self._addr(p), heap_data, links)
            p += self.O + self.L

    def _link_message(self, p):
        b = self.buf
        if b[p] != 1:
            raise Hdf5FormatError("unsupported link message version")
        flags = b[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        n = 1 << (flags & 3)
        nlen = self._u(q, n)
        q += n
        name = b[q:q + nlen].decode("utf-8")
        q += nlen
        return name, (self._addr(q) if ltype == 0 else None)          # soft / external links are ignored

    # ---- datatypes, dataspaces -------------------------------------------------------------------------
    def _datatype(self, p):
        b = self.buf
        cls, ver = b[p] & 0x0F, b[p] >> 4
        bits = self._u(p + 1, 3)
        size = self._u(p + 4, 4)
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return _Datatype(np.dtype(f"{order}{'i' if bits & 0x08 else 'u'}{size}"), size=size)
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5FormatError(f"unsupported float size {size}")
            return _Datatype(np.dtype(f"{order}f{size}"), size=size)
        if cls == 3:
            return _Datatype(np.dtype(f"S{size}"), size=size)
        if cls == 9 and (bits & 0x0F) == 1:                           # variable-length string
            return _Datatype(None, vlen_string=True, size=size)
        raise Hdf5FormatError(f"unsupported datatype class {cls} (version {ver})")

    def _dataspace(self, p):
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if b[p + 3] == 2:                                         # null dataspace
                return None
            q = p + 4
        else:
            raise Hdf5FormatError(f"unsupported dataspace version {ver}")
        return tuple(self._u(q + i * self.L, self.L) for i in range(rank))

    def _vlen_strings(self, p, count):
        out = []
        for i in range(count):
            q = p + i * (4 + self.O + 4)
            coll, idx = self._addr(q + 4), self._u(q + 4 + self.O, 4)
            out.append(self._global_heap_object(coll, idx) if coll is not None else b"")
        return out

    def _global_heap_object(self, pos, index):
        b = self.buf
        if b[pos:pos + 4] != b"GCOL":
            raise Hdf5FormatError("bad global heap collection")
        end = pos + self._u(pos + 8, self.L)
        p = pos + 8 + self.L
        while p + 8 + self.L <= end:
            idx, size = self._u(p, 2), self._u(p + 8, self.L)
            if idx == 0:
                break
            if idx == index:
                return b[p + 8 + self.L:p + 8 + self.L + size]
            p += 8 + self.L + ((size + 7) // 8) * 8
        raise Hdf5FormatError("global heap object not found")

    def _attribute(self, p):
        b = self.buf
        ver = b[p]
        nsz, tsz, ssz = self._u(p + 2, 2), self._u(p + 4, 2), self._u(p + 6, 2)
        if ver == 1:
            pad = lambda n: (n + 7) // 8 * 8
            q = p + 8
        elif ver in (2, 3):
            if b[p + 1] & 0x03:
                raise Hdf5FormatError("shared attribute datatypes / dataspaces are not supported")
            pad = lambda n: n
            q = p + 8 + (1 if ver == 3 else 0)
        else:
            raise Hdf5FormatError(f"unsupported attribute message version {ver}")
        name = b[q:q + nsz].split(b"\x00")[0].decode("utf-8")
        q += pad(nsz)
        dt = self._datatype(q)
        q += pad(tsz)
        shape = self._dataspace(q)
        q += pad(ssz)
        if shape is None:
            return name, None
        count = int(np.prod(shape)) if shape else 1
        if dt.vlen_string:
            vals = np.array(self._vlen_strings(q, count), dtype=object).reshape(shape)
        else:
            vals = np.frombuffer(b, dtype=dt.np_dtype, count=count, offset=q).reshape(shape).copy()
        return name, (vals if shape else vals.reshape(())[()])

    # ---- datasets -----------------------------------------------------------------------------------------
    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        if name not in self._links:
            raise KeyError(name)
        return self._read_dataset(self._links[name])

    def _read_dataset(self, header_pos):
        b = self.buf
        dt = shape = layout = None
        filters = []
        for mtype, mflags, p, size in self._messages(header_pos):
            if mtype == 0x01:
                shape = self._dataspace(p)
            elif mtype == 0x03:
                dt = self._datatype(p)
            elif mtype == 0x08:
                layout = p
            elif mtype == 0x0B:
                filters = self._filters(p)
            elif mtype == 0x11:
                raise Hdf5FormatError("object is a group, not a dataset")
        if dt is None or shape is None or layout is None:
            raise Hdf5FormatError("object lacks a datatype, dataspace or layout message")
        if dt.vlen_string:
            raise Hdf5FormatError("variable-length string datasets are not supported")
        count = int(np.prod(shape)) if shape else 1
        ver = b[layout]
        if ver in (1, 2):                                             # libhdf5 <= 1.6 layout message
            ndim, cls = b[layout + 1], b[layout + 2]
            q = layout + 8
            addr = None
            if cls != 0:
                addr = self._addr(q)
                q += self.O
            cdims = tuple(self._u(q + 4 * i, 4) for i in range(ndim))
            compact_pos = q + 4 * ndim + 4
            btree = addr
        elif ver == 3:
            cls = b[layout + 1]
            compact_pos = layout + 4
            addr = self._addr(layout + 2) if cls == 1 else None
            if cls == 2:
                ndim = b[layout + 2]
                btree = self._addr(layout + 3)
                cdims = tuple(self._u(layout + 3 + self.O + 4 * i, 4) for i in range(ndim))
        else:
            raise Hdf5FormatError(f"unsupported data layout message version {ver}")
        if cls == 0:                                                  # compact
            return np.frombuffer(b, dtype=dt.np_dtype, count=count, offset=compact_pos).reshape(shape).copy()
        if cls == 1:                                                  # contiguous
            if addr is None:                                          # never written: fill value (zeros)
                return np.zeros(shape, dtype=dt.np_dtype)
            return np.frombuffer(b, dtype=dt.np_dtype, count=count, offset=addr).reshape(shape).copy()
        if cls == 2:                                                  # chunked, version-1 B-tree index
            if ndim - 1 != len(shape) or cdims[-1] != dt.size:
                raise Hdf5FormatError("inconsistent chunk dimensions")
            out = np.zeros(shape, dtype=dt.np_dtype)
            if btree is not None:
                self._walk_chunk_btree(btree, ndim, cdims[:-1], dt, filters, out)
            return out
        raise Hdf5FormatError(f"unsupported data layout class {cls}")

    def _filters(self, p):
        b = self.buf
        ver, n = b[p], b[p + 1]
        out = []
        q = p + (8 if ver == 1 else 2)
        for _ in range(n):
            fid = self._u(q, 2)
            q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = self._u(q, 2)
                q += 2
            q += 2                                                    # flags
            ncd = self._u(q, 2)
            q += 2
            q += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [self._u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _walk_chunk_btree(self, pos, ndim, cshape, dt, filters, out):
        b = self.buf
        if b[pos:pos + 4] != b"TREE" or b[pos + 4] != 1:
            raise Hdf5FormatError("bad chunk B-tree node")
        level, n = b[pos + 5], self._u(pos + 6, 2)
        keysize = 8 + 8 * ndim
        p = pos + 8 + 2 * self.O
        for _ in range(n):
            nbytes, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = tuple(self._u(p + 8 + 8 * i, 8) for i in range(ndim - 1))
            child = self._addr(p + keysize)
            if level > 0:
                self._walk_chunk_btree(child, ndim, cshape, dt, filters, out)
            else:
                raw = b[child:child + nbytes]
                for k in range(len(filters) - 1, -1, -1):
                    if mask & (1 << k):
                        continue
                    fid, cd = filters[k]
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:                                    # shuffle: bytes were transposed per element
                        es = cd[0] if cd else dt.size
                        a = np.frombuffer(raw, dtype=np.uint8)
                        nel = a.size // es
                        raw = a[:nel * es].reshape(es, nel).T.tobytes() + a[nel * es:].tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise Hdf5FormatError(f"unsupported filter id {fid}")
                chunk = np.frombuffer(raw, dtype=dt.np_dtype, count=int(np.prod(cshape))).reshape(cshape)
                sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, out.shape))
                sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
                out[sel_out] = chunk[sel_in]
            p += keysize + self.O


def read_athdf(filename):
    """The datasets and the ``VariableNames`` attribute ``AthenakFluidModel`` needs (athenak.py:79-103)."""
    f = Hdf5File(filename)
    out = {k: np.array(f[k]) for k in ('x1v', 'x2v', 'x3v', 'x1f', 'x2f', 'x3f', 'uov', 'B', 'LogicalLocations',
                                       'Levels')}
    names = f.attrs.get('VariableNames')
    if names is None:
        raise Hdf5FormatError("root attribute VariableNames is missing")
    out['VariableNames'] = [n.decode('utf-8') if isinstance(n, bytes) else str(n) for n in np.asarray(names).ravel()]
    return out
