"""h5mini — a small, independent HDF5 *reader* in pure Python (struct + numpy), test
infrastructure only.

libhdf5 / h5py are not installed in this image, so the run.h5 files written by the product's
own dependency-free writer (fv2d_b200/host/H5Lite.h) are checked two ways:
  * this reader, written separately from the C++ code against the HDF5 File Format
    Specification (version 0 superblock, version 1 object headers, symbol-table groups, v1
    B-trees, local and global heaps), must parse them and return the expected content;
  * the same reader must parse a file produced by the real HDF5 library (the MATLAB 7.4
    fixture shipped inside scipy, when present), which pins its reading of the format.

Supported: superblock v0/v1 (optionally behind a user block), old-style groups, object header
v1 (+ continuation blocks), dataspace v1/v2, datatypes fixed-point / IEEE float / fixed string /
variable-length string, contiguous + compact + (unfiltered or deflate) chunked layout v3,
attribute messages v1/v2/v3.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def _pad8(n):
    return (n + 7) & ~7


class Datatype:
    def __init__(self, cls, size, np_dtype=None, is_vlen_str=False, strpad=0, cset=0, base=None):
        self.cls, self.size, self.np_dtype = cls, size, np_dtype
        self.is_vlen_str, self.strpad, self.cset, self.base = is_vlen_str, strpad, cset, base


class Node:
    """A group or a dataset."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.attrs = {}
        self.messages = []
        self.btree = self.heap = None
        self.shape = self.dtype = self.layout = None
        self.filters = []
        f._read_object_header(self)

    @property
    def is_group(self):
        return self.btree is not None

    def keys(self):
        return [k for k, _ in self.f._group_entries(self)]

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            ent = dict(node.f._group_entries(node))
            if part not in ent:
                raise KeyError(path)
            node = Node(node.f, ent[part])
        return node

    def read(self):
        if self.is_group:
            raise H5Error("not a dataset")
        return self.f._read_dataset(self)


class File(Node):
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        off = 0
        while True:
            if self.buf[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise H5Error("HDF5 signature not found")
        self.sb_off = off
        b = self.buf
        self.sb_version = b[off + 8]
        if self.sb_version > 1:
            raise H5Error(f"superblock version {self.sb_version} not supported")
        self.size_off, self.size_len = b[off + 13], b[off + 14]
        if (self.size_off, self.size_len) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, off + 16)
        self.consistency_flags = struct.unpack_from("<I", b, off + 20)[0]
        p = off + 24 + (4 if self.sb_version == 1 else 0)
        self.base, self.free_addr, self.eof, self.driver = struct.unpack_from("<4Q", b, p)
        p += 32
        # root symbol table entry
        self.root_link_off, root_addr, self.root_cache_type = struct.unpack_from("<QQI", b, p)
        self.root_scratch = struct.unpack_from("<QQ", b, p + 24)
        Node.__init__(self, self, root_addr)

    # ---- low level
    def _at(self, addr):
        if addr == UNDEF:
            raise H5Error("undefined address dereferenced")
        return self.base + addr

    def _read_object_header(self, node):
        b = self.buf
        p = self._at(node.addr)
        version, _, nmsg, refcnt, hsize = struct.unpack_from("<BBHII", b, p)
        if version != 1:
            raise H5Error(f"object header version {version} at {node.addr:#x}")
        node.refcount = refcnt
        blocks = [(p + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, q)
                body = b[q + 8:q + 8 + msize]
                q += 8 + msize
                seen += 1
                node.messages.append((mtype, mflags, body))
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._at(caddr), clen))
        for mtype, mflags, body in node.messages:
            if mtype == 0x0001:
                node.shape = self._parse_dataspace(body)
            elif mtype == 0x0003:
                node.dtype = self._parse_datatype(body)[0]
            elif mtype == 0x0008:
                node.layout = self._parse_layout(body)
            elif mtype == 0x000B:
                node.filters = self._parse_filters(body)
            elif mtype == 0x000C:
                name, val = self._parse_attribute(body)
                node.attrs[name] = val
            elif mtype == 0x0011:
                node.btree, node.heap = struct.unpack_from("<QQ", body, 0)

    def _parse_dataspace(self, m):
        ver, rank, flags = m[0], m[1], m[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if m[3] == 2:  # null dataspace
                return None
            p = 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(struct.unpack_from(f"<{rank}Q", m, p)) if rank else ()

    def _parse_datatype(self, m):
        cv, b0, b1, b2 = m[0], m[1], m[2], m[3]
        cls, ver = cv & 0x0F, cv >> 4
        size = struct.unpack_from("<I", m, 4)[0]
        if ver not in (1, 2, 3):
            raise H5Error(f"datatype version {ver}")
        end = ">" if (b0 & 1) else "<"
        if cls == 0:
            _, prec = struct.unpack_from("<HH", m, 8)
            signed = bool(b0 & 0x08)
            if prec != 8 * size:
                raise H5Error("fixed-point with padding bits")
            return Datatype(cls, size, np.dtype(f"{end}{'i' if signed else 'u'}{size}")), 12
        if cls == 1:
            _, prec, eloc, esize, mloc, msize, ebias = struct.unpack_from("<HHBBBBI", m, 8)
            ieee = {4: (32, 23, 8, 0, 23, 127), 8: (64, 52, 11, 0, 52, 1023)}.get(size)
            if ieee is None or (prec, eloc, esize, mloc, msize, ebias) != ieee:
                raise H5Error("non-IEEE floating point type")
            if (b0 >> 4) & 3 != 2 or b1 != prec - 1:
                raise H5Error("unexpected float normalisation / sign location")
            return Datatype(cls, size, np.dtype(f"{end}f{size}")), 20
        if cls == 3:
            return Datatype(cls, size, np.dtype(f"S{size}"), strpad=b0 & 0x0F, cset=b0 >> 4), 8
        if cls == 9:
            base, used = self._parse_datatype(m[8:])
            kind = b0 & 0x0F
            if kind != 1:
                raise H5Error("variable-length sequences not supported")
            return Datatype(cls, size, None, is_vlen_str=True, strpad=b0 >> 4, cset=b1 & 0x0F, base=base), 8 + used
        raise H5Error(f"datatype class {cls} not supported")

    def _parse_layout(self, m):
        ver = m[0]
        if ver in (1, 2):  # HDF5 1.6-era files
            ndim, cls = m[1], m[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack_from("<Q", m, p)[0]
                p += 8
            dims = struct.unpack_from(f"<{ndim}I", m, p)
            p += 4 * ndim
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                return ("chunked", addr, dims)
            n = struct.unpack_from("<I", m, p)[0]
            return ("compact", m[p + 4:p + 4 + n])
        if ver != 3:
            raise H5Error(f"layout version {ver}")
        cls = m[1]
        if cls == 0:
            n = struct.unpack_from("<H", m, 2)[0]
            return ("compact", m[4:4 + n])
        if cls == 1:
            addr, size = struct.unpack_from("<QQ", m, 2)
            return ("contiguous", addr, size)
        if cls == 2:
            rank = m[2]
            addr = struct.unpack_from("<Q", m, 3)[0]
            dims = struct.unpack_from(f"<{rank}I", m, 11)
            return ("chunked", addr, dims)
        raise H5Error("layout class")

    def _parse_filters(self, m):
        ver, n = m[0], m[1]
        if ver != 1:
            raise H5Error("filter pipeline version")
        p, out = 8, []
        for _ in range(n):
            fid, nlen, flags, ncd = struct.unpack_from("<HHHH", m, p)
            p += 8 + _pad8(nlen)
            cd = struct.unpack_from(f"<{ncd}I", m, p)
            p += 4 * ncd + (4 if ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def _parse_attribute(self, m):
        ver = m[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", m, 2)
            p = 8
            name = m[p:p + nsz].split(b"\0")[0].decode()
            p += _pad8(nsz)
            dt, _ = self._parse_datatype(m[p:p + tsz])
            p += _pad8(tsz)
            shape = self._parse_dataspace(m[p:p + ssz])
            p += _pad8(ssz)
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", m, 2)
            p = 8 + (1 if ver == 3 else 0)
            name = m[p:p + nsz].split(b"\0")[0].decode()
            p += nsz
            dt, _ = self._parse_datatype(m[p:p + tsz])
            p += tsz
            shape = self._parse_dataspace(m[p:p + ssz])
            p += ssz
        else:
            raise H5Error(f"attribute version {ver}")
        return name, self._decode(dt, shape, m[p:])

    def _decode(self, dt, shape, raw):
        n = int(np.prod(shape)) if shape else 1
        if dt.is_vlen_str:
            out = []
            for k in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * k)
                out.append(self._global_heap_object(addr, idx)[:ln].decode("utf-8" if dt.cset else "ascii"))
            return out[0] if shape == () else np.array(out, dtype=object).reshape(shape)
        arr = np.frombuffer(raw[:n * dt.size], dtype=dt.np_dtype)
        if dt.cls == 3:
            vals = [bytes(x).split(b"\0")[0].decode() for x in arr]
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        return arr[0] if shape == () else arr.reshape(shape).copy()

    def _global_heap_object(self, addr, idx):
        b = self.buf
        p = self._at(addr)
        if b[p:p + 4] != b"GCOL":
            raise H5Error("bad global heap signature")
        size = struct.unpack_from("<Q", b, p + 8)[0]
        q, end = p + 16, p + size
        while q + 16 <= end:
            oid, _ref, _, osz = struct.unpack_from("<HHIQ", b, q)
            if oid == idx:
                return b[q + 16:q + 16 + osz]
            if oid == 0:
                break
            q += 16 + _pad8(osz)
        raise H5Error("global heap object not found")

    def _heap_string(self, heap_addr, off):
        b = self.buf
        p = self._at(heap_addr)
        if b[p:p + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        dsize, _free, daddr = struct.unpack_from("<QQQ", b, p + 8)
        q = self._at(daddr) + off
        e = b.index(b"\0", q)
        return b[q:e].decode()

    def local_heap_info(self, heap_addr):
        p = self._at(heap_addr)
        return struct.unpack_from("<QQQ", self.buf, p + 8)

    def _group_entries(self, node):
        """[(name, object header address)] in B-tree (= name) order."""
        if node.btree is None:
            raise H5Error("not a group")
        out = []
        self._walk_group_btree(node.btree, node.heap, out, [])
        return out

    def _walk_group_btree(self, addr, heap, out, keycheck):
        b = self.buf
        p = self._at(addr)
        if b[p:p + 4] == b"SNOD":
            ver, _, nsym = struct.unpack_from("<BBH", b, p + 4)
            if ver != 1:
                raise H5Error("SNOD version")
            for k in range(nsym):
                noff, oaddr = struct.unpack_from("<QQ", b, p + 8 + 40 * k)
                out.append((self._heap_string(heap, noff), oaddr))
            return
        if b[p:p + 4] != b"TREE":
            raise H5Error(f"bad B-tree signature at {addr:#x}")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        q = p + 24
        for k in range(used):
            key_lo, child = struct.unpack_from("<QQ", b, q + 16 * k)
            key_hi = struct.unpack_from("<Q", b, q + 16 * k + 16)[0]
            n0 = len(out)
            self._walk_group_btree(child, heap, out, keycheck)
            lo, hi = self._heap_string(heap, key_lo), self._heap_string(heap, key_hi)
            for name, _ in out[n0:]:
                if not (lo < name <= hi):
                    raise H5Error(f"B-tree key order violated: {lo!r} < {name!r} <= {hi!r}")

    def _read_dataset(self, node):
        dt, shape, lay = node.dtype, node.shape, node.layout
        n = int(np.prod(shape)) if shape else 1
        if lay[0] == "compact":
            raw = lay[1]
        elif lay[0] == "contiguous":
            if lay[1] == UNDEF:
                raw = bytes(n * dt.size)
            else:
                size = lay[2] if lay[2] is not None else n * dt.size
                raw = self.buf[self._at(lay[1]):self._at(lay[1]) + size]
        else:
            return self._read_chunked(node)
        return self._decode(dt, shape, raw)

    def _read_chunked(self, node):
        dt, shape = node.dtype, node.shape
        _, addr, cdims = node.layout
        rank = len(shape)
        out = np.zeros(shape, dtype=dt.np_dtype)
        chunks = []
        self._walk_chunk_btree(addr, rank, chunks)
        for off, size, fmask, caddr in chunks:
            raw = self.buf[self._at(caddr):self._at(caddr) + size]
            for fid, _ in reversed(node.filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                else:
                    raise H5Error(f"filter {fid} not supported")
            blk = np.frombuffer(raw, dtype=dt.np_dtype)[:int(np.prod(cdims[:rank]))].reshape(cdims[:rank])
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(off, cdims, shape))
            out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def _walk_chunk_btree(self, addr, rank, out):
        b = self.buf
        p = self._at(addr)
        if b[p:p + 4] != b"TREE":
            raise H5Error("bad chunk B-tree")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        ksz = 8 + 8 * (rank + 1)
        q = p + 24
        for k in range(used):
            size, fmask = struct.unpack_from("<II", b, q)
            off = struct.unpack_from(f"<{rank}Q", b, q + 8)
            child = struct.unpack_from("<Q", b, q + ksz)[0]
            if level == 0:
                out.append((off, size, fmask, child))
            else:
                self._walk_chunk_btree(child, rank, out)
            q += ksz + 8


def dump(path):
    f = File(path)
    print(f"superblock v{f.sb_version} at {f.sb_off}, base {f.base:#x}, eof {f.eof:#x}, K=({f.leaf_k},{f.internal_k})")

    def rec(node, indent):
        for k, v in node.attrs.items():
            print(f"{indent}@{k} = {v!r}")
        if node.is_group:
            for name in node.keys():
                child = node[name]
                print(f"{indent}{name}{'/' if child.is_group else ''}"
                      + ("" if child.is_group else f"  shape={child.shape} layout={child.layout[0]}"))
                rec(child, indent + "  ")

    rec(f, "")
    return f


if __name__ == "__main__":
    import sys

    dump(sys.argv[1])
