"""An independent, read-only walk of the HDF5 file format in pure Python (struct), written against the format
specification and first validated on a file that genuine libhdf5 wrote (the MATLAB v7.3 sample shipped with scipy).
It shares no code with xpcs-eigen_b200/host/h5lite.cpp and is used to check the files h5lite WRITES: superblock v0,
symbol-table groups (v1 B-tree, local heap, SNOD), v1 object headers with dataspace / datatype / fill / layout /
filter-pipeline / attribute messages, contiguous and chunked (v1 chunk B-tree, deflate + shuffle) raw data.
Test infrastructure only."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5Check:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        off = 0
        while self.b[off: off + 8] != SIG:
            off = 512 if off == 0 else off * 2
            if off >= len(self.b):
                raise ValueError("no HDF5 signature")
        self.sb = off
        ver = self.b[off + 8]
        assert ver in (0, 1), "superblock version %d" % ver
        self.so, self.sl = self.b[off + 13], self.b[off + 14]
        assert self.so == 8 and self.sl == 8
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = self.u64(p)
        if self.base == UNDEF:
            self.base = off
        self.eof = self.u64(p + 16)
        self.root_ohdr = self.u64(p + 32 + 8)   # root symbol table entry: link name offset, object header address
        self.datasets, self.groups, self.attrs = {}, [], {}
        # (the genuine sample, with its 512-byte user block, stores the file length here, not length - base)
        assert self.eof <= len(self.b) + 7, "end-of-file address beyond the file"
        self._walk(self.root_ohdr, "")

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def u64(self, o):
        return struct.unpack_from("<Q", self.b, o)[0]

    def _messages(self, addr):
        a = self.base + addr
        assert self.b[a] == 1, "object header version"
        nmsg, hsize = self.u16(a + 2), self.u32(a + 8)
        blocks, out = [(a + 16, hsize)], []
        for start, size in blocks:
            p = start
            while p + 8 <= start + size and len(out) < nmsg:
                t, sz, fl = self.u16(p), self.u16(p + 2), self.b[p + 4]
                assert sz % 8 == 0, "message size not a multiple of 8 in a version-1 header"
                if t == 0x10:
                    blocks.append((self.base + self.u64(p + 8), self.u64(p + 16)))
                out.append((t, p + 8, sz, fl))
                p += 8 + sz
        assert len(out) == nmsg, "object header holds %d of %d messages" % (len(out), nmsg)
        return out

    def _walk(self, addr, path):
        msgs = self._messages(addr)
        types = {m[0] for m in msgs}
        self.attrs[path or "/"] = [self.b[p: p + sz] for t, p, sz, _ in msgs if t == 0x0C]
        if 0x11 in types:
            self.groups.append(path or "/")
            p = [m for m in msgs if m[0] == 0x11][0][1]
            self._group(self.u64(p), self.u64(p + 8), path)
        else:
            self.datasets[path] = self._dataset(msgs)

    def _group(self, btree, heap, path):
        h = self.base + heap
        assert self.b[h: h + 4] == b"HEAP"
        hdata = self.base + self.u64(h + 24)
        names = []

        def node(addr):
            t = self.base + addr
            assert self.b[t: t + 4] == b"TREE" and self.b[t + 4] == 0
            level, n = self.b[t + 5], self.u16(t + 6)
            p = t + 24 + 8
            for _ in range(n):
                child = self.u64(p)
                p += 16
                if level:
                    node(child)
                    continue
                s = self.base + child
                assert self.b[s: s + 4] == b"SNOD"
                for i in range(self.u16(s + 6)):
                    ep = s + 8 + 40 * i
                    no, oa = self.u64(ep), self.u64(ep + 8)
                    end = self.b.index(b"\0", hdata + no)
                    name = self.b[hdata + no: end].decode()
                    names.append(name)
                    self._walk(oa, path + "/" + name)

        node(btree)
        assert names == sorted(names), "symbol table entries must be in name order: %s" % names

    def _dataset(self, msgs):
        m = {t: (p, sz) for t, p, sz, _ in msgs}
        sp, _ = m[0x01]
        sver, rank = self.b[sp], self.b[sp + 1]
        dp = sp + (8 if sver == 1 else 4)
        dims = [self.u64(dp + 8 * i) for i in range(rank)]
        tp, _ = m[0x03]
        cls, bits0, size = self.b[tp] & 15, self.b[tp + 1], self.u32(tp + 4)
        if cls == 0:
            dt = np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
        elif cls == 1:
            dt = np.dtype("<f%d" % size)
        elif cls == 3:
            dt = np.dtype("S%d" % size)
        else:
            raise ValueError("datatype class %d" % cls)
        n = int(np.prod(dims)) if rank else 1
        lp, _ = m[0x08]
        if self.b[lp] in (1, 2):   # older layout message (the genuine sample): dimensionality, class, then the address
            assert self.b[lp + 2] == 1, "only contiguous data in a version-1/2 layout here"
            addr = self.u64(lp + 8)
            raw = self.b[self.base + addr: self.base + addr + n * size]
            return np.frombuffer(raw, dt).reshape(dims).copy()
        assert self.b[lp] == 3, "layout version"
        lcls = self.b[lp + 1]
        if lcls == 1:
            addr, nbytes = self.u64(lp + 2), self.u64(lp + 10)
            assert nbytes == n * size
            raw = self.b[self.base + addr: self.base + addr + nbytes] if addr != UNDEF else bytes(n * size)
        elif lcls == 0:
            sz = self.u16(lp + 2)
            raw = self.b[lp + 4: lp + 4 + sz]
        else:
            assert self.b[lp + 2] == rank + 1
            bt = self.u64(lp + 3)
            cd = [self.u32(lp + 11 + 4 * i) for i in range(rank + 1)]
            assert cd[-1] == size, "last chunk extent = element size"
            filters = []
            if 0x0B in m:
                pp, _ = m[0x0B]
                assert self.b[pp] == 1
                f = pp + 8
                for _ in range(self.b[pp + 1]):
                    fid, nlen, ncd = self.u16(f), self.u16(f + 2), self.u16(f + 6)
                    f += 8 + ((nlen + 7) & ~7) + 4 * ncd + (4 if ncd & 1 else 0)
                    filters.append(fid)
            out = np.zeros(dims, dt)
            self._chunks(bt, rank, dims, cd[:-1], size, filters, out, dt)
            return out
        return np.frombuffer(raw[: n * size], dt).reshape(dims).copy()

    def _chunks(self, addr, rank, dims, cd, es, filters, out, dt):
        if addr == UNDEF:
            return
        t = self.base + addr
        assert self.b[t: t + 4] == b"TREE" and self.b[t + 4] == 1
        level, n = self.b[t + 5], self.u16(t + 6)
        keysz = 8 + 8 * (rank + 1)
        p = t + 24
        for _ in range(n):
            csize, fmask = self.u32(p), self.u32(p + 4)
            off = [self.u64(p + 8 + 8 * i) for i in range(rank)]
            child = self.u64(p + keysz)
            p += keysz + 8
            if level:
                self._chunks(child, rank, dims, cd, es, filters, out, dt)
                continue
            raw = self.b[self.base + child: self.base + child + csize]
            for k in reversed(range(len(filters))):
                if fmask & (1 << k):
                    continue
                if filters[k] == 1:
                    raw = zlib.decompress(raw)
                elif filters[k] == 2:
                    a = np.frombuffer(raw, np.uint8)
                    raw = a.reshape(es, -1).T.tobytes()
                else:
                    raise ValueError("filter %d" % filters[k])
            chunk = np.frombuffer(raw, dt).reshape(cd)
            sl_out = tuple(slice(o, min(o + c, d)) for o, c, d in zip(off, cd, dims))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]
