"""Minimal fast5 / HDF5 chunk reader: enough of the HDF5 file format to find every chunked dataset of
a fast5 file, its filter pipeline and its raw chunk payloads WITHOUT libhdf5, so that all VBZ-compressed reads of a
file (filter 32020, ref vbz_plugin/vbz_plugin.h:5) can be handed to the GPU as ONE batch instead of one filter
callback per chunk (ref vbz_plugin/vbz_plugin.cpp:97-229; python/fast5compress/fast5vbz.py:17-55 is the reference's
CPU-side walk over the same datasets).

Supported: superblock versions 0 and 1, version-1 object headers (with continuation blocks), old-style groups
(symbol-table message -> v1 B-tree + local heap + SNOD nodes), chunked data layout message version 3 (v1 chunk
B-tree), filter pipeline message versions 1 and 2.  This is what MinKNOW-era fast5 files and the reference's
test_data/*.fast5 use.  Anything else (superblock 2/3, "OHDR" object headers, fractal-heap groups, layout v4) raises
Fast5FormatError with the feature's name - nothing is guessed.

No sample arithmetic happens here: bytes are located, the codec runs in libvbz_b200 (csrc/)."""

import numpy as np

from ._lib import CompressionOptions

VBZ_FILTER_ID = 32020
_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Fast5FormatError(Exception):
    pass


class Chunk:
    __slots__ = ("offset", "filter_mask", "address", "size")

    def __init__(self, offset, filter_mask, address, size):
        self.offset, self.filter_mask, self.address, self.size = offset, filter_mask, address, size


class Dataset:
    """A chunked dataset: name (path inside the file), shape, element size, filters [(id, [cd_values])], chunks."""

    def __init__(self, name, shape, itemsize, filters, chunk_shape, chunks):
        self.name, self.shape, self.itemsize, self.filters, self.chunk_shape, self.chunks = name, shape, itemsize, filters, chunk_shape, chunks

    @property
    def vbz_cd_values(self):
        """cd_values of the VBZ filter if it is the dataset's only filter, else None."""
        if len(self.filters) == 1 and self.filters[0][0] == VBZ_FILTER_ID:
            return self.filters[0][1]
        return None


def options_from_cd_values(cd):
    """cd_values = {version, integer_size, zigzag, level}; level defaults to 1 with three values, extra values are
    ignored, fewer than three are an error (ref vbz_plugin/vbz_plugin.cpp:112-140)."""
    if cd is None or len(cd) < 3:
        raise Fast5FormatError("VBZ filter with fewer than 3 cd_values")
    level = cd[3] if len(cd) > 3 else 1
    return CompressionOptions(bool(cd[2]), int(cd[1]), int(level), int(cd[0]))


class Fast5File:
    def __init__(self, data):
        self.b = memoryview(data) if not isinstance(data, memoryview) else data
        self._parse_superblock()

    @classmethod
    def open(cls, path):
        return cls(np.fromfile(path, dtype=np.uint8).data)

    # ---- primitives -------------------------------------------------------------------------------------------------------
    def _u(self, pos, n):
        return int.from_bytes(self.b[pos:pos + n], "little")

    def _off(self, pos):
        return self._u(pos, self.so)

    def _len(self, pos):
        return self._u(pos, self.sl)

    def _parse_superblock(self):
        b = self.b
        base = -1
        for cand in (0, 512, 1024, 2048, 4096):  # the superblock sits at 0 or at a power of two >= 512
            if bytes(b[cand:cand + 8]) == _SIG:
                base = cand
                break
        if base < 0:
            raise Fast5FormatError("not an HDF5 file")
        ver = b[base + 8]
        if ver > 1:
            raise Fast5FormatError("superblock version %d (only 0 and 1 are supported)" % ver)
        self.so, self.sl = b[base + 13], b[base + 14]
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise Fast5FormatError("unsupported offset/length sizes")
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_address = self._off(p)
        p += 4 * self.so  # base, free-space info, end of file, driver info
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
        self.root_header = self._off(p + self.so)

    # ---- object headers -------------------------------------------------------------------------------------------------
    def _messages(self, addr):
        """[(type, data_start, size)] of a version-1 object header, continuation blocks included."""
        b = self.b
        addr += self.base_address
        if bytes(b[addr:addr + 4]) == b"OHDR":
            raise Fast5FormatError("version-2 object headers (OHDR)")
        if b[addr] != 1:
            raise Fast5FormatError("object header version %d" % b[addr])
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                data = p + 8
                out.append((mtype, data, msize))
                if mtype == 0x0010:  # continuation
                    blocks.append((self._off(data) + self.base_address, self._len(data + self.so)))
                p = data + msize
        return out

    # ---- groups -------------------------------------------------------------------------------------------------------------
    def _heap_data(self, addr):
        addr += self.base_address
        if bytes(self.b[addr:addr + 4]) != b"HEAP":
            raise Fast5FormatError("local heap signature")
        return self._off(addr + 8 + 2 * self.sl) + self.base_address

    def _name(self, heap_data, off):
        p = heap_data + off
        e = p
        while self.b[e] != 0:
            e += 1
        return bytes(self.b[p:e]).decode("utf-8", "replace")

    def _group_entries(self, btree, heap):
        """(name, object header address) of every link of an old-style group."""
        heap_data = self._heap_data(heap)
        out = []
        stack = [btree + self.base_address]
        while stack:
            a = stack.pop()
            sig = bytes(self.b[a:a + 4])
            if sig == b"TREE":
                if self.b[a + 4] != 0:
                    raise Fast5FormatError("group B-tree node type")
                used = self._u(a + 6, 2)
                p = a + 8 + 2 * self.so
                for _ in range(used):
                    p += self.sl  # key
                    stack.append(self._off(p) + self.base_address)
                    p += self.so
            elif sig == b"SNOD":
                n = self._u(a + 6, 2)
                p = a + 8
                esz = 2 * self.so + 4 + 4 + 16
                for i in range(n):
                    q = p + i * esz
                    out.append((self._name(heap_data, self._off(q)), self._off(q + self.so)))
            else:
                raise Fast5FormatError("group node signature %r" % sig)
        return out

    # ---- datasets ---------------------------------------------------------------------------------------------------------
    def _chunks(self, btree, rank):
        out = []
        if btree == _UNDEF or btree == (1 << (8 * self.so)) - 1:
            return out
        stack = [btree + self.base_address]
        while stack:
            a = stack.pop()
            if bytes(self.b[a:a + 4]) != b"TREE" or self.b[a + 4] != 1:
                raise Fast5FormatError("chunk B-tree node")
            level = self.b[a + 5]
            used = self._u(a + 6, 2)
            p = a + 8 + 2 * self.so
            key = 8 + 8 * (rank + 1)
            for _ in range(used):
                size, mask = self._u(p, 4), self._u(p + 4, 4)
                offs = tuple(self._u(p + 8 + 8 * k, 8) for k in range(rank))
                child = self._off(p + key)
                if level == 0:
                    out.append(Chunk(offs, mask, child + self.base_address, size))
                else:
                    stack.append(child + self.base_address)
                p += key + self.so
        out.sort(key=lambda c: c.offset)
        return out

    def _dataset(self, name, msgs):
        shape, itemsize, filters, layout = None, None, [], None
        for mtype, d, size in msgs:
            b = self.b
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = b[d], b[d + 1], b[d + 2]
                p = d + (8 if ver == 1 else 4)
                shape = tuple(self._len(p + k * self.sl) for k in range(rank))
            elif mtype == 0x0003:  # datatype
                itemsize = self._u(d + 4, 4)
            elif mtype == 0x000B:  # filter pipeline
                ver, nf = b[d], b[d + 1]
                p = d + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self._u(p, 2)
                    p += 2
                    namelen = 0
                    if ver == 1 or fid >= 256:
                        namelen = self._u(p, 2)
                        p += 2
                    p += 2  # flags
                    ncd = self._u(p, 2)
                    p += 2
                    p += (namelen + 7) // 8 * 8 if ver == 1 else namelen
                    cd = [self._u(p + 4 * k, 4) for k in range(ncd)]
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
            elif mtype == 0x0008:  # data layout
                ver = b[d]
                if ver != 3:
                    raise Fast5FormatError("data layout message version %d" % ver)
                if b[d + 1] == 2:  # chunked
                    dims = b[d + 2]
                    btree = self._off(d + 3)
                    cshape = tuple(self._u(d + 3 + self.so + 4 * k, 4) for k in range(dims))
                    layout = (btree, dims - 1, cshape[:-1])
        if layout is None or shape is None:
            return None
        return Dataset(name, shape, itemsize, filters, layout[2], self._chunks(layout[0], layout[1]))

    def datasets(self):
        """Every chunked dataset of the file, depth first from the root group."""
        out = []
        seen = set()
        stack = [("", self.root_header)]
        while stack:
            path, hdr = stack.pop()
            if hdr in seen:
                continue
            seen.add(hdr)
            msgs = self._messages(hdr)
            sym = [m for m in msgs if m[0] == 0x0011]
            if sym:
                d = sym[0][1]
                for name, child in self._group_entries(self._off(d), self._off(d + self.so)):
                    stack.append((path + "/" + name, child))
                continue
            if any(m[0] == 0x0002 for m in msgs):
                raise Fast5FormatError("new-style groups (link info message)")
            ds = self._dataset(path, msgs)
            if ds is not None:
                out.append(ds)
        out.sort(key=lambda d: d.name)
        return out

    def chunk_bytes(self, chunk):
        return self.b[chunk.address:chunk.address + chunk.size]


def read_vbz_signals(path_or_bytes, batch, name_filter=lambda name: name.endswith("/Signal")):
    """Decode every VBZ-compressed 1-D dataset of a fast5 file as ONE batch on `batch` (api.Batch).
    -> {dataset name: numpy array}.  Datasets with other filters are left alone."""
    f = Fast5File.open(path_or_bytes) if isinstance(path_or_bytes, str) else Fast5File(path_or_bytes)
    groups = {}
    for ds in f.datasets():
        cd = ds.vbz_cd_values
        if cd is None or not name_filter(ds.name) or len(ds.shape) != 1:
            continue
        groups.setdefault(tuple(cd[:4] if len(cd) > 3 else list(cd[:3]) + [1]), []).append(ds)
    out = {}
    for cd, dss in groups.items():
        opts = options_from_cd_values(list(cd))
        work = []
        for ds in dss:
            for c in ds.chunks:
                if c.filter_mask & 1:  # the filter was skipped for this chunk: stored raw
                    raise Fast5FormatError("unfiltered chunk in a VBZ dataset")
                n = min(ds.chunk_shape[0], ds.shape[0] - c.offset[0])  # elements of the chunk that lie inside the dataset
                if c.size < 4:
                    raise Fast5FormatError("VBZ chunk shorter than its size header")
                stored = int.from_bytes(f.b[c.address:c.address + 4], "little")  # HDF5 stores edge chunks whole
                if stored < n * ds.itemsize:
                    raise Fast5FormatError("VBZ chunk of %s holds fewer bytes than the dataset needs" % ds.name)
                work.append((ds, c, n, stored))
        res = batch.decompress([np.frombuffer(f.chunk_bytes(c), np.uint8) for _, c, _, _ in work], [st for _, _, _, st in work], opts, sized=True)
        dtype = {1: np.int8, 2: np.int16, 4: np.int32}[opts.integer_size] if opts.perform_delta_zig_zag else {1: np.uint8, 2: np.uint16, 4: np.uint32}[opts.integer_size]
        for (ds, c, n, stored), (code, blob) in zip(work, res):
            if code != stored:
                raise Fast5FormatError("chunk of %s failed to decode: 0x%08X" % (ds.name, code))
            arr = out.setdefault(ds.name, np.empty(ds.shape[0], dtype))
            arr[c.offset[0]:c.offset[0] + n] = np.frombuffer(blob, dtype)[:n]
    return out
