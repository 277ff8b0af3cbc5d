"""Minimal fast5 / HDF5 chunk reader and transcoder: enough of the HDF5 file format to find every chunked dataset of
a fast5 file, its filter pipeline and its raw chunk payloads WITHOUT libhdf5, so that all VBZ-compressed reads of a
file (filter 32020, ref vbz_plugin/vbz_plugin.h:5) can be handed to the GPU as ONE batch instead of one filter
callback per chunk (ref vbz_plugin/vbz_plugin.cpp:97-229; python/fast5compress/fast5vbz.py:17-55 is the reference's
CPU-side walk over the same datasets).

Supported: superblock versions 0 and 1, version-1 object headers (with continuation blocks), old-style groups
(symbol-table message -> v1 B-tree + local heap + SNOD nodes), chunked data layout message version 3 (v1 chunk
B-tree), filter pipeline message versions 1 and 2.  This is what MinKNOW-era fast5 files and the reference's
test_data/*.fast5 use.  Anything else (superblock 2/3, "OHDR" object headers, fractal-heap groups, layout v4) raises
Fast5FormatError with the feature's name - nothing is guessed.

`transcode` is the writer half: what python/fast5compress/fast5vbz.py:17-55 does through h5py (copy the file, replace every
read's Raw/Signal by a single-chunk dataset with the other compression), done on the file's bytes - all reads of the file
go through the codec as ONE batch.

No sample arithmetic happens here: bytes are located, the codec runs in libvbz_b200 (csrc/)."""

import zlib

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

    def __init__(self, name, shape, itemsize, filters, chunk_shape, chunks, signed=True):
        self.name, self.shape, self.itemsize, self.filters, self.chunk_shape, self.chunks = name, shape, itemsize, filters, chunk_shape, chunks
        self.signed = signed

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
        self.sb_pos, self.sb_ver = base, ver
        if ver > 1:
            raise Fast5FormatError("superblock version %d (only 0 and 1 are supported)" % ver)
        self.so, self.sl = b[base + 13], b[base + 14]
        if self.so not in (4, 8) or self.sl not in (4, 8):
            raise Fast5FormatError("unsupported offset/length sizes")
        self.chunk_k = self._u(base + 24, 2) if ver == 1 else 32  # "indexed storage internal node K": a field from version 1 on
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_address = self._off(p)
        self.eof_pos = p + 2 * self.so  # where the end-of-file address is kept
        p += 4 * self.so  # base, free-space info, end of file, driver info
        # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
        self.root_header = self._off(p + self.so)

    # ---- object headers -------------------------------------------------------------------------------------------------
    def _messages(self, addr, extents=None):
        """[(type, data_start, size)] of a version-1 object header, continuation blocks included.  `extents` (a list)
        collects the (position, size) of the header's blocks; a message's flags byte sits at data_start - 4."""
        b = self.b
        addr += self.base_address
        if bytes(b[addr:addr + 4]) == b"OHDR":
            raise Fast5FormatError("version-2 object headers (OHDR)")
        if b[addr] != 1:
            raise Fast5FormatError("object header version %d" % b[addr])
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        if extents is not None:
            extents.append((addr, 16 + hsize))
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
                    if extents is not None:
                        extents.append(blocks[-1])
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
        """(name, object header address, position of that address in the file) of every link of an old-style group."""
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
                    out.append((self._name(heap_data, self._off(q)), self._off(q + self.so), q + self.so))
            else:
                raise Fast5FormatError("group node signature %r" % sig)
        return out

    # ---- datasets ---------------------------------------------------------------------------------------------------------
    def chunk_node_size(self, rank):
        """Bytes of a chunk B-tree node: libhdf5 always writes (and reads) room for 2K children and 2K + 1 keys."""
        return 8 + 2 * self.so + (2 * self.chunk_k + 1) * (8 + 8 * (rank + 1)) + 2 * self.chunk_k * self.so

    def _chunks(self, btree, rank, nodes=None):
        out = []
        if btree == _UNDEF or btree == (1 << (8 * self.so)) - 1:
            return out
        stack = [btree + self.base_address]
        while stack:
            a = stack.pop()
            if bytes(self.b[a:a + 4]) != b"TREE" or self.b[a + 4] != 1:
                raise Fast5FormatError("chunk B-tree node")
            if nodes is not None:
                nodes.append((a, self.chunk_node_size(rank)))
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
        shape, itemsize, filters, layout, signed = None, None, [], None, True
        for mtype, d, size in msgs:
            b = self.b
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = b[d], b[d + 1], b[d + 2]
                p = d + (8 if ver == 1 else 4)
                shape = tuple(self._len(p + k * self.sl) for k in range(rank))
            elif mtype == 0x0003:  # datatype
                itemsize = self._u(d + 4, 4)
                signed = bool(b[d + 1] & 8)  # fixed-point class: bit 3 of the class bit field
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
        return Dataset(name, shape, itemsize, filters, layout[2], self._chunks(layout[0], layout[1]), signed)

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
                for name, child, _ in self._group_entries(self._off(d), self._off(d + self.so)):
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

    # ---- attributes -----------------------------------------------------------------------------------------------------
    def _attribute_parts(self, d):
        """(name, datatype position, datatype size, dataspace position, dataspace size, data position) of an attribute message."""
        b = self.b
        ver = b[d]
        nlen, tlen, slen = self._u(d + 2, 2), self._u(d + 4, 2), self._u(d + 6, 2)
        if ver == 1:
            p = d + 8
            r8 = lambda v: (v + 7) // 8 * 8
        elif ver in (2, 3):
            if b[d + 1] & 3:
                raise Fast5FormatError("attribute with a shared datatype or dataspace")
            p = d + (9 if ver == 3 else 8)
            r8 = lambda v: v
        else:
            raise Fast5FormatError("attribute message version %d" % ver)
        name = bytes(b[p:p + nlen]).split(b"\0")[0].decode("utf-8", "replace")
        tp = p + r8(nlen)
        sp = tp + r8(tlen)
        return name, tp, tlen, sp, slen, sp + r8(slen)

    def _elements(self, sp):
        """Number of elements a dataspace message describes."""
        b = self.b
        ver, rank = b[sp], b[sp + 1]
        if ver == 2 and b[sp + 3] == 2:
            return 0
        p = sp + (8 if ver == 1 else 4)
        n = 1
        for k in range(rank):
            n *= self._len(p + k * self.sl)
        return n

    def _heap_object(self, collection, index):
        a = collection + self.base_address
        if bytes(self.b[a:a + 4]) != b"GCOL":
            raise Fast5FormatError("global heap collection signature")
        end = a + self._len(a + 8)
        p = a + 8 + self.sl
        while p + 8 + self.sl <= end:
            idx, size = self._u(p, 2), self._len(p + 8)
            if idx == index:
                return bytes(self.b[p + 8 + self.sl:p + 8 + self.sl + size])
            if idx == 0:
                break
            p += 8 + self.sl + (size + 7) // 8 * 8
        raise Fast5FormatError("global heap object %d" % index)

    def attributes(self, hdr):
        """{name: (datatype bytes, dataspace bytes, value)} of the object at `hdr`; value = the raw data, or for
        variable-length types the list of the elements' bytes fetched from the global heap."""
        out = {}
        for mtype, d, size in self._messages(hdr):
            if mtype != 0x000C:
                continue
            name, tp, tlen, sp, slen, dp = self._attribute_parts(d)
            esize = self._u(tp + 4, 4)
            n = self._elements(sp)
            if self.b[tp] & 15 == 9:
                value = []
                for k in range(n):
                    q = dp + k * (8 + self.so)
                    value.append(self._heap_object(self._off(q + 4), self._u(q + 4 + self.so, 4)) if self._u(q, 4) else b"")
            else:
                value = bytes(self.b[dp:dp + n * esize])
            out[name] = (bytes(self.b[tp:tp + tlen]), bytes(self.b[sp:sp + slen]), value)
        return out

    def objects(self):
        """{path: object header address} of every group and dataset reachable from the root."""
        out, stack = {}, [("", self.root_header)]
        while stack:
            path, hdr = stack.pop()
            if (path or "/") in out:
                continue
            out[path or "/"] = hdr
            for m in self._messages(hdr):
                if m[0] == 0x0011:
                    for name, child, _ in self._group_entries(self._off(m[1]), self._off(m[1] + self.so)):
                        stack.append((path + "/" + name, child))
        return out


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


# ---- the writer half: re-compress every read's Raw/Signal in place of the old one -----------------------------------------------

_DEFLATE, _SHUFFLE, _FLETCHER32 = 1, 2, 3


def is_read_signal(name):
    """The datasets python/fast5compress/fast5vbz.py:39-53 replaces: /read_*/Raw/Signal of a multi-read file."""
    parts = name.split("/")
    return len(parts) == 4 and parts[0] == "" and parts[1].startswith("read_") and parts[2] == "Raw" and parts[3] == "Signal"


class _Space:
    """The bytes of the file the replaced datasets gave up.  First fit at the lowest address, neighbours merge, what is
    free at the end of the file is cut off - the role libhdf5's free-space manager plays while the reference's script
    deletes and re-creates datasets inside one session."""

    def __init__(self, eof):
        self.free, self.eof = [], eof

    def release(self, start, size):
        if size <= 0:
            return
        merged = []
        for a, e in sorted(self.free + [[start, start + size]]):
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([a, e])
        self.free = merged

    def take(self, size):
        for i, (a, e) in enumerate(self.free):
            if e - a >= size:
                if e - a == size:
                    del self.free[i]
                else:
                    self.free[i][0] = a + size
                return a
        a = self.eof
        if self.free and self.free[-1][1] == self.eof:  # grow the free tail instead of leaving it behind as a hole
            a = self.free.pop()[0]
        self.eof = a + size
        return a

    def finish(self):
        while self.free and self.free[-1][1] == self.eof:
            self.eof = self.free.pop()[0]
        return self.eof


def _pad8(b):
    return b + bytes(-len(b) % 8)


def filter_pipeline_message(filters):
    """Body of a version-1 filter pipeline message, [(id, name, flags, cd_values)] -> bytes, as libhdf5 lays it out: the
    name NUL-padded to a multiple of eight (and its length field says the padded length), an odd number of cd_values
    padded by four bytes."""
    out = bytes([1, len(filters)]) + bytes(6)
    for fid, name, flags, cd in filters:
        nm = _pad8(name + b"\0") if name else b""
        out += fid.to_bytes(2, "little") + len(nm).to_bytes(2, "little") + flags.to_bytes(2, "little") + len(cd).to_bytes(2, "little") + nm
        out += b"".join(int(v).to_bytes(4, "little") for v in cd) + (bytes(4) if len(cd) % 2 else b"")
    return out


def vbz_pipeline(vbz_version, integer_size, zigzag, level):
    """What h5py writes for compression=32020, compression_opts=(version, size, zigzag, level): an optional filter (flag 1)
    named as the plugin registers itself (ref vbz_plugin/vbz_plugin.cpp:236)."""
    return filter_pipeline_message([(VBZ_FILTER_ID, b"vbz", 1, [vbz_version, integer_size, int(bool(zigzag)), level])])


def gzip_pipeline(level):
    return filter_pipeline_message([(_DEFLATE, b"deflate", 1, [level])])


def chunk_btree_node(f, entries, end_offset, itemsize):
    """A leaf node of a rank-1 chunk B-tree: entries [(stored size, filter mask, element offset, address)], then the key
    that closes the node (no size, offset = one past the last chunk, the element size in the extra dimension)."""
    undef = b"\xff" * f.so
    out = b"TREE" + bytes([1, 0]) + len(entries).to_bytes(2, "little") + undef + undef
    for size, mask, off, addr in entries:
        out += size.to_bytes(4, "little") + mask.to_bytes(4, "little") + off.to_bytes(8, "little") + bytes(8) + addr.to_bytes(f.so, "little")
    out += bytes(8) + end_offset.to_bytes(8, "little") + itemsize.to_bytes(8, "little")
    return out + bytes(f.chunk_node_size(1) - len(out))


def chunked_layout_message(f, btree, chunk_elems, itemsize):
    """Version-3 data layout message, chunked, rank 1 (the extra dimension is the element)."""
    return _pad8(bytes([3, 2, 2]) + btree.to_bytes(f.so, "little") + chunk_elems.to_bytes(4, "little") + itemsize.to_bytes(4, "little"))


def object_header(messages, refcount=1):
    """Version-1 object header in one block: [(type, flags, body)] with bodies already padded to eight."""
    body = b"".join(t.to_bytes(2, "little") + len(b).to_bytes(2, "little") + bytes([fl, 0, 0, 0]) + b for t, fl, b in messages)
    return bytes([1, 0]) + len(messages).to_bytes(2, "little") + refcount.to_bytes(4, "little") + len(body).to_bytes(4, "little") + bytes(4) + body


def _unshuffle(raw, itemsize):
    n = len(raw) // itemsize
    body = np.frombuffer(raw, np.uint8, n * itemsize).reshape(itemsize, n).T.tobytes()
    return body + raw[n * itemsize:]


def _unfilter(ds, chunk, payload):
    """Undo a pipeline of libhdf5's own filters on one chunk (last filter first; a set bit in the mask = filter skipped)."""
    raw = bytes(payload)
    for k in range(len(ds.filters) - 1, -1, -1):
        if chunk.filter_mask >> k & 1:
            continue
        fid = ds.filters[k][0]
        if fid == _DEFLATE:
            raw = zlib.decompress(raw)
        elif fid == _SHUFFLE:
            raw = _unshuffle(raw, ds.itemsize)
        elif fid == _FLETCHER32:
            raw = raw[:-4]
        else:
            raise Fast5FormatError("filter %d in the pipeline of %s" % (fid, ds.name))
    return raw


class _Target:
    """A dataset to replace: where its link points from, its header's messages and everything it owns in the file."""

    def __init__(self, ds, link_pos, refcount, messages, owned):
        self.ds, self.link_pos, self.refcount, self.messages, self.owned = ds, link_pos, refcount, messages, owned
        self.raw = None


def _targets(f, name_filter):
    out, seen = [], set()
    stack = [("", f.root_header, None)]
    while stack:
        path, hdr, link_pos = stack.pop()
        if hdr in seen:
            continue
        seen.add(hdr)
        extents = []
        msgs = f._messages(hdr, extents)
        sym = [m for m in msgs if m[0] == 0x0011]
        if sym:
            d = sym[0][1]
            for name, child, pos in f._group_entries(f._off(d), f._off(d + f.so)):
                stack.append((path + "/" + name, child, pos))
            continue
        if link_pos is None or not name_filter(path):
            continue
        ds = f._dataset(path, msgs)
        if ds is None:
            raise Fast5FormatError("%s is not a chunked dataset" % path)
        if len(ds.shape) != 1 or ds.itemsize not in (1, 2, 4):
            raise Fast5FormatError("%s: only one-dimensional integer datasets are re-compressed" % path)
        if any(f.b[d - 4] & 2 for _, d, _ in msgs):
            raise Fast5FormatError("%s: shared header messages" % path)
        owned = list(extents)
        layout = [m for m in msgs if m[0] == 0x0008][0][1]
        f._chunks(f._off(layout + 3), 1, owned)
        owned += [(c.address, c.size) for c in ds.chunks]
        keep = [(t, f.b[d - 4], bytes(f.b[d:d + size])) for t, d, size in msgs if t not in (0x0000, 0x0010)]
        out.append(_Target(ds, link_pos, f._u(hdr + f.base_address + 4, 4), keep, owned))
    out.sort(key=lambda t: t.ds.name)
    return out


def _sample_dtype(ds):
    return np.dtype({1: "i1", 2: "<i2", 4: "<i4"}[ds.itemsize] if ds.signed else {1: "u1", 2: "<u2", 4: "<u4"}[ds.itemsize])


def _read_raw(f, targets, batch):
    """The samples of every target: VBZ chunks of the whole file as one batch per cd_values, anything else by zlib/numpy."""
    groups = {}
    for t in targets:
        ds = t.ds
        t.raw = np.zeros(ds.shape[0], _sample_dtype(ds))
        cd = ds.vbz_cd_values
        for c in ds.chunks:
            n = min(ds.chunk_shape[0], ds.shape[0] - c.offset[0])
            if cd is not None and not c.filter_mask & 1:
                if c.size < 4:
                    raise Fast5FormatError("VBZ chunk shorter than its size header")
                stored = int.from_bytes(f.b[c.address:c.address + 4], "little")
                if stored < n * ds.itemsize:
                    raise Fast5FormatError("VBZ chunk of %s holds fewer bytes than the dataset needs" % ds.name)
                groups.setdefault(tuple(cd[:4] if len(cd) > 3 else list(cd[:3]) + [1]), []).append((t, c, n, stored))
            else:
                if any(fid == VBZ_FILTER_ID for fid, _ in ds.filters) and cd is None:
                    raise Fast5FormatError("VBZ next to other filters in the pipeline of %s" % ds.name)
                raw = f.chunk_bytes(c) if cd is not None else _unfilter(ds, c, f.chunk_bytes(c))
                t.raw[c.offset[0]:c.offset[0] + n] = np.frombuffer(raw, t.raw.dtype, n)
    for cd, work in groups.items():
        opts = options_from_cd_values(list(cd))
        res = batch.decompress([np.frombuffer(f.chunk_bytes(c), np.uint8) for _, c, _, _ in work], [st for _, _, _, st in work], opts, sized=True)
        for (t, c, n, stored), (code, blob) in zip(work, res):
            if code != stored:
                raise Fast5FormatError("chunk of %s failed to decode: 0x%08X" % (t.ds.name, code))
            t.raw[c.offset[0]:c.offset[0] + n] = np.frombuffer(blob, t.raw.dtype, n)


def _datatype_span(b, p):
    """(holds_addresses, bytes) of the datatype message that starts at b[p]: variable-length and reference types store file addresses
    (global heap IDs, object addresses) in the data; compound, array, enumeration and variable-length types nest other datatypes."""
    u = lambda q, n: int.from_bytes(b[q:q + n], "little")
    cls, ver = b[p] & 15, b[p] >> 4
    bits = u(p + 1, 3)
    size = u(p + 4, 4)
    q = p + 8
    if cls in (0, 4):
        return False, 12
    if cls == 1:
        return False, 20
    if cls == 2:
        return False, 10
    if cls == 3:
        return False, 8
    if cls == 5:
        return False, 8 + ((bits & 0xFF) + 7) // 8 * 8
    if cls == 7:
        return True, 8
    if cls == 9:
        return True, 8 + _datatype_span(b, q)[1]
    if cls == 6:
        holds = False
        for _ in range(bits & 0xFFFF):
            e = q
            while b[e] != 0:
                e += 1
            q = q + (e - q) // 8 * 8 + 8 if ver < 3 else e + 1  # the name, NUL-padded to a multiple of eight before version 3
            if ver == 1:
                q += 4 + 1 + 3 + 4 + 4 + 16  # offset, dimensionality, reserved, permutation, reserved, four dimension sizes
            elif ver == 2:
                q += 4
            else:
                q += max(1, (max(size, 1).bit_length() + 7) // 8)  # the offset takes as many bytes as the compound's size needs
            h, n = _datatype_span(b, q)
            holds = holds or h
            q += n
        return holds, q - p
    if cls == 8:
        h, n = _datatype_span(b, q)
        q += n
        for _ in range(bits & 0xFFFF):
            e = q
            while b[e] != 0:
                e += 1
            q = q + (e - q) // 8 * 8 + 8 if ver < 3 else e + 1
        return h, q - p + (bits & 0xFFFF) * u(p + 8 + 4, 4)
    if cls == 10:
        rank = b[q]
        q += 4 + 8 * rank if ver < 3 else 1 + 4 * rank
        h, n = _datatype_span(b, q)
        return h, q - p + n
    raise Fast5FormatError("datatype class %d" % cls)


def _address_sites(f):
    """Position of every file address stored in the metadata reachable from the superblock (undefined addresses left
    out).  Anything that may hold an address this walk does not understand raises Fast5FormatError - `compact` must never
    move a byte something still points at by an address it has not found."""
    b, so, sl = f.b, f.so, f.sl
    undef = (1 << (8 * so)) - 1
    sites = []

    def site(pos):
        v = f._off(pos)
        if v == undef:
            return None
        sites.append(pos)
        return v

    def entry(q):  # symbol table entry: name offset, header address, cache type, reserved, scratch
        site(q + so)
        if f._u(q + 2 * so, 4) == 1:  # the group's B-tree and heap addresses cached in the scratch space
            site(q + 2 * so + 8)
            site(q + 2 * so + 8 + so)
        return f._off(q + so)

    def chunk_tree(addr, rank):
        stack = [addr + f.base_address]
        while stack:
            a = stack.pop()
            if bytes(b[a:a + 4]) != b"TREE" or b[a + 4] != 1:
                raise Fast5FormatError("chunk B-tree node")
            site(a + 8)
            site(a + 8 + so)
            key = 8 + 8 * (rank + 1)
            p = a + 8 + 2 * so
            for _ in range(f._u(a + 6, 2)):
                child = site(p + key)
                if b[a + 5] and child is not None:
                    stack.append(child + f.base_address)
                p += key + so

    def group(btree, heap):
        h = heap + f.base_address
        if bytes(b[h:h + 4]) != b"HEAP":
            raise Fast5FormatError("local heap signature")
        site(h + 8 + 2 * sl)
        children = []
        stack = [btree + f.base_address]
        while stack:
            a = stack.pop()
            sig = bytes(b[a:a + 4])
            if sig == b"TREE":
                if b[a + 4] != 0:
                    raise Fast5FormatError("group B-tree node type")
                site(a + 8)
                site(a + 8 + so)
                p = a + 8 + 2 * so
                for _ in range(f._u(a + 6, 2)):
                    p += sl
                    stack.append(site(p) + f.base_address)
                    p += so
            elif sig == b"SNOD":
                for i in range(f._u(a + 6, 2)):
                    q = a + 8 + i * (2 * so + 24)
                    if f._u(q + 2 * so, 4) == 2:  # a symbolic link: no object behind it
                        continue
                    children.append(entry(q))
            else:
                raise Fast5FormatError("group node signature %r" % sig)
        return children

    def attribute(d):
        name, tp, tlen, sp, slen, dp = f._attribute_parts(d)
        cls = b[tp] & 15
        if cls != 9 and _datatype_span(b, tp)[0]:
            raise Fast5FormatError("attribute %s: its type holds file addresses inside a compound, array or reference" % name)
        if cls == 9:
            if _datatype_span(b, tp + 8)[0]:
                raise Fast5FormatError("attribute %s: nested variable-length type" % name)
            for k in range(f._elements(sp)):
                q = dp + k * (8 + so)
                if f._u(q, 4):
                    site(q + 4)

    p = f.sb_pos + 24 + (4 if f.sb_ver == 1 else 0)
    site(p + so)       # free-space info
    site(p + 3 * so)   # driver information block
    seen, stack = set(), [entry(p + 4 * so)]
    while stack:
        hdr = stack.pop()
        if hdr in seen:
            continue
        seen.add(hdr)
        msgs = f._messages(hdr)
        if any(m[0] == 0x0008 for m in msgs):  # a dataset: its elements must not be addresses (variable-length data, references)
            for mtype, d, size in msgs:
                if mtype == 0x0003 and not b[d - 4] & 2 and _datatype_span(b, d)[0]:
                    raise Fast5FormatError("a dataset of variable-length or reference type stores file addresses in its data")
        for mtype, d, size in msgs:
            if mtype != 0 and b[d - 4] & 2:
                raise Fast5FormatError("shared header message")
            if mtype == 0x0010:
                site(d)
            elif mtype == 0x0011:
                stack += group(site(d), site(d + so))
            elif mtype == 0x0008:
                if b[d] != 3:
                    raise Fast5FormatError("data layout message version %d" % b[d])
                if b[d + 1] == 2:
                    t = site(d + 3)
                    if t is not None:
                        chunk_tree(t, b[d + 2] - 1)
                elif b[d + 1] == 1:
                    site(d + 2)
            elif mtype == 0x000C:
                attribute(d)
            elif mtype in (0x0002, 0x0006, 0x0007, 0x000F, 0x0014, 0x0015):
                raise Fast5FormatError("header message 0x%04X holds addresses this walk does not follow" % mtype)
    return sites


def compact(data, holes):
    """Cut the byte ranges `holes` [(start, end)] out of a file and re-aim every address at where its target went."""
    f = Fast5File(memoryview(bytes(data)))
    holes = sorted((a, e) for a, e in holes if e > a)
    starts = [a for a, _ in holes]
    gone = [0]
    for a, e in holes:
        gone.append(gone[-1] + e - a)

    def moved(rel):
        pos = rel + f.base_address
        k = int(np.searchsorted(starts, pos, side="right"))
        if k and pos < holes[k - 1][1]:
            raise Fast5FormatError("an address points into space that was given up")
        return rel - gone[k]

    out = bytearray(f.b)
    for pos in _address_sites(f):
        out[pos:pos + f.so] = moved(f._off(pos)).to_bytes(f.so, "little")
    eof = len(out) - f.base_address
    k = len(holes)
    out[f.eof_pos:f.eof_pos + f.so] = (eof - gone[k]).to_bytes(f.so, "little")
    for a, e in reversed(holes):
        del out[a:e]
    return bytes(out)


def transcode(data, batch, decompress=False, vbz_version=1, zstd_level=1, name_filter=is_read_signal, compact_file=True, chunk_elems=None):
    """bytes of a fast5 file -> bytes of the same file with every read's Raw/Signal stored as ONE chunk, VBZ-compressed
    (cd_values (vbz_version, element size, zigzag for signed samples, zstd_level)) or, with `decompress`, gzip level 1 -
    the two directions of python/fast5compress/fast5vbz.py:17-55.  All reads are decoded as one batch and encoded as one
    batch on `batch` (api.Batch).  Everything else in the file stays where it is, byte for byte: a replaced dataset gets a
    new object header, chunk index and chunk in the space the old ones gave up (or behind the end of the file), and the
    one link that pointed at it is re-aimed.  With `compact_file` the space that stays unused is then cut out and every
    address in the file's metadata moved along (`compact`), so the file shrinks by what the new compression saves - a file
    with structures `compact` does not know raises Fast5FormatError; pass compact_file=False to keep such a file's holes
    (what the reference's script leaves too).  Empty datasets are left alone (h5py cannot chunk them either).
    `chunk_elems` cuts every dataset into chunks of that many samples instead of one (the last chunk is stored whole, padded with
    zeros, as HDF5 stores edge chunks); up to 2K = 64 chunks fit the single B-tree node written here."""
    f = Fast5File(memoryview(bytes(data)))
    targets = [t for t in _targets(f, name_filter) if t.ds.shape[0] > 0]
    _read_raw(f, targets, batch)
    pieces = []  # (target index, first sample, samples of one chunk)
    for i, t in enumerate(targets):
        n = t.ds.shape[0]
        c = n if chunk_elems is None else int(chunk_elems)
        if c <= 0 or (n + c - 1) // c > 2 * f.chunk_k:
            raise Fast5FormatError("%s: %d samples do not fit %d chunks of %d" % (t.ds.name, n, 2 * f.chunk_k, c))
        t.chunk_len = c
        for a in range(0, n, c):
            part = t.raw[a:a + c]
            if part.size < c:
                part = np.concatenate([part, np.zeros(c - part.size, part.dtype)])
            pieces.append((i, a, part))
    payloads = [None] * len(pieces)
    if decompress:
        payloads = [zlib.compress(part.tobytes(), 1) for _, _, part in pieces]
    else:
        groups = {}
        for k, (i, _, _) in enumerate(pieces):
            groups.setdefault((targets[i].ds.itemsize, targets[i].ds.signed), []).append(k)
        for (itemsize, signed), idx in groups.items():
            opts = CompressionOptions(bool(signed), itemsize, int(zstd_level), int(vbz_version))
            for k, (code, blob) in zip(idx, batch.compress([pieces[k][2] for k in idx], opts, sized=True)):
                if code >= 0xFFFFFFF9 or code != len(blob):
                    raise Fast5FormatError("%s failed to compress: 0x%08X" % (targets[pieces[k][0]].ds.name, code))
                payloads[k] = blob
    out = bytearray(f.b)
    space = _Space(len(out))
    for t in targets:
        for pos, size in t.owned:
            out[pos:pos + size] = bytes(size)
            space.release(pos, size)

    def put(blob):
        pos = space.take(len(blob))
        if pos + len(blob) > len(out):
            out.extend(bytes(pos + len(blob) - len(out)))
        out[pos:pos + len(blob)] = blob
        return pos - f.base_address

    for t in targets:
        t.entries = []
    for (i, a, _), payload in zip(pieces, payloads):  # big pieces first, so that they get the low, big holes
        targets[i].entries.append((len(payload), 0, a, put(payload)))
    for t in targets:
        ds = t.ds
        n = t.chunk_len
        tree = put(chunk_btree_node(f, t.entries, len(t.entries) * n, ds.itemsize))
        pipeline = gzip_pipeline(1) if decompress else vbz_pipeline(vbz_version, ds.itemsize, ds.signed, zstd_level)
        msgs, placed = [], False
        for mtype, flags, body in t.messages:
            if mtype == 0x000B:
                continue
            if mtype == 0x0008:
                msgs += [(0x000B, 1, pipeline), (0x0008, 0, chunked_layout_message(f, tree, n, ds.itemsize))]
                placed = True
            else:
                msgs.append((mtype, flags, _pad8(body)))
        assert placed
        hdr = put(object_header(msgs, t.refcount))
        out[t.link_pos:t.link_pos + f.so] = hdr.to_bytes(f.so, "little")
    eof = space.finish()
    del out[eof:]
    out[f.eof_pos:f.eof_pos + f.so] = (eof - f.base_address).to_bytes(f.so, "little")
    if compact_file and space.free:
        return compact(out, [(a, e) for a, e in space.free])
    return bytes(out)


def compress_fast5(filename, output_suffix, batch, vbz_version=0, decompress=False, compact_file=True, chunk_elems=None):
    """python/fast5compress/fast5vbz.py:17-55 with a batch handle in place of h5py: writes filename + output_suffix and
    returns its name."""
    import os
    filename = os.path.abspath(filename)
    out_filename = filename + output_suffix
    with open(filename, "rb") as fh:
        data = fh.read()
    result = transcode(data, batch, decompress=decompress, vbz_version=vbz_version, compact_file=compact_file, chunk_elems=chunk_elems)
    with open(out_filename, "wb") as fh:
        fh.write(result)
    return out_filename
