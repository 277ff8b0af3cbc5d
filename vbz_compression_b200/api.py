"""Python mirror of the reference's pyvbz module (python/pyvbz/vbz/__init__.py:12-76): same function
names, argument meaning and error behaviour, running on libvbz_b200.so; plus `Batch`, a thin wrapper
over the batch C-ABI (include/vbz_b200.h) for many reads per launch."""
import ctypes

import numpy as np

from . import _lib
from ._lib import CompressionOptions


class VbzError(Exception):
    pass


def compression_options(zigzag, integer_size, zstd_level=1, version=0):
    """ref python/pyvbz/vbz/__init__.py:12-18"""
    return CompressionOptions(bool(zigzag), int(integer_size), int(zstd_level), int(version))


def _default_options(dtype):
    dtype = np.dtype(dtype)
    return compression_options(dtype.kind == "i", dtype.itemsize, 1, 0)


def _check(lib, ret):
    if lib.vbz_is_error(ret):
        raise VbzError("VBZ error: " + lib.vbz_error_string(ret).decode())
    return ret


def compress(data, options=None):
    """ref python/pyvbz/vbz/__init__.py:21-44: numpy array in, numpy uint8 array (sized format) out."""
    lib = _lib.load()
    data = np.ascontiguousarray(data)
    if options is None:
        options = _default_options(data.dtype)
    raw = data.view(np.uint8).reshape(-1)
    cap = _check(lib, lib.vbz_max_compressed_size(raw.size, ctypes.byref(options)))
    out = np.empty(cap, np.uint8)
    n = _check(lib, lib.vbz_compress_sized(raw.ctypes.data, raw.size, out.ctypes.data, cap, ctypes.byref(options)))
    return out[:n]


def decompress(data, dtype, options=None):
    """ref python/pyvbz/vbz/__init__.py:47-76: sized-format bytes in, numpy array of `dtype` out."""
    lib = _lib.load()
    if options is None:
        options = _default_options(dtype)
    src = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    size = _check(lib, lib.vbz_decompressed_size(src.ctypes.data, src.size, ctypes.byref(options)))
    out = np.empty(size, np.uint8)
    n = _check(lib, lib.vbz_decompress_sized(src.ctypes.data, src.size, out.ctypes.data, size, ctypes.byref(options)))
    return out[:n].view(dtype)


class Batch:
    """Many reads per launch through the batch C-ABI.  Owns one vbz_b200_ctx (one device)."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or _lib.load()
        self.ctx = self.lib.vbz_b200_create(device)
        if not self.ctx:
            raise VbzError("vbz_b200_create failed: " + self.lib.vbz_b200_last_error().decode() + " (no CPU fallback)")
        self.device = device

    def close(self):
        if self.ctx:
            self.lib.vbz_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fail(self, what):
        raise VbzError(what + ": " + self.lib.vbz_b200_last_error().decode())

    @property
    def kernel_launches(self):
        return int(self.lib.vbz_b200_kernel_launches(self.ctx))

    # ---- host-pointer batches: lists of numpy arrays in, list of (code, bytes) out ------------------
    def compress(self, reads, options, sized=True):
        n = len(reads)
        raws = [np.ascontiguousarray(r).view(np.uint8).reshape(-1) for r in reads]
        sizes = (ctypes.c_uint32 * n)(*[r.size for r in raws])
        caps_l = [self.lib.vbz_max_compressed_size(r.size, ctypes.byref(options)) for r in raws]
        outs = [np.empty(c if c < 0xFFFFFFF9 else 4, np.uint8) for c in caps_l]
        caps = (ctypes.c_uint32 * n)(*[o.size for o in outs])
        srcp = (ctypes.c_void_p * n)(*[r.ctypes.data for r in raws])
        dstp = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
        res = (ctypes.c_uint32 * n)()
        if self.lib.vbz_b200_compress_batch(self.ctx, srcp, sizes, dstp, caps, res, n, ctypes.byref(options), int(sized)) != 0:
            self._fail("vbz_b200_compress_batch")
        return [(int(res[i]), outs[i][:res[i]].tobytes() if res[i] < 0xFFFFFFF9 and res[i] <= outs[i].size else b"") for i in range(n)]

    def decompress(self, chunks, sizes_out, options, sized=True):
        n = len(chunks)
        srcs = [np.frombuffer(bytes(c), np.uint8) if not isinstance(c, np.ndarray) else c for c in chunks]
        sizes = (ctypes.c_uint32 * n)(*[s.size for s in srcs])
        outs = [np.empty(max(int(s), 1), np.uint8) for s in sizes_out]
        dsz = (ctypes.c_uint32 * n)(*[int(s) for s in sizes_out])
        srcp = (ctypes.c_void_p * n)(*[s.ctypes.data if s.size else None for s in srcs])
        dstp = (ctypes.c_void_p * n)(*[o.ctypes.data for o in outs])
        res = (ctypes.c_uint32 * n)()
        if self.lib.vbz_b200_decompress_batch(self.ctx, srcp, sizes, dstp, dsz, res, n, ctypes.byref(options), int(sized)) != 0:
            self._fail("vbz_b200_decompress_batch")
        return [(int(res[i]), outs[i][:res[i]].tobytes() if res[i] < 0xFFFFFFF9 and res[i] <= outs[i].size else b"") for i in range(n)]

    # ---- device-resident batches: raw device pointers (e.g. torch tensor .data_ptr()) ------------------
    def compress_device(self, d_src, src_off, src_size, d_dst, dst_off, dst_cap, d_results, options, sized=True, stream=0):
        n = len(src_size)
        so = np.ascontiguousarray(src_off, np.uint64); ss = np.ascontiguousarray(src_size, np.uint32)
        do = np.ascontiguousarray(dst_off, np.uint64); dc = np.ascontiguousarray(dst_cap, np.uint32)
        u64p, u32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
        rc = self.lib.vbz_b200_compress_device(self.ctx, d_src, so.ctypes.data_as(u64p), ss.ctypes.data_as(u32p), d_dst,
                                               do.ctypes.data_as(u64p), dc.ctypes.data_as(u32p), d_results, n, ctypes.byref(options),
                                               int(sized), stream)
        if rc != 0:
            self._fail("vbz_b200_compress_device")

    def decompress_device(self, d_src, src_off, src_size, d_dst, dst_off, dst_size, d_results, options, sized=True, stream=0):
        n = len(src_size)
        so = np.ascontiguousarray(src_off, np.uint64); ss = np.ascontiguousarray(src_size, np.uint32)
        do = np.ascontiguousarray(dst_off, np.uint64); dc = np.ascontiguousarray(dst_size, np.uint32)
        u64p, u32p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
        rc = self.lib.vbz_b200_decompress_device(self.ctx, d_src, so.ctypes.data_as(u64p), ss.ctypes.data_as(u32p), d_dst,
                                                 do.ctypes.data_as(u64p), dc.ctypes.data_as(u32p), d_results, n, ctypes.byref(options),
                                                 int(sized), stream)
        if rc != 0:
            self._fail("vbz_b200_decompress_device")

    # ---- torch tensors that already live on the device: zero-copy ------------------------------------
    def compress_tensors(self, reads, options, sized=True, stream=None):
        """reads: list of 1-D contiguous CUDA tensors on this context's device (any integer dtype matching
        options.integer_size).  Nothing is copied: the kernels read the tensors where they are.  Returns
        (out, offsets, sizes): `out` one uint8 CUDA tensor, read i's bytes at out[offsets[i] : offsets[i] + sizes[i]],
        `sizes` an int32 CUDA tensor holding per read what vbz_compress[_sized] would have returned (view as uint32:
        values >= 0xFFFFFFF9 are VBZ error codes).  Enqueued on the current torch stream; not synchronised."""
        import torch
        n = len(reads)
        dev = torch.device("cuda", self.device)
        if n == 0:
            return torch.empty(0, dtype=torch.uint8, device=dev), np.zeros(0, np.uint64), torch.empty(0, dtype=torch.int32, device=dev)
        for t in reads:
            if not (t.is_cuda and t.device.index == self.device and t.is_contiguous() and t.data_ptr() % 16 == 0):
                raise VbzError("compress_tensors needs contiguous, 16-byte aligned CUDA tensors on device %d" % self.device)
        ptrs = np.array([t.data_ptr() for t in reads], np.uint64)
        base = int(ptrs.min())
        src_off = ptrs - np.uint64(base)
        src_size = np.array([t.numel() * t.element_size() for t in reads], np.uint32)
        caps = np.array([self.lib.vbz_b200_compress_bound(int(z), ctypes.byref(options), int(sized)) for z in src_size], np.uint64)
        dst_off = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
        out = torch.empty(int(caps.sum()) + 16, dtype=torch.uint8, device=dev)
        sizes = torch.zeros(n, dtype=torch.int32, device=dev)
        st = (stream or torch.cuda.current_stream(dev)).cuda_stream
        self.compress_device(base, src_off, src_size, out.data_ptr(), dst_off, caps.astype(np.uint32), sizes.data_ptr(), options, sized, st)
        return out, dst_off, sizes

    def decompress_tensors(self, chunks, out_dtype, out_lengths, options, sized=True, stream=None):
        """chunks: list of 1-D contiguous uint8 CUDA tensors (4-byte aligned) each holding one compressed read (e.g. slices
        of compress_tensors' output); out_lengths: elements of `out_dtype` per read.  Returns (list of CUDA tensors,
        results int32 CUDA tensor); enqueued on the current torch stream, not synchronised."""
        import torch
        n = len(chunks)
        dev = torch.device("cuda", self.device)
        esz = torch.empty(0, dtype=out_dtype).element_size()
        res = torch.zeros(n, dtype=torch.int32, device=dev)
        if n == 0:
            return [], res
        for t in chunks:
            if not (t.is_cuda and t.device.index == self.device and t.is_contiguous() and t.data_ptr() % 4 == 0):
                raise VbzError("decompress_tensors needs contiguous, 4-byte aligned CUDA tensors on device %d" % self.device)
        ptrs = np.array([t.data_ptr() for t in chunks], np.uint64)
        base = int(ptrs.min())
        src_off = ptrs - np.uint64(base)
        src_size = np.array([t.numel() for t in chunks], np.uint32)
        dst_size = np.array([int(x) * esz for x in out_lengths], np.uint32)
        slot = (dst_size.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15)
        dst_off = np.concatenate([[0], np.cumsum(slot)[:-1]]).astype(np.uint64)
        out = torch.empty(int(slot.sum()) + 16, dtype=torch.uint8, device=dev)
        st = (stream or torch.cuda.current_stream(dev)).cuda_stream
        self.decompress_device(base, src_off, src_size, out.data_ptr(), dst_off, dst_size, res.data_ptr(), options, sized, st)
        outs = [out[int(o):int(o) + int(z)].view(out_dtype) for o, z in zip(dst_off, dst_size)]
        return outs, res

    def synchronize(self):
        if self.lib.vbz_b200_synchronize(self.ctx) != 0:
            self._fail("vbz_b200_synchronize")
