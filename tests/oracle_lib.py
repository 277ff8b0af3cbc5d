"""ctypes access to the two checkers (TEST INFRASTRUCTURE ONLY):

  port : oracle/liboracle_port.so      - this repo's plain-C restatement (oracle/vbz_oracle.c, zstd_oracle.c)
  ref  : oracle/_ref/libvbz_ref.so     - the unmodified reference compiled from /root/reference by oracle/Makefile
         oracle/_ref/libvbz_ref_generic.so - same, without -mssse3 (what ARM builds run)

Both expose the reference's C API shape: f(src, n, dst, cap, options*) -> uint32.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ERR_ZSTD = 0xFFFFFFFF
ERR_INPUT_SIZE = 0xFFFFFFFE
ERR_INTEGER_SIZE = 0xFFFFFFFD
ERR_DEST_SIZE = 0xFFFFFFFC
ERR_STREAM = 0xFFFFFFFB
ERR_VERSION = 0xFFFFFFFA
ERR_OOM = 0xFFFFFFF9


def is_error(v):
    return v >= ERR_OOM


import sys
sys.path.insert(0, ROOT)
# one ctypes declaration of the 16-byte options struct (ref vbz/vbz.h:29-53) for oracle and product alike
from vbz_compression_b200._lib import CompressionOptions as Options  # noqa: E402


assert ctypes.sizeof(Options) == 16


class VbzLib:
    """Uniform wrapper over any library exporting the vbz C API under a name prefix."""

    def __init__(self, path, prefix="vbz_"):
        self.lib = ctypes.CDLL(path)
        self.path = path
        P = ctypes.POINTER(Options)
        for n in ("compress", "decompress", "compress_sized", "decompress_sized"):
            f = getattr(self.lib, prefix + n)
            f.restype = ctypes.c_uint32
            f.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, P]
            setattr(self, "_" + n, f)
        f = getattr(self.lib, prefix + "max_compressed_size")
        f.restype = ctypes.c_uint32
        f.argtypes = [ctypes.c_uint32, P]
        self._max = f
        f = getattr(self.lib, prefix + "decompressed_size")
        f.restype = ctypes.c_uint32
        f.argtypes = [ctypes.c_void_p, ctypes.c_uint32, P]
        self._dsize = f
        f = getattr(self.lib, prefix + "is_error")
        f.restype = ctypes.c_bool if prefix == "vbz_" else ctypes.c_int
        f.argtypes = [ctypes.c_uint32]
        self._is_error = f
        f = getattr(self.lib, prefix + "error_string")
        f.restype = ctypes.c_char_p
        f.argtypes = [ctypes.c_uint32]
        self._error_string = f

    @staticmethod
    def _buf(b):
        a = np.frombuffer(bytes(b), np.uint8) if not isinstance(b, np.ndarray) else np.ascontiguousarray(b).view(np.uint8).reshape(-1)
        return a

    def max_compressed_size(self, n, opts):
        return self._max(n, ctypes.byref(opts))

    def decompressed_size(self, src, opts):
        a = self._buf(src)
        return self._dsize(a.ctypes.data if a.size else None, a.size, ctypes.byref(opts))

    def is_error(self, v):
        return bool(self._is_error(v))

    def error_string(self, v):
        return self._error_string(v).decode()

    def _call(self, fn, src, cap, opts):
        a = self._buf(src)
        dst = np.zeros(max(int(cap), 1), np.uint8)
        r = fn(a.ctypes.data if a.size else None, a.size, dst.ctypes.data, int(cap), ctypes.byref(opts))
        return r, dst

    def compress(self, src, opts, sized=True, cap=None):
        a = self._buf(src)
        if cap is None:
            cap = self.max_compressed_size(a.size, opts)
            if is_error(cap):
                return cap, b""
        r, dst = self._call(self._compress_sized if sized else self._compress, a, cap, opts)
        return r, (dst[:r].tobytes() if not is_error(r) and r <= cap else b"")

    def decompress(self, src, opts, dst_size, sized=True):
        r, dst = self._call(self._decompress_sized if sized else self._decompress, src, dst_size, opts)
        return r, (dst[:r].tobytes() if not is_error(r) and r <= dst_size else b"")


_cache = {}


def port():
    if "port" not in _cache:
        p = VbzLib(os.path.join(ROOT, "oracle", "liboracle_port.so"), prefix="vbzo_")
        L = p.lib
        L.vbzo_stage1_encode.restype = ctypes.c_uint32
        L.vbzo_stage1_encode.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.vbzo_stage1_decode.restype = ctypes.c_uint32
        L.vbzo_stage1_decode.argtypes = L.vbzo_stage1_encode.argtypes
        L.vbzo_zstd_decompress.restype = ctypes.c_size_t
        L.vbzo_zstd_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.vbzo_libzstd_decompress.restype = ctypes.c_size_t
        L.vbzo_libzstd_decompress.argtypes = L.vbzo_zstd_decompress.argtypes
        L.vbzo_libzstd_compress.restype = ctypes.c_size_t
        L.vbzo_libzstd_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        L.vbzo_zstd_frame_content_size.restype = ctypes.c_ulonglong
        L.vbzo_zstd_frame_content_size.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.vbzo_zstd_compress_bound.restype = ctypes.c_size_t
        L.vbzo_zstd_compress_bound.argtypes = [ctypes.c_size_t]
        L.vbzo_set_int16_x86_semantics.argtypes = [ctypes.c_int]
        L.vbzo_set_use_restated_zstd_decoder.argtypes = [ctypes.c_int]
        _cache["port"] = p
    return _cache["port"]


def ref(generic=False):
    key = "ref_generic" if generic else "ref"
    if key not in _cache:
        path = os.path.join(ROOT, "oracle", "_ref", "libvbz_ref_generic.so" if generic else "libvbz_ref.so")
        _cache[key] = VbzLib(path) if os.path.exists(path) else None
    return _cache[key]


# ---- helpers over the port's extra entry points --------------------------------
def stage1_encode(arr, isz, zz, version=0):
    a = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    n = a.size // isz
    dst = np.zeros((n + 3) // 4 + 4 * n + 16, np.uint8)
    r = port().lib.vbzo_stage1_encode(a.ctypes.data if a.size else None, a.size, dst.ctypes.data, dst.size, isz, int(zz), version)
    assert not is_error(r), r
    return dst[:r].tobytes()


def stage1_decode(buf, nbytes, isz, zz, version=0):
    a = np.frombuffer(bytes(buf), np.uint8)
    dst = np.zeros(max(nbytes, 1), np.uint8)
    r = port().lib.vbzo_stage1_decode(a.ctypes.data if a.size else None, a.size, dst.ctypes.data, nbytes, isz, int(zz), version)
    return r, dst[:nbytes].tobytes()


def zstd_decompress_port(frame, cap):
    a = np.frombuffer(bytes(frame), np.uint8)
    dst = np.zeros(max(cap, 1), np.uint8)
    r = port().lib.vbzo_zstd_decompress(dst.ctypes.data, cap, a.ctypes.data if a.size else None, a.size)
    return (None if r == 2**64 - 1 else dst[:r].tobytes())


def zstd_decompress_lib(frame, cap):
    a = np.frombuffer(bytes(frame), np.uint8)
    dst = np.zeros(max(cap, 1), np.uint8)
    r = port().lib.vbzo_libzstd_decompress(dst.ctypes.data, cap, a.ctypes.data if a.size else None, a.size)
    return (None if r == 2**64 - 1 else dst[:r].tobytes())


def zstd_compress_lib(data, level):
    a = np.frombuffer(bytes(data), np.uint8)
    cap = port().lib.vbzo_zstd_compress_bound(a.size) + 64
    dst = np.zeros(cap, np.uint8)
    r = port().lib.vbzo_libzstd_compress(dst.ctypes.data, cap, a.ctypes.data if a.size else None, a.size, level)
    assert r != 2**64 - 1
    return dst[:r].tobytes()


def frame_content_size(frame):
    a = np.frombuffer(bytes(frame), np.uint8)
    return port().lib.vbzo_zstd_frame_content_size(a.ctypes.data if a.size else None, a.size)
