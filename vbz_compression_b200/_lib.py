"""ctypes binding of libvbz_b200.so (the C-ABI of include/vbz.h, include/vbz_b200.h, include/vbz_plugin.h).

There is no Python or CPU implementation behind this module: if the shared library is missing it is
built with nvcc (vbz_compression_b200/build.py); if that fails the import raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvbz_b200.so")

ERR_ZSTD = 0xFFFFFFFF
ERR_INPUT_SIZE = 0xFFFFFFFE
ERR_INTEGER_SIZE = 0xFFFFFFFD
ERR_DESTINATION_SIZE = 0xFFFFFFFC
ERR_STREAM = 0xFFFFFFFB
ERR_VERSION = 0xFFFFFFFA
ERR_OUT_OF_MEMORY = 0xFFFFFFF9


class CompressionOptions(ctypes.Structure):
    """reference vbz/vbz.h:29-53 (16 bytes)"""
    _fields_ = [("perform_delta_zig_zag", ctypes.c_bool), ("integer_size", ctypes.c_uint),
                ("zstd_compression_level", ctypes.c_uint), ("vbz_version", ctypes.c_uint)]


def bind(lib):
    """Declare argtypes/restypes of every exported symbol on a loaded CDLL."""
    P = ctypes.POINTER(CompressionOptions)
    vp, u32, u64p, u32p = ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
    for n in ("vbz_compress", "vbz_decompress", "vbz_compress_sized", "vbz_decompress_sized"):
        f = getattr(lib, n)
        f.restype, f.argtypes = u32, [vp, u32, vp, u32, P]
    lib.vbz_max_compressed_size.restype, lib.vbz_max_compressed_size.argtypes = u32, [u32, P]
    lib.vbz_decompressed_size.restype, lib.vbz_decompressed_size.argtypes = u32, [vp, u32, P]
    lib.vbz_is_error.restype, lib.vbz_is_error.argtypes = ctypes.c_bool, [u32]
    lib.vbz_error_string.restype, lib.vbz_error_string.argtypes = ctypes.c_char_p, [u32]
    lib.vbz_b200_device_count.restype, lib.vbz_b200_device_count.argtypes = ctypes.c_int, []
    lib.vbz_b200_create.restype, lib.vbz_b200_create.argtypes = vp, [ctypes.c_int]
    lib.vbz_b200_destroy.restype, lib.vbz_b200_destroy.argtypes = None, [vp]
    lib.vbz_b200_last_error.restype, lib.vbz_b200_last_error.argtypes = ctypes.c_char_p, []
    lib.vbz_b200_compress_bound.restype, lib.vbz_b200_compress_bound.argtypes = ctypes.c_uint64, [u32, P, ctypes.c_int]
    lib.vbz_b200_stream.restype, lib.vbz_b200_stream.argtypes = vp, [vp]
    lib.vbz_b200_synchronize.restype, lib.vbz_b200_synchronize.argtypes = ctypes.c_int, [vp]
    lib.vbz_b200_kernel_launches.restype, lib.vbz_b200_kernel_launches.argtypes = ctypes.c_uint64, [vp]
    lib.vbz_b200_set_stage_timing.restype, lib.vbz_b200_set_stage_timing.argtypes = ctypes.c_int, [vp, ctypes.c_int]
    lib.vbz_b200_stage_times.restype = ctypes.c_int
    lib.vbz_b200_stage_times.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    for n in ("vbz_b200_compress_batch", "vbz_b200_decompress_batch"):
        f = getattr(lib, n)
        f.restype, f.argtypes = ctypes.c_int, [vp, ctypes.POINTER(vp), u32p, ctypes.POINTER(vp), u32p, u32p, u32, P, ctypes.c_int]
    for n in ("vbz_b200_compress_device", "vbz_b200_decompress_device"):
        f = getattr(lib, n)
        f.restype, f.argtypes = ctypes.c_int, [vp, vp, u64p, u32p, vp, u64p, u32p, vp, u32, P, ctypes.c_int, vp]
    lib.H5PLget_plugin_type.restype, lib.H5PLget_plugin_type.argtypes = ctypes.c_int, []
    lib.H5PLget_plugin_info.restype, lib.H5PLget_plugin_info.argtypes = vp, []
    lib.vbz_plugin_info.restype, lib.vbz_plugin_info.argtypes = vp, []
    lib.vbz_b200_h5z_filter.restype = ctypes.c_size_t
    lib.vbz_b200_h5z_filter.argtypes = [ctypes.c_uint, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint), ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(vp)]
    lib.vbz_b200_decode_stats.restype, lib.vbz_b200_decode_stats.argtypes = ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_uint32)]
    lib.vbz_b200_h5z_filter_batch.restype = ctypes.c_int
    lib.vbz_b200_h5z_filter_batch.argtypes = [vp, ctypes.c_uint, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint), u32, ctypes.POINTER(vp), u32p,
                                              ctypes.POINTER(vp), u32p, u32p]
    return lib


EXPORTED_SYMBOLS = [
    "vbz_is_error", "vbz_error_string", "vbz_max_compressed_size", "vbz_compress", "vbz_decompress", "vbz_compress_sized",
    "vbz_decompress_sized", "vbz_decompressed_size",
    "vbz_b200_device_count", "vbz_b200_create", "vbz_b200_destroy", "vbz_b200_last_error", "vbz_b200_compress_batch",
    "vbz_b200_decompress_batch", "vbz_b200_compress_device", "vbz_b200_decompress_device", "vbz_b200_synchronize",
    "vbz_b200_kernel_launches", "vbz_b200_compress_bound", "vbz_b200_stream", "vbz_b200_set_stage_timing", "vbz_b200_stage_times", "vbz_b200_decode_stats",
    "H5PLget_plugin_type", "H5PLget_plugin_info", "vbz_plugin_info", "vbz_b200_h5z_filter", "vbz_b200_h5z_filter_batch",
]

_lib = None


def load(path=None):
    """Load (building it first if needed) libvbz_b200.so.  Raises if it cannot be had."""
    global _lib
    if path is None:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_native()
        _lib = bind(ctypes.CDLL(LIB_PATH))
        return _lib
    return bind(ctypes.CDLL(path))
