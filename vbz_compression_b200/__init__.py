"""vbz_compression_b200 - B200-native VBZ codec (delta + zigzag + StreamVByte, then zstd) behind the
reference's C API.  The product is the shared library libvbz_b200.so (CUDA, sm_100a); this package is
the Python mirror of the reference's pyvbz module (python/pyvbz/vbz/__init__.py) on top of it, plus
batch helpers.  No CPU implementation exists here."""
from ._lib import CompressionOptions, load, LIB_PATH  # noqa: F401
from .api import compress, decompress, compression_options, VbzError, Batch  # noqa: F401
