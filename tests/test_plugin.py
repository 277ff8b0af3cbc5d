"""The HDF5 filter 32020 entry points of the product library, driven the way libhdf5 drives them (one malloc'd chunk
buffer per call through H5PLget_plugin_info()->filter), without libhdf5: restates vbz_plugin/test/vbz_hdf_plugin_test.cpp
and python/test/test_vbz_filter.py:57-73 over raw chunk bytes.  Where oracle/_ref holds the compiled
reference plugin, chunks written by one library are read by the other."""
import ctypes
import os

import numpy as np
import pytest

import data_gen
import golden_data
import oracle_lib as ol
import product

libc = ctypes.CDLL(None)
libc.malloc.restype = ctypes.c_void_p
libc.malloc.argtypes = [ctypes.c_size_t]
libc.free.argtypes = [ctypes.c_void_p]

REVERSE = 0x0100


class H5ZClass2(ctypes.Structure):
    _fields_ = [("version", ctypes.c_int), ("id", ctypes.c_int), ("encoder_present", ctypes.c_uint), ("decoder_present", ctypes.c_uint),
                ("name", ctypes.c_char_p), ("can_apply", ctypes.c_void_p), ("set_local", ctypes.c_void_p), ("filter", ctypes.c_void_p)]


FILTER = ctypes.CFUNCTYPE(ctypes.c_size_t, ctypes.c_uint, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint), ctypes.c_size_t,
                          ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p))


class Plugin:
    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.H5PLget_plugin_type.restype = ctypes.c_int
        self.lib.H5PLget_plugin_info.restype = ctypes.c_void_p
        self.info = H5ZClass2.from_address(self.lib.H5PLget_plugin_info())
        self.filter = FILTER(self.info.filter)

    def run(self, flags, cd, data):
        n = len(data)
        buf = libc.malloc(max(n, 1))
        ctypes.memmove(buf, bytes(data), n)
        p = ctypes.c_void_p(buf)
        size = ctypes.c_size_t(n)
        cdv = (ctypes.c_uint * len(cd))(*cd)
        used = self.filter(flags, len(cd), cdv, n, ctypes.byref(size), ctypes.byref(p))
        out = ctypes.string_at(p.value, used) if used else b""
        same_buffer = p.value == buf
        libc.free(p)
        return used, out, size.value, same_buffer


@pytest.fixture(params=product.BACKENDS)
def plugin(request):
    capi, _, _ = product.get(request.param)
    return Plugin(capi.path)


def test_plugin_identity(plugin):
    assert plugin.lib.H5PLget_plugin_type() == 0  # H5PL_TYPE_FILTER
    i = plugin.info
    assert (i.version, i.id, i.encoder_present, i.decoder_present, i.name) == (1, 32020, 1, 1, b"vbz")
    assert i.can_apply is None and i.set_local is None


def test_stored_fixture_chunks_decode(plugin):
    """python/test/test_vbz_filter.py:57-73: Raw/Signal is the same in the zip, vbz and vbz_v1 files."""
    cd0, cd1 = golden_data.fast5_cd_values()
    for rd in golden_data.fast5_reads()[:4]:
        for cd in (cd0, cd1):
            used, out, _, same = plugin.run(REVERSE, cd, rd["vbz_chunk"])
            assert used == rd["raw_bytes"] and out == rd["samples"].tobytes() and not same


@pytest.mark.parametrize("dtype,cd", [(np.int16, (0, 2, 1, 1)), (np.int16, (1, 2, 1, 5)), (np.int32, (0, 4, 1, 1)), (np.uint32, (0, 4, 0, 1)),
                                      (np.int8, (0, 1, 1, 1)), (np.int16, (0, 2, 1)), (np.int16, (0, 2, 1, 0)), (np.int16, (0, 2, 1, 1, 1))])
def test_write_then_read(plugin, dtype, cd):
    """vbz_hdf_plugin_test.cpp:15-136 without the file: chunk = count/8 elements; cd_nelmts 3 means level 1; extras ignored."""
    rng = np.random.default_rng(3)
    ref_path = os.path.join(ol.ROOT, "oracle", "_ref", "libvbz_hdf_plugin_ref.so")
    ref = Plugin(ref_path) if os.path.exists(ref_path) else None
    for n in (100, 12500):
        arr = (rng.integers(0, 1 << 20, n).astype(dtype) if dtype == np.uint32 else data_gen.nanopore_like(n, rng).astype(dtype))
        used, chunk, alloc, same = plugin.run(0, cd, arr.tobytes())
        assert used > 4 and not same and alloc >= used
        assert int.from_bytes(chunk[:4], "little") == arr.nbytes
        used2, back, _, _ = plugin.run(REVERSE, cd, chunk)
        assert used2 == arr.nbytes and back == arr.tobytes()
        if ref is not None:  # cross-library: the reference reads what we wrote, we read what it wrote
            ur, br, _, _ = ref.run(REVERSE, cd, chunk)
            assert ur == arr.nbytes and br == arr.tobytes()
            uw, cw, _, _ = ref.run(0, cd, arr.tobytes())
            u3, b3, _, _ = plugin.run(REVERSE, cd, cw)
            assert u3 == arr.nbytes and b3 == arr.tobytes()


def test_failures_leave_the_buffer_alone(plugin):
    arr = np.arange(100, dtype=np.int16).tobytes()
    assert plugin.run(0, (0, 2), arr)[0] == 0                      # cd_nelmts < 3 (vbz_plugin.cpp:109-112)
    used, _, _, same = plugin.run(0, (0, 4, 1, 1), arr[:-2])       # size % integer_size (vbz_plugin.cpp:194-199)
    assert used == 0 and same
    used, _, _, same = plugin.run(REVERSE, (0, 2, 1, 1), b"\\x10\\x00\\x00\\x00garbage-not-a-zstd-frame")
    assert used == 0 and same
    used, _, _, same = plugin.run(REVERSE, (0, 2, 1, 1), b"\\x01")
    assert used == 0 and same


def test_pyvbz_style_calls(plugin):
    """python/pyvbz/tests: compress()/decompress() defaults from the dtype, through the package's mirror of pyvbz."""
    from vbz_compression_b200 import _lib as vlib
    from vbz_compression_b200 import api
    lib = vlib.load(plugin.lib._name)
    saved = api._lib_handle if hasattr(api, "_lib_handle") else None
    rng = np.random.default_rng(1)
    for dtype in (np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32):
        for arr in (np.arange(1, 11).astype(dtype), rng.integers(0, 100, 20000).astype(dtype)):
            o = vlib.CompressionOptions(np.issubdtype(dtype, np.signedinteger), arr.dtype.itemsize, 1, 0)
            cap = lib.vbz_max_compressed_size(arr.nbytes, ctypes.byref(o))
            out = np.empty(cap, np.uint8)
            n = lib.vbz_compress_sized(arr.ctypes.data, arr.nbytes, out.ctypes.data, cap, ctypes.byref(o))
            assert not lib.vbz_is_error(n)
            assert lib.vbz_decompressed_size(out.ctypes.data, n, ctypes.byref(o)) == arr.nbytes
            back = np.empty(arr.size, dtype)
            m = lib.vbz_decompress_sized(out.ctypes.data, n, back.ctypes.data, back.nbytes, ctypes.byref(o))
            assert m == arr.nbytes and (back == arr).all()
    del saved


@pytest.mark.parametrize("backend", product.BACKENDS)
def test_batch_form_of_the_filter_callback(backend):
    """vbz_b200_h5z_filter_batch (include/vbz_b200.h): many chunks of one dataset per call, results[i] = what the reference's sized call
    returns for chunk i (vbz_plugin.cpp:136-229 per chunk).  Both directions, three and four cd_values, chunks written by the reference,
    a damaged chunk among good ones, and fewer than three cd_values refused."""
    capi, batch, lib = product.get(backend)
    ref = ol.ref() or ol.port()
    rng = np.random.default_rng(17)
    reads = [data_gen.nanopore_like(int(n), rng) for n in (1, 777, 40_000, 0, 123_456, 9)]
    vp = ctypes.c_void_p

    def call(flags, cd, srcs, caps):
        n = len(srcs)
        arrs = [np.frombuffer(bytes(s), np.uint8) for s in srcs]
        outs = [np.full(max(c, 1), 0xEE, np.uint8) for c in caps]
        res = (ctypes.c_uint32 * n)()
        rc = lib.vbz_b200_h5z_filter_batch(batch.ctx, flags, len(cd), (ctypes.c_uint * max(len(cd), 1))(*cd), n, (vp * n)(*[a.ctypes.data if a.size else None for a in arrs]),
                                           (ctypes.c_uint32 * n)(*[a.size for a in arrs]), (vp * n)(*[o.ctypes.data for o in outs]), (ctypes.c_uint32 * n)(*caps), res)
        return rc, [int(r) for r in res], outs

    for cd in ([1, 2, 1, 1], [0, 2, 1], [1, 2, 1, 3]):
        o = ol.Options(True, 2, cd[3] if len(cd) > 3 else 1, cd[0])
        caps = [int(capi.max_compressed_size(r.nbytes, o)) + 4 for r in reads]
        rc, sizes, outs = call(0, cd, [r.tobytes() for r in reads], caps)
        assert rc == 0
        chunks = []
        for r, z, out in zip(reads, sizes, outs):
            assert not ol.is_error(z) and z <= out.size
            code, raw = ref.decompress(out[:z].tobytes(), o, r.nbytes, sized=True)  # the reference reads what the batch call wrote
            assert code == r.nbytes and raw == r.tobytes()
            chunks.append(out[:z].tobytes())
        theirs = [ref.compress(r, o, sized=True)[1] for r in reads]  # and the batch call reads what the reference wrote
        for written in (chunks, theirs):
            rc, sizes, outs = call(REVERSE, cd, written, [r.nbytes for r in reads])
            assert rc == 0
            for r, z, out in zip(reads, sizes, outs):
                assert z == r.nbytes and out[:z].tobytes() == r.tobytes()
        broken = list(theirs)
        broken[2] = broken[2][:len(broken[2]) // 2]
        rc, sizes, outs = call(REVERSE, cd, broken, [r.nbytes for r in reads])
        assert rc == 0
        for i, (r, z, out) in enumerate(zip(reads, sizes, outs)):
            want, _ = ref.decompress(broken[i], o, r.nbytes, sized=True)
            assert z == want and (ol.is_error(z) or out[:z].tobytes() == r.tobytes()), i
    rc, _, _ = call(REVERSE, [1, 2], [b"\0" * 8], [8])
    assert rc == -1
