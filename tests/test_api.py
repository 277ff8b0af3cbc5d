"""the pyvbz mirror (vbz_compression_b200.api compress / decompress / compression_options) has the reference
module's semantics (python/pyvbz/vbz/__init__.py:12-76) - restates the reference's python/pyvbz/tests/unit: default options from
the dtype (zigzag for signed integers, integer_size = itemsize, level 1, version 0), sized format out, `output[:size]` back -
and the README example.  What the product writes must decode with the reference, and the other way round.  Also: the product's
vbz_max_compressed_size equals the reference's value for value (ref vbz/vbz.cpp:79-114)."""
import numpy as np
import pytest

import oracle_lib as ol
import product
from oracle_lib import Options
from vbz_compression_b200 import _lib as vlib
from vbz_compression_b200 import api


@pytest.fixture(params=product.BACKENDS)
def pyvbz(request, monkeypatch):
    _, _, lib = product.get(request.param)
    monkeypatch.setattr(vlib, "_lib", lib)  # api.compress / api.decompress load the library through _lib.load()
    return request.param


DTYPES = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32]


@pytest.mark.parametrize("version", [None, 1])
@pytest.mark.parametrize("dtype", DTYPES)
def test_basic_roundtrip(pyvbz, dtype, version):
    data = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], dtype=dtype)
    options = None
    if version:
        options = api.compression_options(np.issubdtype(data.dtype, np.signedinteger), data.dtype.itemsize, 1, version)
    res = api.compress(data, options)
    assert res.dtype == np.uint8 and int.from_bytes(res[:4].tobytes(), "little") == data.nbytes  # the sized format
    rec = api.decompress(res, dtype, options)
    assert rec.dtype == np.dtype(dtype) and np.array_equal(rec, data)


@pytest.mark.parametrize("version", [None, 1])
@pytest.mark.parametrize("dtype", DTYPES)
def test_random_roundtrip_and_cross_decode(pyvbz, dtype, version):
    rng = np.random.default_rng(11)
    info = np.iinfo(dtype)
    size = 20_000 if pyvbz == "cpusim" else 200_000
    hi = (1 << (8 * np.dtype(dtype).itemsize - 1)) - 1
    data = rng.integers(info.min, hi, size=size, dtype=dtype)
    zz = bool(np.issubdtype(np.dtype(dtype), np.signedinteger))
    options = api.compression_options(zz, np.dtype(dtype).itemsize, 1, version) if version else None
    res = api.compress(data, options)
    assert np.array_equal(api.decompress(res, dtype, options), data)
    # both directions against the unmodified reference (the bytes of a zstd frame are the writer's own: decode, do not compare)
    ref = ol.ref() or ol.port()
    o = Options(zz, np.dtype(dtype).itemsize, 1, version or 0)
    r, back = ref.decompress(res, o, data.nbytes, sized=True)
    assert r == data.nbytes and back == data.tobytes()
    r, blob = ref.compress(data, o, sized=True)
    assert not ol.is_error(r)
    assert np.array_equal(api.decompress(np.frombuffer(blob, np.uint8), dtype, options), data)


def test_readme_example(pyvbz):
    signal = np.arange(0, 1000, dtype=np.int16)
    compressed = api.compress(signal)
    assert compressed.nbytes <= 40  # python/pyvbz/README.md:18-23: the reference's libzstd writes 27 bytes
    assert np.array_equal(api.decompress(compressed, dtype=np.int16), signal)


def test_errors_raise(pyvbz):
    with pytest.raises(api.VbzError):
        api.compress(np.zeros(10, np.int16), api.compression_options(True, 3))  # integer size 3
    with pytest.raises(api.VbzError):
        api.decompress(np.zeros(2, np.uint8), np.int16)  # shorter than the sized header


@pytest.mark.parametrize("backend", product.BACKENDS)
def test_max_compressed_size_equals_reference(backend):
    capi = product.get(backend)[0]
    ref = ol.ref() or ol.port()
    sizes = [0, 1, 2, 3, 4, 7, 8, 100, 1000, 2000, 4096, 65535, 65536, 131071, 131072, 131073, 2_000_000, 33_333_333, 0x7FFFFFFF, 0xFFFFFFFF]
    for isz in (0, 1, 2, 3, 4, 8):
        for level in (0, 1, 3, 22):
            for version in (0, 1, 2):
                for zz in (False, True):
                    o = Options(zz, isz, level, version)
                    for n in sizes:
                        assert capi.max_compressed_size(n, o) == ref.max_compressed_size(n, o), (isz, level, version, zz, n)
    # the values SURVEY.md 8a2 quotes
    assert capi.max_compressed_size(2_000_000, Options(True, 2, 1, 0)) == 4_266_621
    assert capi.max_compressed_size(2_000_000, Options(True, 2, 0, 0)) == 4_250_020
