"""the minimal fast5 / HDF5 chunk reader finds every VBZ chunk of a file without libhdf5 and decodes them
as one batch.  Restates python/test/test_vbz_filter.py:57-73: Raw/Signal must be equal across the gzip, VBZ v0 and VBZ v1
fixtures.  The three fixture files (the reference's test_data/*.fast5, reference-held test vectors like the rest of tests/golden/)
travel with the repo in tests/golden/fast5/, so the GPU tests run on the box, where /root/reference does not exist."""
import os
import zlib

import numpy as np
import pytest

import product
from vbz_compression_b200 import fast5

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fast5")


def _zip_signals(path):
    f = fast5.Fast5File.open(path)
    out = {}
    for ds in f.datasets():
        if not ds.name.endswith("/Signal"):
            continue
        assert [fid for fid, _ in ds.filters] == [1], ds.filters  # deflate
        parts = [np.frombuffer(zlib.decompress(bytes(f.chunk_bytes(c))), np.int16) for c in ds.chunks]
        out[ds.name] = np.concatenate(parts)[:ds.shape[0]]
    return out


def test_container_structure():
    f = fast5.Fast5File.open(os.path.join(DATA, "multi_fast5_vbz.fast5"))
    sig = [d for d in f.datasets() if d.name.endswith("/Raw/Signal")]
    assert len(sig) == 10
    for d in sig:
        assert d.itemsize == 2 and len(d.shape) == 1 and d.vbz_cd_values[:4] == [0, 2, 1, 1]
        assert len(d.chunks) == 1 and d.chunks[0].offset == (0,)
        assert int.from_bytes(f.chunk_bytes(d.chunks[0])[:4], "little") == 2 * d.shape[0]
    assert sorted(d.shape[0] for d in sig) == [9885, 14510, 15643, 37440, 123627, 136370, 161547, 206976, 337876, 505057]
    f1 = fast5.Fast5File.open(os.path.join(DATA, "multi_fast5_vbz_v1.fast5"))
    assert all(d.vbz_cd_values == [1, 2, 1, 1] for d in f1.datasets() if d.name.endswith("/Raw/Signal"))
    o = fast5.options_from_cd_values([1, 2, 1])
    assert (o.vbz_version, o.integer_size, o.perform_delta_zig_zag, o.zstd_compression_level) == (1, 2, True, 1)
    with pytest.raises(fast5.Fast5FormatError):
        fast5.Fast5File(b"not hdf5 at all" * 100)


@pytest.mark.parametrize("backend", product.BACKENDS)
@pytest.mark.parametrize("name", ["multi_fast5_vbz.fast5", "multi_fast5_vbz_v1.fast5"])
def test_whole_file_decodes_as_one_batch(backend, name):
    _, batch, _ = product.get(backend)
    want = _zip_signals(os.path.join(DATA, "multi_fast5_zip.fast5"))
    l0 = batch.kernel_launches
    got = fast5.read_vbz_signals(os.path.join(DATA, name), batch)
    assert batch.kernel_launches > l0
    assert set(got) == set(want) and len(got) == 10
    for k in want:
        assert got[k].dtype == np.int16 and np.array_equal(got[k], want[k]), k
