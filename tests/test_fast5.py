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


# ---- the writer half (python/fast5compress/fast5vbz.py:17-55) ------------------------------------------------------------------------------

def _signal_header(f, which=0):
    """(dataset, {message type: body bytes}) of one /read_*/Raw/Signal as libhdf5 wrote it."""
    name = sorted(p for p in f.objects() if fast5.is_read_signal(p))[which]
    msgs = f._messages(f.objects()[name])
    return f._dataset(name, msgs), {t: bytes(f.b[d:d + size]) for t, d, size in msgs}


def test_written_structures_equal_what_libhdf5_writes():
    """The pieces the transcoder generates, byte for byte against the same pieces in the reference's files (written by h5py/libhdf5)."""
    v1 = fast5.Fast5File.open(os.path.join(DATA, "multi_fast5_vbz_v1.fast5"))
    ds, msgs = _signal_header(v1)
    assert fast5.vbz_pipeline(1, 2, True, 1) == msgs[0x000B]
    layout = msgs[0x0008]
    tree = int.from_bytes(layout[3:11], "little")
    assert fast5.chunked_layout_message(v1, tree, ds.shape[0], 2) == layout
    c = ds.chunks[0]
    assert fast5.chunk_btree_node(v1, [(c.size, 0, 0, c.address)], ds.shape[0], 2) == bytes(v1.b[tree:tree + v1.chunk_node_size(1)])
    assert v1.chunk_node_size(1) == 2096
    z = fast5.Fast5File.open(os.path.join(DATA, "multi_fast5_zip.fast5"))
    zds, zmsgs = _signal_header(z)
    assert fast5.gzip_pipeline(1) == zmsgs[0x000B]
    # a whole header: the zip fixture's Signal header has no padding message, so it must come out identical
    name = zds.name
    hdr = z.objects()[name]
    raw = [(t, z.b[d - 4], bytes(z.b[d:d + size])) for t, d, size in z._messages(hdr)]
    size = 16 + z._u(hdr + 8, 4)
    assert fast5.object_header(raw, 1) == bytes(z.b[hdr:hdr + size])


@pytest.mark.parametrize("backend", product.BACKENDS)
@pytest.mark.parametrize("vbz_version", [0, 1])
def test_transcode_gzip_to_vbz_and_back(backend, vbz_version):
    capi, batch, _ = product.get(backend)
    src = open(os.path.join(DATA, "multi_fast5_zip.fast5"), "rb").read()
    want = _zip_signals(os.path.join(DATA, "multi_fast5_zip.fast5"))
    l0 = batch.kernel_launches
    out = fast5.transcode(src, batch, vbz_version=vbz_version)
    assert batch.kernel_launches > l0
    f, g = fast5.Fast5File(src), fast5.Fast5File(out)
    # the signal: through this library, and every chunk through the unmodified reference
    got = fast5.read_vbz_signals(out, batch)
    assert set(got) == set(want)
    ref = product.oracle_lib.ref() or product.oracle_lib.port()
    for d in g.datasets():
        assert np.array_equal(got[d.name], want[d.name]), d.name
        assert d.vbz_cd_values == [vbz_version, 2, 1, 1] and d.chunk_shape == d.shape and len(d.chunks) == 1
        opts = product.oracle_lib.Options(True, 2, 1, vbz_version)
        code, raw = ref.decompress(np.frombuffer(g.chunk_bytes(d.chunks[0]), np.uint8), opts, 2 * d.shape[0], sized=True)
        assert code == 2 * d.shape[0] and np.array_equal(np.frombuffer(raw, np.int16), want[d.name])
    # everything else: same objects, same attributes (variable-length strings followed into the global heap)
    fo, go = f.objects(), g.objects()
    assert set(fo) == set(go) and len(go) == 151
    assert sum(len(f.attributes(h)) for h in fo.values()) == 701
    for p in fo:
        assert f.attributes(fo[p]) == g.attributes(go[p]), p
    # the file shrank by what the chunks shrank, give or take the new headers, and says so itself
    saved = sum(c.size for d in f.datasets() for c in d.chunks) - sum(c.size for d in g.datasets() for c in d.chunks)
    assert abs((len(src) - len(out)) - saved) < 4096
    assert g._off(g.eof_pos) == len(out)
    assert len(out) <= len(open(os.path.join(DATA, "multi_fast5_vbz.fast5"), "rb").read())  # the reference's file with the same objects
    # without compaction: same content, the size of the source
    holes = fast5.transcode(src, batch, vbz_version=vbz_version, compact_file=False)
    assert len(holes) == len(src)
    assert all(np.array_equal(v, want[k]) for k, v in fast5.read_vbz_signals(holes, batch).items())
    # and back to gzip (-d): what the zip fixture holds, chunk for chunk
    back = fast5.transcode(out, batch, decompress=True)
    b = fast5.Fast5File(back)
    for d in b.datasets():
        assert [fid for fid, _ in d.filters] == [1] and len(d.chunks) == 1
        assert np.array_equal(np.frombuffer(zlib.decompress(bytes(b.chunk_bytes(d.chunks[0]))), np.int16), want[d.name])
    assert len(back) == len(src)
    bo = b.objects()
    for p in fo:
        assert f.attributes(fo[p]) == b.attributes(bo[p]), p


def test_compact_moves_every_address():
    """Cutting nothing changes nothing; cutting a hole in front of everything (a dead region made by hand) moves all
    addresses - and an address into a hole is refused."""
    src = open(os.path.join(DATA, "multi_fast5_vbz_v1.fast5"), "rb").read()
    assert fast5.compact(src, []) == src
    f = fast5.Fast5File(src)
    d = sorted(f.datasets(), key=lambda d: d.chunks[0].address)[0]
    c = d.chunks[0]
    with pytest.raises(fast5.Fast5FormatError):
        fast5.compact(src, [(c.address, c.address + 16)])
    # give the first chunk's space up by pointing its dataset at a copy behind the end of the file
    grown = bytearray(src) + bytes(f.chunk_bytes(c))
    sites = [p for p in fast5._address_sites(f) if f._off(p) == c.address]
    assert len(sites) == 1
    grown[sites[0]:sites[0] + 8] = len(src).to_bytes(8, "little")
    out = fast5.compact(bytes(grown), [(c.address, c.address + c.size)])
    assert len(out) == len(src)
    g = fast5.Fast5File(out)
    a, b = {x.name: x for x in f.datasets()}, {x.name: x for x in g.datasets()}
    for k in a:
        assert bytes(f.chunk_bytes(a[k].chunks[0])) == bytes(g.chunk_bytes(b[k].chunks[0])), k
    fo, go = f.objects(), g.objects()
    assert all(f.attributes(fo[p]) == g.attributes(go[p]) for p in fo)


def test_fast5compress_command_line(tmp_path, capsys):
    """The reference script's interface: files in, '<file><suffix>' out, names printed (fast5vbz.py:58-76)."""
    import shutil
    from vbz_compression_b200 import fast5compress
    _, batch, lib = product.get("cpusim")
    src = tmp_path / "reads.fast5"
    shutil.copy(os.path.join(DATA, "multi_fast5_zip.fast5"), src)
    fast5compress.Batch = lambda device: product.Batch(device, lib=lib)
    fast5compress.main([str(src), "-s", ".vbz", "--vbz-version", "0"])
    printed = capsys.readouterr().out.strip()
    assert printed == str(src) + ".vbz"
    want = _zip_signals(str(src))
    got = fast5.read_vbz_signals(printed, batch)
    assert all(np.array_equal(got[k], want[k]) for k in want) and len(got) == 10
    fast5compress.main([printed, "-d", "-s", ".gz"])
    back = _zip_signals(printed + ".gz")
    assert all(np.array_equal(back[k], want[k]) for k in want)


def test_types_that_hold_addresses_are_recognised():
    """`compact` moves bytes: data whose elements are file addresses (variable-length values = global heap IDs, references) must be found
    wherever a datatype can hide them - or the file refused."""
    import struct
    i16 = bytes([0x10, 0x08, 0, 0]) + struct.pack("<I", 2) + bytes([0, 0, 16, 0])
    fstr = bytes([0x13, 0, 0, 0]) + struct.pack("<I", 7)
    f64 = bytes([0x11, 0x20, 0x3F, 0]) + struct.pack("<I", 8) + bytes(12)
    vlen = bytes([0x19, 0x01, 0, 0]) + struct.pack("<I", 16) + fstr
    ref = bytes([0x17, 0, 0, 0]) + struct.pack("<I", 8)

    def member_v1(name, off, dt):
        nm = name + b"\0"
        return nm + bytes(-len(nm) % 8) + struct.pack("<I", off) + bytes(28) + dt

    def compound_v1(members, size):
        return bytes([0x16, len(members), 0, 0]) + struct.pack("<I", size) + b"".join(member_v1(n, o, d) for n, o, d in members)

    plain = compound_v1([(b"start", 0, i16), (b"a_name_longer_than_eight", 2, f64), (b"s", 10, fstr)], 17)
    hiding = compound_v1([(b"start", 0, i16), (b"label", 2, vlen)], 18)
    nested = compound_v1([(b"inner", 0, compound_v1([(b"r", 0, ref)], 8))], 8)
    array_v2 = bytes([0x2A, 0, 0, 0]) + struct.pack("<I", 6) + bytes([1, 0, 0, 0]) + struct.pack("<II", 3, 0) + i16
    array_of_vlen = bytes([0x3A, 0, 0, 0]) + struct.pack("<I", 32) + bytes([1]) + struct.pack("<I", 2) + vlen
    compound_v3 = bytes([0x36, 2, 0, 0]) + struct.pack("<I", 10) + b"x\0" + bytes([0]) + i16 + b"yy\0" + bytes([2]) + fstr
    for msg, want in ((i16, False), (fstr, False), (f64, False), (vlen, True), (ref, True), (plain, False), (hiding, True), (nested, True),
                      (array_v2, False), (array_of_vlen, True), (compound_v3, False)):
        holds, used = fast5._datatype_span(msg + b"\xEE" * 8, 0)
        assert holds == want and used == len(msg), (msg.hex(), holds, used, len(msg))
    # the datatype messages libhdf5 wrote into the fixture are walked to their exact ends too
    f = fast5.Fast5File.open(os.path.join(DATA, "multi_fast5_zip.fast5"))
    seen = 0
    for hdr in f.objects().values():
        for mtype, d, size in f._messages(hdr):
            if mtype == 0x0003:
                holds, used = fast5._datatype_span(f.b, d)
                assert not holds and used <= size and size - used < 8
                seen += 1
            elif mtype == 0x000C:
                name, tp, tlen, sp, slen, dp = f._attribute_parts(d)
                holds, used = fast5._datatype_span(f.b, tp)
                assert used == tlen, (name, used, tlen)
                seen += 1
    assert seen > 700


@pytest.mark.parametrize("backend", product.BACKENDS)
def test_transcode_between_vbz_versions_is_stable(backend):
    """VBZ v0 -> v1 keeps the samples, doing it again changes nothing (same frames, same layout), and a file with nothing to replace comes
    back byte for byte."""
    _, batch, _ = product.get(backend)
    v0 = open(os.path.join(DATA, "multi_fast5_vbz.fast5"), "rb").read()
    want = fast5.read_vbz_signals(v0, batch)
    a = fast5.transcode(v0, batch, vbz_version=1)
    got = fast5.read_vbz_signals(a, batch)
    assert set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)
    assert all(d.vbz_cd_values == [1, 2, 1, 1] for d in fast5.Fast5File(a).datasets())
    assert fast5.transcode(a, batch, vbz_version=1) == a
    assert fast5.transcode(v0, batch, name_filter=lambda name: False) == v0


@pytest.mark.parametrize("backend", product.BACKENDS)
def test_transcode_into_several_chunks(backend):
    """chunk_elems: datasets of several chunks, edge chunks stored whole - written, read back (the VBZ batch path and the zlib path both
    walk multi-chunk datasets), and every chunk decoded by the unmodified reference."""
    _, batch, _ = product.get(backend)
    src = open(os.path.join(DATA, "multi_fast5_zip.fast5"), "rb").read()
    want = _zip_signals(os.path.join(DATA, "multi_fast5_zip.fast5"))
    out = fast5.transcode(src, batch, vbz_version=0, chunk_elems=20_000)
    g = fast5.Fast5File(out)
    ref = product.oracle_lib.ref() or product.oracle_lib.port()
    opts = product.oracle_lib.Options(True, 2, 1, 0)
    for d in g.datasets():
        n = d.shape[0]
        assert d.chunk_shape == (20_000,) and len(d.chunks) == (n + 19_999) // 20_000 and [c.offset[0] for c in d.chunks] == list(range(0, n, 20_000))
        for c in d.chunks:
            code, raw = ref.decompress(np.frombuffer(g.chunk_bytes(c), np.uint8), opts, 40_000, sized=True)
            part = want[d.name][c.offset[0]:c.offset[0] + 20_000]
            assert code == 40_000 and np.array_equal(np.frombuffer(raw, np.int16)[:part.size], part) and not np.frombuffer(raw, np.int16)[part.size:].any()
    got = fast5.read_vbz_signals(out, batch)
    assert all(np.array_equal(got[k], want[k]) for k in want) and len(got) == 10
    back = fast5.transcode(out, batch, decompress=True, chunk_elems=30_000)  # several VBZ chunks in, several gzip chunks out
    b = fast5.Fast5File(back)
    for d in b.datasets():
        parts = [np.frombuffer(zlib.decompress(bytes(b.chunk_bytes(c))), np.int16) for c in d.chunks]
        assert all(p.size == 30_000 for p in parts) and np.array_equal(np.concatenate(parts)[:d.shape[0]], want[d.name])
    again = fast5.transcode(back, batch, vbz_version=1)  # and several gzip chunks in
    got = fast5.read_vbz_signals(again, batch)
    assert all(np.array_equal(got[k], want[k]) for k in want)
    with pytest.raises(fast5.Fast5FormatError):
        fast5.transcode(src, batch, chunk_elems=1000)  # 506 chunks for the longest read: more than one B-tree node holds


def test_libhdf5_filters_are_undone_in_reverse_order():
    """shuffle, deflate, fletcher32 - the other filters a fast5 Signal may carry - and the filter mask that says which were skipped."""
    rng = np.random.default_rng(4)
    arr = rng.integers(-500, 500, 1001).astype("<i2")
    raw = arr.tobytes()
    shuffled = np.frombuffer(raw, np.uint8).reshape(-1, 2).T.tobytes()
    ds = fast5.Dataset("/x", (1001,), 2, [(2, [2]), (1, [4]), (3, [])], (1001,), [])
    payload = zlib.compress(shuffled, 4) + bytes([0x12, 0x34, 0x56, 0x78])
    assert fast5._unfilter(ds, fast5.Chunk((0,), 0, 0, len(payload)), payload) == raw
    assert fast5._unfilter(ds, fast5.Chunk((0,), 0b010, 0, 0), shuffled + b"abcd") == raw          # deflate skipped for this chunk
    assert fast5._unfilter(ds, fast5.Chunk((0,), 0b101, 0, 0), zlib.compress(raw, 1)) == raw      # shuffle and checksum skipped
    with pytest.raises(fast5.Fast5FormatError):
        fast5._unfilter(fast5.Dataset("/x", (4,), 2, [(32015, [])], (4,), []), fast5.Chunk((0,), 0, 0, 0), b"12345678")
