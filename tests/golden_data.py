"""Loaders for the committed golden vectors (tests/golden/, made by tests/golden/make_golden.py)."""
import hashlib
import json
import os
import zlib

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kats():
    return json.load(open(os.path.join(HERE, "kat.json")))


def fast5_reads():
    """-> list of dict(raw_bytes, vbz_chunk (bytes incl. 4-byte size header), samples (int16 array))"""
    z = np.load(os.path.join(HERE, "fast5_chunks.npz"))
    out = []
    vp, zp = 0, 0
    for i, n in enumerate(z["raw_bytes"]):
        vl, zl = int(z["vbz_payload_len"][i]), int(z["zip_payload_len"][i])
        chunk = z["vbz_payload"][vp:vp + vl].tobytes()
        raw = zlib.decompress(z["zip_payload"][zp:zp + zl].tobytes())
        assert len(raw) == n and hashlib.sha256(raw).hexdigest() == str(z["sample_sha256"][i])
        out.append(dict(raw_bytes=int(n), vbz_chunk=chunk, samples=np.frombuffer(raw, np.int16),
                        v1_identical=bool(z["v1_identical"][i])))
        vp += vl
        zp += zl
    return out


def fast5_cd_values():
    z = np.load(os.path.join(HERE, "fast5_chunks.npz"))
    return [int(v) for v in z["cd_values_v0"]], [int(v) for v in z["cd_values_v1"]]


def fuzz_corpus():
    z = np.load(os.path.join(HERE, "fuzz_corpus.npz"))
    out, p = [], 0
    for n in z["lens"]:
        out.append(z["data"][p:p + int(n)].tobytes())
        p += int(n)
    return out
