"""Access to the PRODUCT library for tests, in its two builds:

  "gpu"     vbz_compression_b200/libvbz_b200.so  - the shipped CUDA library (needs a B200; -m gpu tests)
  "cpusim"  tests/cpusim/libvbz_b200_cpusim.so   - the SAME kernel + host sources compiled against the
            fiber-based CUDA emulator in tests/cpusim (test infrastructure; lets the CPU suite run the
            kernels' logic in the GPU-less build container).  Never shipped, never a fallback.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle_lib  # noqa: E402
from vbz_compression_b200 import _lib as vlib  # noqa: E402
from vbz_compression_b200 import build as vbuild  # noqa: E402
from vbz_compression_b200.api import Batch  # noqa: E402

_cache = {}


def get(kind):
    """-> (capi, batch): capi has the oracle_lib.VbzLib interface over the vbz_* symbols, batch is api.Batch."""
    if kind not in _cache:
        if kind == "cpusim":
            path = vbuild.build_cpusim()
        else:
            import torch
            if not torch.cuda.is_available():
                pytest.skip("no CUDA device")
            path = vbuild.build_native()
        capi = oracle_lib.VbzLib(path)
        lib = vlib.load(path)
        _cache[kind] = (capi, Batch(0, lib=lib), lib)
    return _cache[kind]


BACKENDS = ["cpusim", pytest.param("gpu", marks=pytest.mark.gpu)]
