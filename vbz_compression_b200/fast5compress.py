"""Command line of the fast5 transcoder.  It answers to the flags of the reference's demo script (python/fast5compress/fast5vbz.py:58-76:
input files, -d, -s SUFFIX, --vbz-version, -v) and, like it, writes `<file><suffix>` next to every input and prints that name - but a
whole file's signal goes through the GPU codec as one batch (fast5.transcode), and the result is compacted.

    python -m vbz_compression_b200.fast5compress [-d] [-s SUFFIX] [--vbz-version {0,1}] [--chunk N] [--keep-holes] [--device K] files..."""

import argparse

from . import fast5
from .api import Batch

__version__ = "0.1.1"  # the interface version of the script this stands in for

_FLAGS = (
    (("-d", "--decompress"), dict(action="store_true", help="VBZ -> gzip level 1 instead of gzip -> VBZ")),
    (("-s", "--output-suffix"), dict(default=".tmp", metavar="SUFFIX", help="appended to each input name (default .tmp)")),
    (("--vbz-version",), dict(type=int, default=1, choices=(0, 1), help="chunk format written (cd_values[0])")),
    (("--chunk",), dict(type=int, default=None, metavar="N", help="samples per chunk; default: one chunk per read")),
    (("--keep-holes",), dict(action="store_true", help="leave the space of the old chunks unused instead of compacting the file")),
    (("--device",), dict(type=int, default=0, metavar="K", help="CUDA device of the batch")),
)


def _arguments(argv):
    ap = argparse.ArgumentParser(prog="fast5compress", description=__doc__.split("\n\n")[0])
    ap.add_argument("files", nargs="*", help="fast5 files to re-compress")
    ap.add_argument("-v", "--version", action="version", version=__version__)
    for names, spec in _FLAGS:
        ap.add_argument(*names, **spec)
    return ap.parse_args(argv)


def main(argv=None):
    args = _arguments(argv)
    batch = Batch(args.device)
    try:
        for name in args.files:
            written = fast5.compress_fast5(name, args.output_suffix, batch, vbz_version=args.vbz_version, decompress=args.decompress,
                                           compact_file=not args.keep_holes, chunk_elems=args.chunk)
            print(written)
    finally:
        batch.close()


if __name__ == "__main__":
    main()
