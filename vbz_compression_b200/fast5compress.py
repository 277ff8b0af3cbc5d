"""Command line of the fast5 transcoder: the reference's python/fast5compress/fast5vbz.py:58-76 (same arguments, same
output naming, prints the name of every file it wrote), with the signal of a whole file going through the GPU codec as one
batch (fast5.transcode) instead of one h5py dataset at a time.

    python -m vbz_compression_b200.fast5compress [-d] [-s SUFFIX] [--vbz-version {0,1}] [--chunk N] [--keep-holes] files..."""

import argparse

from . import fast5
from .api import Batch

__version__ = "0.1.1"


def main(argv=None):
    parser = argparse.ArgumentParser("fast5compress")
    parser.add_argument("files", nargs="*", help="input files")
    parser.add_argument("-d", "--decompress", action="store_true", default=False)
    parser.add_argument("-s", "--output-suffix", default=".tmp")
    parser.add_argument("-v", "--version", action="version", version=__version__)
    parser.add_argument("--vbz-version", type=int, default=1)
    parser.add_argument("--device", type=int, default=0, help="CUDA device the batch runs on")
    parser.add_argument("--chunk", type=int, default=None, help="samples per chunk (default: one chunk per read, as the reference's script)")
    parser.add_argument("--keep-holes", action="store_true", default=False, help="do not compact the file (leave the space the old chunks held unused, as h5py would)")
    args = parser.parse_args(argv)
    batch = Batch(args.device)
    try:
        for filename in args.files:
            print(fast5.compress_fast5(filename, args.output_suffix, batch, args.vbz_version, args.decompress, compact_file=not args.keep_holes, chunk_elems=args.chunk))
    finally:
        batch.close()


if __name__ == "__main__":
    main()
