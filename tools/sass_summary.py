#!/usr/bin/env python3
"""Per-kernel instruction census of the built library (cuobjdump -sass): what profiles/rNN_sass_summary.txt holds.
usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vbz_compression_b200", "libvbz_b200.so")
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "VIADD.16", "VIADD", "REDUX", "MATCH", "SHFL", "VOTE", "PRMT", "SHF.", "LOP3", "IMAD", "LDS", "STS", "ATOMS", "LDG", "STG", "ATOMG", "RED.",
         "BAR.SYNC", "WARPSYNC", "NANOSLEEP", "MEMBAR", "UTMALDG", "UTCMMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
print("# SASS of %s (cuobjdump -sass; arch = sm_100a): instruction mnemonics per kernel" % os.path.relpath(lib, ROOT))
print("# UBLKCP = 1-D TMA bulk copy (cp.async.bulk), SYNCS = mbarrier, LDGSTS = cp.async, VIADD.16 = packed 16-bit SIMD add, REDUX = warp reduction, MATCH = match.any,")
print("# ATOMS = shared-memory atomic, NANOSLEEP / *.STRONG.GPU = the spin-wait hand-offs (look-back, executor chains).  No UTMALDG / UTCMMA: no tensor maps, no contraction.")
print()
for name, body in zip(names, re.split(r"Function : \S+", sass)[1:]):
    ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    if not ops:
        continue
    c = collections.Counter()
    for op in ops:
        for w in WATCH:
            if w == "VIADD" and op.startswith("VIADD.16"):
                continue
            if op.startswith(w) or (w.endswith(".") and op.startswith(w[:-1] + ".")):
                c[w] += 1
                break
        if ".STRONG.GPU" in op:
            c["*.STRONG.GPU"] += 1
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"\(.*$", "", short).replace(" ", "")
    print("%-46s %5d instructions  %s" % (short, len(ops), "  ".join("%s %d" % (w, c[w]) for w in WATCH + ["*.STRONG.GPU"] if c[w])))
