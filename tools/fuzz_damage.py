#!/usr/bin/env python3
"""Emulator fuzz (no GPU needed; tests/cpusim build): batches that mix good chunks with damaged ones (bit flips, truncation, overwritten bytes), reference- and own-written: a good neighbour must decode, an accepted damaged chunk must give exactly what the reference gives.
usage: fuzz_damage.py SEED SECONDS"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, product, oracle_lib as ol, data_gen
from oracle_lib import Options
capi, batch, lib = product.get("cpusim")
ref = ol.ref()
rng=np.random.default_rng(int(sys.argv[1]))
o=Options(True,2,1,0)
t0=time.time(); it=0; bad=0; dmg_ok=0; dmg_err=0
while time.time()-t0 < float(sys.argv[2]):
    n=int(rng.integers(2,9))
    reads=[data_gen.nanopore_like(int(np.exp(rng.uniform(np.log(50),np.log(400_000)))), rng) for _ in range(n)]
    src = ref if rng.random()<0.5 else None
    if src: blobs=[ref.compress(r,o,sized=True)[1] for r in reads]
    else: blobs=[b for _,b in batch.compress(reads,o,sized=True)]
    damaged=set()
    chunks=[]
    for i,b in enumerate(blobs):
        if rng.random()<0.4 and len(b)>12:
            bb=bytearray(b); k=rng.integers(0,3)
            if k==0: bb[int(rng.integers(4,len(bb)))]^=1<<int(rng.integers(0,8))
            elif k==1: bb=bb[:int(rng.integers(5,len(bb)))]
            else:
                p=int(rng.integers(4,len(bb))); bb[p:p+int(rng.integers(1,9))]=rng.integers(0,256,int(rng.integers(1,9)),dtype=np.uint8).tobytes()
            damaged.add(i); chunks.append(bytes(bb))
        else: chunks.append(b)
    res=batch.decompress(chunks,[r.nbytes for r in reads],o,sized=True)
    for i,(r,(code,bl)) in enumerate(zip(reads,res)):
        if i in damaged:
            rc,rb=ref.decompress(chunks[i],o,r.nbytes,sized=True)
            if not ol.is_error(code):
                # the product accepted it: then the reference must have too, with the same bytes
                if ol.is_error(rc) or rb!=bl or rc!=code: bad+=1; print("ACCEPTED DIFFERENTLY", hex(code), hex(rc), len(chunks[i]))
                dmg_ok+=1
            else: dmg_err+=1
        elif code!=r.nbytes or bl!=r.tobytes():
            bad+=1; print("GOOD NEIGHBOUR BROKEN", i, hex(code), r.size, sorted(damaged))
    it+=1
print("iterations",it,"bad",bad,"damaged accepted",dmg_ok,"damaged refused",dmg_err)
