#!/usr/bin/env python3
"""Emulator fuzz (no GPU needed; tests/cpusim build): batches of reads of every integer type through compress_batch in ONE context: every chunk must decode with the unmodified reference and with this library.
usage: fuzz_encode.py SEED SECONDS"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, product, oracle_lib as ol, data_gen
from oracle_lib import Options
capi, batch, lib = product.get("cpusim")
ref = ol.ref() or ol.port()
rng=np.random.default_rng(int(sys.argv[1]))
def sig(n, dt):
    k=rng.integers(0,6)
    if k==0: a=data_gen.nanopore_like(n, rng).astype(np.int64)
    elif k==1: a=np.tile(data_gen.nanopore_like(max(1,int(rng.integers(50,3000))), rng), n//50+1)[:n].astype(np.int64)
    elif k==2: a=np.zeros(n,np.int64)+int(rng.integers(-100,100))
    elif k==3: a=np.arange(n,dtype=np.int64)*int(rng.integers(0,5))
    elif k==4: a=rng.integers(-3,4,n).cumsum()
    else: a=rng.integers(-30000,30000,n)
    info=np.iinfo(dt)
    return (a % (int(info.max)-int(info.min)+1) + int(info.min)).astype(dt) if k==5 else np.clip(a, info.min, info.max).astype(dt)
t0=time.time(); it=0; bad=0
while time.time()-t0 < float(sys.argv[2]):
    dt,o = [(np.int16,Options(True,2,1,1)),(np.int16,Options(True,2,1,0)),(np.int8,Options(True,1,1,1)),(np.uint32,Options(False,4,3,0)),(np.int32,Options(True,4,1,0)),(np.uint8,Options(False,1,2,0))][int(rng.integers(0,6))]
    lens=[int(x) for x in np.exp(rng.uniform(0, np.log(300_000), int(rng.integers(1,8))))]
    if rng.random()<0.2: lens.append(0)
    reads=[sig(n,dt) for n in lens]
    res=batch.compress(reads,o,sized=True)
    blobs=[]
    for r,(code,blob) in zip(reads,res):
        if ol.is_error(code): bad+=1; print("compress error",hex(code),r.size,dt); continue
        rc,back=ref.decompress(blob,o,r.nbytes,sized=True)
        if rc!=r.nbytes or back!=r.tobytes(): bad+=1; print("REF cannot decode",dt,r.size,hex(rc)); open('/tmp/badenc_%d.npy'%bad,'wb').write(r.tobytes())
        blobs.append((r,blob))
    back=batch.decompress([b for _,b in blobs],[r.nbytes for r,_ in blobs],o,sized=True)
    for (r,_),(code,bl) in zip(blobs,back):
        if code!=r.nbytes or bl!=r.tobytes(): bad+=1; print("own decode mismatch",dt,r.size,hex(code))
    it+=1
print("iterations",it,"bad",bad)
