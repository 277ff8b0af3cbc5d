#!/usr/bin/env python3
"""Emulator fuzz (no GPU needed; tests/cpusim build): libzstd frames of six kinds of input at levels 1-19, six per batch, ONE context for the whole run (stale state between batches is what it is after): every frame must decode to its input; prints the batches that needed the general kernel.
usage: fuzz_decode.py SEED SECONDS"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, product, oracle_lib as ol, ctypes
from oracle_lib import Options
capi, batch, lib = product.get("cpusim")
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def gen():
    kind=rng.integers(0,6)
    n=int(rng.integers(1,400_000))
    if kind==0:   # small-period patterns (offset 1..7 matches) with noise
        per=int(rng.integers(1,8)); base=rng.integers(0,256,per,dtype=np.uint8)
        a=np.tile(base,n//per+1)[:n].copy()
        idx=rng.integers(0,n,max(1,n//int(rng.integers(20,400)))); a[idx]=rng.integers(0,256,idx.size)
        return a.tobytes()
    if kind==1:   # random copies at random distances
        a=bytearray(rng.integers(0,256,n,dtype=np.uint8).tobytes()); pos=300
        while pos<n-300:
            m=int(rng.integers(3,250)); back=int(rng.integers(1,min(pos,70000)))
            for q in range(m): a[pos+q]=a[pos+q-back]
            pos+=m+int(rng.integers(0,300))
        return bytes(a)
    if kind==2:   # stage-1 bytes of signal
        import data_gen
        return ol.stage1_encode(data_gen.nanopore_like(max(n//2,10), rng),2,True)
    if kind==3:   # long runs
        parts=[]
        while sum(map(len,parts))<n:
            parts.append(bytes([int(rng.integers(0,256))])*int(rng.integers(1,5000))); parts.append(rng.integers(0,256,int(rng.integers(0,40)),dtype=np.uint8).tobytes())
        return b"".join(parts)[:n]
    if kind==4:   # text-like
        words=[bytes(rng.integers(97,123,int(rng.integers(2,9)),dtype=np.uint8)) for _ in range(200)]
        out=[]; tot=0
        while tot<n:
            w=words[int(rng.integers(0,200))]; out.append(w+b" "); tot+=len(w)+1
        return b"".join(out)[:n]
    return rng.integers(0,4,n,dtype=np.uint8).tobytes()  # low-entropy
t0=time.time(); it=0; bad=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 120:
    datas=[gen() for _ in range(6)]
    level=int(rng.choice([1,1,2,3,4,5,7,9,12,16,19]))
    frames=[ol.zstd_compress_lib(d,level) for d in datas]
    res=batch.decompress(frames,[len(d) for d in datas],Options(False,0,1,0),sized=False)
    for d,(code,blob) in zip(datas,res):
        if code!=len(d) or blob!=d:
            bad+=1; print("MISMATCH level",level,len(d),hex(code)); open('/tmp/bad_%d.bin'%bad,'wb').write(d)
    out=(ctypes.c_uint32*4)(); lib.vbz_b200_decode_stats(batch.ctx,out)
    if out[1]!=0: print("general kernel used", level, [len(d) for d in datas], list(out))
    it+=1
print("iterations",it,"bad",bad)
