"""Shared-memory bank simulation of the FAST corner phases: wavefronts per warp-wide LDS.U8 for a list order (lane -> row permutation)
and a row pitch, on random corner sets of the benchmark density.  python tools/bank_sim.py   (no GPU)"""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from hyslam_b200 import synth
rng=np.random.default_rng(1)
def wavefronts(addrs):
    # addrs: byte addresses of 32 lanes (array), returns wavefronts for an LDS.U8 (distinct words per bank, max over banks)
    words=np.unique(addrs//4)
    banks=words%32
    return np.bincount(banks,minlength=32).max()
def sim(perm, pitch=112, off=7, density=0.245, tiles=40, order='lane'):
    tot=0; n=0
    for t in range(tiles):
        flags=rng.random((64,96))<density   # pixel rows x cols
        flags[:3]=False; flags[61:]=False; flags[:,:3]=False; flags[:,93:]=False
        lst=[]
        warps=[(s,b) for b in range(2) for s in range(3)]
        rng.shuffle(warps)
        for (s,b) in warps:
            for lane in range(32):
                row=b*32+perm[lane]
                cols=np.nonzero(flags[row,32*s:32*s+32])[0]
                for c in cols: lst.append(off+row*pitch+32*s+c)
        lst=np.array(lst)
        for i0 in range(0,len(lst),192):
            chunk=lst[i0:i0+192]
            for w0 in range(0,len(chunk),32):
                a=chunk[w0:w0+32]
                tot+=wavefronts(a); n+=1
    return tot/n
ident=list(range(32))
p2=[((l&15)<<1)|(l>>4) for l in range(32)]
p4=[((l&7)<<2)|(l>>3) for l in range(32)]
for name,p in (('ident',ident),('step2',p2),('step4',p4)):
    print(name, 'pix pitch112:', round(sim(p),3), ' pitch 96:', round(sim(p,pitch=96),3), ' pitch 128:', round(sim(p,pitch=128),3), 'pitch 144', round(sim(p,pitch=144),3), 'pitch 160', round(sim(p,pitch=160),3))
