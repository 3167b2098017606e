#!/usr/bin/env python
"""Differential fuzzing of every voxel mode on the CPU emulation of the C ABI (tests/emu/): adversarial grids (1-pixel
axes, 24 bins), windows of 1-9 events, few distinct / unsorted / far-apart timestamps, one-pixel event streams, maps that
send every pixel to one cell, mostly outside the grid, NaN / inf / border entries or integer coordinates.  Checks:
per-bin counts equal everywhere and to the oracle's; TILED == GLOBAL, BANDED == BANDED2 == FACTORED, EXACT == the oracle,
bit for bit; GLOBAL and FACTORED inside the raw-grid bound.  usage: emu_fuzz.py [seed] [seconds]   (CPU only)
Round 1: seeds 1 and 2, 1 560 cases, no failure.  Round 2 (final kernels: 32 x 16 gather tiles, even-aligned boxes,
side-stream plans): seeds 11 and 12, 600 s each, 4 434 cases, no failure."""
import sys, ctypes, numpy as np, time
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'tests'), os.path.join(ROOT,'tests','emu')]
import build_emu, test_emu_abi as T
from oracle import cmda_oracle as O
from cmda_b200 import synth
L=T._bind(build_emu.build_abi())
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+120
it=0; fails=0
while time.time()<t_end:
    it+=1
    H=int(rng.choice([1,2,3,int(rng.integers(4,40)),int(rng.integers(40,130))]))
    W=int(rng.choice([1,2,3,5,int(rng.integers(4,60)),int(rng.integers(60,200))]))
    B=int(rng.choice([1,2,3,5,9,24]))
    n=int(rng.choice([1,2,7,8,9,int(rng.integers(10,400)),int(rng.integers(400,9000))]))
    t,x,y,p=synth.make_events(n,H,W,seed=int(rng.integers(1<<30)))
    kind=rng.integers(0,5)
    if kind==1: t=np.sort(rng.integers(0,3,size=n)).astype(np.uint32)+7           # very few distinct timestamps
    if kind==2: t=rng.permutation(t)                                               # unsorted
    if kind==3: t=(t.astype(np.int64)*80000 % (1<<32)).astype(np.uint32); t.sort() # large spans
    if kind==4: x[:]=x[0]; y[:]=y[0]                                               # one pixel
    m=synth.make_rectify_map(H,W,seed=int(rng.integers(1<<20)))
    mk=rng.integers(0,5)
    if mk==1: m[...,0]=m[0,0,0]; m[...,1]=m[0,0,1]                                 # every pixel -> one cell (overflow lists)
    if mk==2: m=m*np.float32(1.7)-np.float32(3.3)                                  # much of it outside
    if mk==3:
        idx=rng.integers(0,H*W,size=max(1,H*W//10)); mm=m.reshape(-1,2); mm[idx,rng.integers(0,2,size=idx.size)]=rng.choice(np.array([np.nan,np.inf,-np.inf,-0.5,-1.0,W-1,W,H-0.5,1e9,-1e9],np.float32),size=idx.size)
    if mk==4: m=np.round(m)                                                        # integer coordinates
    maps=m[None].astype(np.float32)
    S=int(rng.integers(1,4))
    starts=np.sort(rng.integers(0,n,size=S)); fins=np.minimum(starts+rng.integers(-1,n,size=S),n-1)
    try:
        raws={}
        for mode in (T.GLOBAL,T.TILED,T.FACTORED,T.BANDED,T.BANDED2,T.EXACT):
            if mode==T.TILED and not L.cmda_events_vg_workspace_bytes(n,S,H,W,B,mode): continue
            try:
                raws[mode]=T._vg_batch(L,t,x,y,p,starts,fins,maps,None,H,W,B,mode,normalize=0)
            except AssertionError as e:
                if b'unsupported' in str(e).encode() or 'unsupported' in str(e): continue
                raise
        g,gc=raws[T.GLOBAL]
        for mode,(r,c) in raws.items():
            assert np.array_equal(c,gc),("counts",mode)
        if T.TILED in raws: assert np.array_equal(T.bits(raws[T.TILED][0]),T.bits(g)),"tiled!=global"
        for md in (T.BANDED,T.BANDED2):
            if md in raws and T.FACTORED in raws: assert np.array_equal(T.bits(raws[md][0]),T.bits(raws[T.FACTORED][0])),("banded!=factored",md)
        for s in range(S):
            if fins[s]<starts[s]:
                for mode,(r,c) in raws.items(): assert not r[s].any()
                continue
            sl=slice(int(starts[s]),int(fins[s])+1)
            tf,xf,yf,pf=O.rectify_events(t[sl],x[sl],y[sl],p[sl],maps[0])
            ref,aux=O.events_to_voxel_grid(tf,xf,yf,pf,W,H,B,return_aux=True)
            if T.EXACT in raws: assert np.array_equal(T.bits(raws[T.EXACT][0][s]),T.bits(ref)),"exact!=oracle"
            tol=1e-5*np.maximum(np.abs(ref),aux["abs_weight_sum"])+aux["n_contrib"]*2.0**-31
            for mode in (T.GLOBAL,T.FACTORED):
                if mode in raws:
                    err=np.abs(raws[mode][0][s].astype(np.float64)-ref.astype(np.float64))
                    assert np.all(err<=tol),("tolerance",mode,float(np.max(err-tol)))
            assert np.array_equal(gc[s],aux["bin_counts"]),"bin counts vs oracle"
    except Exception as e:
        fails+=1
        print("FAIL",it,dict(H=H,W=W,B=B,n=n,kind=int(kind),mk=int(mk),S=S,starts=starts.tolist(),fins=fins.tolist()),repr(e)[:300],flush=True)
        if fails>5: break
print("iterations",it,"fails",fails)
