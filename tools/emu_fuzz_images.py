#!/usr/bin/env python
"""Differential fuzzing of both pseudo-event generators and of events_norm on the CPU emulation of the C ABI (tests/emu/)
against the oracle: 1-pixel axes, shifts from 0 to the image size, every direction, random / constant / two-level / ramp
images, the parameter sets of the reference plus extremes (threshold 0, clip 2.0, log_add 0.5), grids that are all zero,
all positive, all negative or almost empty.  Pseudo-events must match bit for bit, events_norm within 1e-5.
usage: emu_fuzz_images.py <seed> <seconds>   (CPU only).  Round 1: seed 1, 3 674 cases, no failure."""
import sys, os, numpy as np, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/emu']
import build_emu, test_emu_abi as T
from oracle import cmda_oracle as O
from cmda_b200 import image_change as ic
L=T._bind(build_emu.build_abi())
rng=np.random.default_rng(int(sys.argv[1])); t_end=time.time()+float(sys.argv[2]); it=fails=0
while time.time()<t_end:
    it+=1
    H=int(rng.choice([1,2,3,int(rng.integers(4,50))])); W=int(rng.choice([1,2,3,4,5,int(rng.integers(6,90))]))
    S=int(rng.integers(1,4))
    kind=rng.integers(0,4)
    if kind==0: imgs=rng.integers(0,256,size=(S,H,W),dtype=np.uint8)
    elif kind==1: imgs=np.full((S,H,W),int(rng.integers(0,256)),dtype=np.uint8)
    elif kind==2: imgs=(rng.integers(0,2,size=(S,H,W))*255).astype(np.uint8)
    else: imgs=np.clip(np.add.outer(np.arange(H)*3,np.arange(W)*2)[None]+rng.integers(-2,3,size=(S,H,W)),0,255).astype(np.uint8)
    vr=[(1,100),(0.01,1.01),(1e-5,255+1e-5),(500,1000),(1,10)][int(rng.integers(0,5))]
    thr_f=float(rng.choice([0.0,0.005,0.04,0.3])); clip_f=float(rng.choice([0.04,0.1,0.2,0.9]))
    shift=int(rng.integers(0,min(H,W)+1)); name=list(T.DIRECTIONS)[int(rng.integers(0,5))]
    try:
        lut=ic.log_lut_val_range(tuple(float(v) for v in vr)); span=np.log(vr[1])-np.log(vr[0])
        thr,clip=np.float32(span*thr_f),np.float32(span*clip_f)
        out=np.full((S,1,H,W),np.nan,np.float32); need=L.cmda_image_workspace_bytes(S,H,W,1); ws=T.workspace(need)
        rc=L.cmda_isr_shift_u8(T.ptr(imgs),1,S,H,W,shift,T.DIRECTIONS[name],T.ptr(lut),float(thr),float(clip),T.ptr(out),T.ptr(ws),need,None)
        assert rc==0,rc
        for s in range(S):
            want=O.get_image_change_from_pil(imgs[s],W,H,shift_pixel=shift,val_range=vr,_threshold=thr_f,_clip_range=clip_f,shift_direction=name)
            assert np.array_equal(T.bits(out[s]),T.bits(want)),("isr",s)
        front=rng.integers(0,256,size=(S,H,W),dtype=np.uint8) if rng.random()<0.7 else imgs.copy()
        la=float(rng.choice([50,1,0.5])); th=float(rng.choice([0.1,0.0,0.5])); cr=float(rng.choice([0.8,0.05,2.0]))
        lut2=ic.log_lut_log_add(la); f32=np.full((S,H,W),np.nan,np.float32); u8=np.zeros((S,H,W),np.uint8)
        rc=L.cmda_logdiff_pair_u8(T.ptr(imgs),T.ptr(front),S,H,W,T.ptr(lut2),float(np.float32(th)),float(np.float32(cr)),T.ptr(f32),T.ptr(u8),T.ptr(ws),need,None)
        assert rc==0
        for s in range(S):
            assert np.array_equal(T.bits(f32[s]),T.bits(O.get_image_change(imgs[s],front[s],log_add=la,threshold=th,clip_range=cr,return_float=True))),("pair f32",s)
            assert np.array_equal(u8[s],O.get_image_change(imgs[s],front[s],log_add=la,threshold=th,clip_range=cr)),("pair u8",s)
        # events_norm on random grids
        V=int(rng.integers(1,4000)); g=rng.normal(0,float(rng.choice([0.01,1,30])),size=V).astype(np.float32)
        g[rng.random(V)<float(rng.choice([0,0.5,0.99,1.0]))]=0
        if rng.random()<0.2: g=np.abs(g)
        if rng.random()<0.2: g=-np.abs(g)
        clipn=np.array([float(rng.choice([0.018,0.75,15.0]))],np.float32); enf=int(rng.integers(0,2))
        gg=g.copy(); need2=L.cmda_events_norm_workspace_bytes(1); ws2=T.workspace(need2)
        assert L.cmda_events_norm_batch(T.ptr(gg),1,V,T.ptr(clipn),1.0,enf,T.ptr(ws2),need2,None)==0
        ref=O.events_norm(g.copy(),clipn[0],1.0,bool(enf))
        assert np.allclose(gg,ref,rtol=0,atol=1e-5,equal_nan=True),("norm",float(np.nanmax(np.abs(gg-ref))))
    except Exception as e:
        fails+=1; print("FAIL",it,dict(H=H,W=W,S=S,kind=int(kind),vr=vr,thr=thr_f,clip=clip_f,shift=shift,dir=name),repr(e)[:300],flush=True)
        if fails>5: break
print("iterations",it,"fails",fails)
