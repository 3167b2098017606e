#!/usr/bin/env python
"""Differential fuzzing of the fused post-voxel augmentation (mean over bins, crop, flip, bilinear resize, repeat:
dsec.py:304-319) on the CPU emulation of the C ABI against the reference's torch statements applied to the unfused grid:
random grid / crop / output sizes down to one pixel.  usage: <seed> <seconds>.  Round 1: seed 1, 3 868 cases, no failure."""
import sys, os, numpy as np, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/emu']
import build_emu, test_emu_abi as T
from oracle import cmda_oracle as O
from cmda_b200 import synth
L=T._bind(build_emu.build_abi())
rng=np.random.default_rng(int(sys.argv[1])); t_end=time.time()+float(sys.argv[2]); it=fails=0
while time.time()<t_end:
    it+=1
    H,W=int(rng.integers(2,50)),int(rng.integers(2,70)); B=int(rng.choice([1,2,5])); n=int(rng.integers(50,4000)); S=int(rng.integers(1,4))
    t,x,y,p=synth.make_events(n,H,W,seed=int(rng.integers(1<<30)))
    rmap=synth.make_rectify_map(H,W,seed=int(rng.integers(1<<20)))[None]
    starts=np.sort(rng.integers(0,n-1,size=S)); fins=np.minimum(starts+rng.integers(1,n,size=S),n-1)
    cw,ch=int(rng.integers(1,W+1)),int(rng.integers(1,H+1)); ow,oh=int(rng.integers(1,90)),int(rng.integers(1,70))
    if rng.random()<0.2: ow,oh=cw,ch
    xy=[(int(rng.integers(0,W-cw+1)),int(rng.integers(0,H-ch+1))) for _ in range(S)]; flips=[int(v) for v in rng.integers(0,2,size=S)]
    avg=bool(rng.integers(0,2)) and B>1; rep=int(rng.choice([1,3]))
    try:
        grid,_=T._vg_batch(L,t,x,y,p,starts,fins,rmap,None,H,W,B,T.FACTORED)
        st=np.ascontiguousarray(starts,dtype=np.int64); en=np.ascontiguousarray(fins,dtype=np.int64)+1
        clips=np.array([O.default_clip_range(int(f),int(s)) for s,f in zip(starts,fins)],dtype=np.float32)
        table=np.array([[cx,cy,fl] for (cx,cy),fl in zip(xy,flips)],dtype=np.int32)
        Bo=1 if avg else B
        out=np.full((S,rep*Bo,oh,ow),np.nan,np.float32)
        need=L.cmda_events_vg_augmented_workspace_bytes(int((en-st).sum()),S,H,W,B,T.FACTORED); ws=T.workspace(need)
        rc=L.cmda_events_vg_augmented_batch(T.ptr(t),T.ptr(x),T.ptr(y),T.ptr(p),T.ptr(st),T.ptr(en),S,T.ptr(rmap),None,H,W,B,T.ptr(clips),1.0,1,T.ptr(table),cw,ch,ow,oh,int(avg),rep,T.ptr(out),None,None,T.ptr(ws),need,T.FACTORED,None,None)
        assert rc==0,rc
        for s in range(S):
            g=grid[s]
            if avg: g=g.mean(axis=0,keepdims=True,dtype=np.float32)
            import torch, torch.nn.functional as F
            e=torch.from_numpy(np.ascontiguousarray(g[:, xy[s][1]:xy[s][1]+ch, xy[s][0]:xy[s][0]+cw]))
            if flips[s]: e=e.flip(-1)
            e=F.interpolate(e[None],size=(oh,ow),mode='bilinear',align_corners=False)[0].repeat(rep,1,1).numpy()
            assert np.allclose(out[s],e,rtol=0,atol=1e-5),("aug",float(np.abs(out[s]-e).max()))
    except Exception as ex:
        fails+=1; print("FAIL",it,dict(H=H,W=W,B=B,n=n,S=S,cw=cw,ch=ch,ow=ow,oh=oh,avg=avg,rep=rep),repr(ex)[:300],flush=True)
        if fails>5: break
print("iterations",it,"fails",fails)
