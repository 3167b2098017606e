#!/usr/bin/env python
"""Differential fuzzing of the PIL-exact resize (against Pillow, both channel counts, up- and down-scaling) and of the
window slicer (against the oracle, timestamps before, inside and after the stream, IndexError / ValueError cases as
status codes) on the CPU emulation of the C ABI.  usage: <seed> <seconds>.  Round 1: seed 1, 90 s, no failure."""
import sys, os, numpy as np, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, ROOT+'/tests', ROOT+'/tests/emu']
import build_emu, test_emu_abi as T
from oracle import cmda_oracle as O
from cmda_b200 import synth
from PIL import Image
L=T._bind(build_emu.build_abi())
rng=np.random.default_rng(int(sys.argv[1])); t_end=time.time()+float(sys.argv[2]); it=fails=0
while time.time()<t_end:
    it+=1
    try:
        # resize vs Pillow
        H,W=int(rng.integers(1,60)),int(rng.integers(1,80)); oh,ow=int(rng.integers(1,70)),int(rng.integers(1,90)); C=int(rng.choice([1,3])); S=int(rng.integers(1,3))
        img=rng.integers(0,256,size=(S,H,W)+((3,) if C==3 else ()),dtype=np.uint8)
        dst=np.zeros((S,oh,ow)+((3,) if C==3 else ()),dtype=np.uint8)
        need=L.cmda_resize_bilinear_workspace_bytes(S,H,W,C,oh,ow); ws=T.workspace(max(need,1))
        rc=L.cmda_resize_bilinear_u8(T.ptr(img),C,S,H,W,oh,ow,T.ptr(dst),T.ptr(ws),need,None)
        if rc==-4: continue      # more filter taps than the kernel is built for (documented limit)
        assert rc==0,("resize rc",rc)
        for s in range(S):
            want=np.asarray(Image.fromarray(img[s],mode="RGB" if C==3 else "L").resize((ow,oh),Image.BILINEAR))
            assert np.array_equal(dst[s],want),("resize",H,W,oh,ow,C)
        # slicer vs oracle
        n=int(rng.integers(50,20000)); dur=int(rng.integers(3000,200000))
        st=synth.make_event_store(n,dur,seed=int(rng.integers(1<<30)))
        t,ms,toff=st[0],np.ascontiguousarray(st[4],dtype=np.int64),int(st[5])
        nts=int(rng.integers(1,40)); ts=(rng.integers(-2000,dur+3000,size=nts)+toff).astype(np.int64)
        idx=np.full(nts,-9,np.int64); status=np.full(nts,-9,np.int32)
        rc=L.cmda_images_to_events_index(T.ptr(t),t.shape[0],T.ptr(ms),ms.shape[0],toff,T.ptr(ts),nts,T.ptr(idx),T.ptr(status),None); assert rc==0
        for i in range(nts):
            try:
                want=O.images_to_events_index(t,toff,ms,ts[i:i+1])[0]
                assert status[i]==0 and idx[i]==want,("index",i,int(idx[i]),int(want))
            except ValueError:
                assert status[i]==1,("range error expected",i)
            except IndexError:
                assert status[i]==2,("index error expected",i)
    except Exception as e:
        fails+=1; print("FAIL",it,repr(e)[:300],flush=True)
        if fails>5: break
print("iterations",it,"fails",fails)
