#!/usr/bin/env python
"""Minimal driver for ncu captures of the pseudo-event kernels (bench.py's C3 workload: 32 pairs of 2048x1024 gray
frames; frame pair with f32 + u8 output, then the shift-pair ISR with the shipped cs2dsec parameters).  Never a bench
number: runs under a profiler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cmda_b200
from cmda_b200 import synth

S, H, W = 32, 1024, 2048
pairs = [synth.make_frame_pair(H, W, seed=synth.seed_for(3, s)) for s in range(4)]
now = torch.from_numpy(np.stack([pairs[s % 4][0] for s in range(S)])).cuda()
front = torch.from_numpy(np.stack([pairs[s % 4][1] for s in range(S)])).cuda()
for _ in range(2):
    f32, u8 = cmda_b200.image_change_batch(now, front, want_f32=True, want_u8=True)
    u8_only = cmda_b200.image_change_batch(now, front, want_f32=False, want_u8=True)      # the table-driven uint8 pass
    isr = cmda_b200.isr_batch(now, 1, (0.01, 1.01), 0.005, 0.1, "rightdown")
torch.cuda.synchronize()
print("ok", float(f32.abs().sum()), float(isr.abs().sum()))
