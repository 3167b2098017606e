#!/usr/bin/env python
"""Phase times (cmda_profiler_attach) of the C2 step for the library named by CMDA_B200_LIB: a lean probe for kernel-shape
sweeps (tools/build_variant.sh).  usage: phase_times.py [--bins B] [--mode M] [--store p4|soa] [--steps K]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cmda_b200

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=5)
ap.add_argument("--mode", default="auto")
ap.add_argument("--store", default="p4")
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
L = cmda_b200.lib()
t, x, y, p, rmap, starts, fins = bench.make_workload(16, 5_000_000, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev, plan=False)
if a.store == "p4":
    store = cmda_b200.PackedEventStore.from_event_store(store, plan=False)
out = torch.empty((16, a.bins, bench.H, bench.W), dtype=torch.float32, device=dev)
for _ in range(3):
    cmda_b200.events_vg_batch(store, starts, fins, a.bins, mode=a.mode, out=out)
torch.cuda.synchronize()
n_phase = 8
evs = [[L.cmda_event_create() for _ in range(n_phase)] for _ in range(a.steps)]
arrs = [(ctypes.c_void_p * n_phase)(*pe) for pe in evs]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
used = 0
for k in range(a.steps):
    L.cmda_profiler_attach(arrs[k], n_phase)
    cmda_b200.events_vg_batch(store, starts, fins, a.bins, mode=a.mode, out=out)
    used = L.cmda_profiler_detach()
e1.record()
torch.cuda.synchronize()
ph = np.zeros(max(used - 1, 0))
for pe in evs:
    for j in range(used - 1):
        ms = ctypes.c_float()
        L.cmda_event_elapsed_ms(pe[j], pe[j + 1], ctypes.byref(ms))
        ph[j] += ms.value / a.steps
print(os.path.basename(os.environ.get("CMDA_B200_LIB", "shipped")), f"B={a.bins} {a.mode} {a.store}", round(e0.elapsed_time(e1) / a.steps, 3),
      [round(float(v), 3) for v in ph], "checksum", float(out.double().abs().sum()))
