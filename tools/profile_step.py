#!/usr/bin/env python
"""Minimal driver for ncu captures: a few device-resident steps of the voxel path (bench.py's
C2 workload, no end-to-end leg, no CPU leg).  Never a bench number: runs under a profiler."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import cmda_b200

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=5)
ap.add_argument("--mode", default="tiled")
ap.add_argument("--events", type=int, default=bench.EVENTS_PER_WINDOW)
ap.add_argument("--windows", type=int, default=bench.WINDOWS_PER_GPU)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--store", default="p4", choices=["p4", "soa"], help="device-resident event format (bench.py's headline uses p4)")
a = ap.parse_args()
dev = torch.device("cuda:0")
t, x, y, p, rmap, starts, fins = bench.make_workload(a.windows, a.events, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev, plan=False)
if a.store == "p4":
    store = cmda_b200.PackedEventStore.from_event_store(store, plan=False)
out = torch.empty((a.windows, a.bins, bench.H, bench.W), dtype=torch.float32, device=dev)
for _ in range(a.steps):
    cmda_b200.events_vg_batch(store, starts, fins, a.bins, mode=a.mode, out=out)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
