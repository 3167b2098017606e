#!/usr/bin/env python
"""Does splitting the C2 step's windows over k streams (k calls of 16 / k windows, each with its own workspace) let the
gather / normaliser of one group run under the RED kernel of another?  Prints ms per step for k = 1, 2, 4 and for the
same k calls issued on ONE stream (the control: what the split costs without overlap)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cmda_b200

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=5)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--store", default="p4")
a = ap.parse_args()
dev = torch.device("cuda:0")
t, x, y, p, rmap, starts, fins = bench.make_workload(16, 5_000_000, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev, plan=False)
if a.store == "p4":
    store = cmda_b200.PackedEventStore.from_event_store(store, plan=False)
out = torch.empty((16, a.bins, bench.H, bench.W), dtype=torch.float32, device=dev)
ref = cmda_b200.events_vg_batch(store, starts, fins, a.bins).clone()
main = torch.cuda.current_stream()


def step(k, streams):
    g = 16 // k
    ev = torch.cuda.Event()
    ev.record(main)
    for j in range(k):
        st = streams[j % len(streams)]
        st.wait_event(ev)
        with torch.cuda.stream(st):
            cmda_b200.events_vg_batch(store, starts[j * g:(j + 1) * g], fins[j * g:(j + 1) * g], a.bins, out=out[j * g:(j + 1) * g])
    for st in streams:
        done = torch.cuda.Event()
        done.record(st)
        main.wait_event(done)


for k in (1, 2, 4, 8):
    for label, streams in (("streams", [torch.cuda.Stream() for _ in range(k)]), ("one stream", [torch.cuda.Stream()])):
        for _ in range(3):
            step(k, streams)
        torch.cuda.synchronize()
        assert torch.equal(out, ref), (k, label)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(a.steps):
            step(k, streams)
        e1.record(main)
        torch.cuda.synchronize()
        print(f"B={a.bins} {a.store} k={k} {label}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step")
