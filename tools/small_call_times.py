#!/usr/bin/env python
"""Per-call time of small batches (the launch-bound end of the path): one window of 1 M events (C1) and two windows of
330 k events (C5's voxel part), B = 5 and B = 1, with the map plans built per call (side stream) and prebuilt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cmda_b200

dev = torch.device("cuda:0")
for label, S, n in (("C1: 1 x 1 M", 1, 1_000_000), ("C5: 2 x 330 k", 2, 330_000)):
    t, x, y, p, rmap, starts, fins = bench.make_workload(S, n, seed_base=3)
    for plan in (False, True):
        soa = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev, plan=plan)
        for store_name, store in (("soa", soa), ("p4", cmda_b200.PackedEventStore.from_event_store(soa, plan=plan))):
            for bins in (5, 1):
                out = torch.empty((S, bins, bench.H, bench.W), dtype=torch.float32, device=dev)
                for _ in range(5):
                    cmda_b200.events_vg_batch(store, starts, fins, bins, out=out)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(200):
                    cmda_b200.events_vg_batch(store, starts, fins, bins, out=out)
                e1.record()
                torch.cuda.synchronize()
                eager = e0.elapsed_time(e1) / 200 * 1e3
                # the same call replayed from a CUDA graph: the device time of the call without the host's launch work
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cmda_b200.events_vg_batch(store, starts, fins, bins, out=out)
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(200):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                print(f"{label}  B={bins}  store={store_name}  plans {'prebuilt' if plan else 'per call'}: {eager:.1f} us per call, "
                      f"{e0.elapsed_time(e1) / 200 * 1e3:.1f} us replayed from a CUDA graph")
