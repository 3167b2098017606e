#!/usr/bin/env python
"""Probe: how much of the SM-bound stage B hides under the L2-atomic-bound stage A when two halves of the C2
batch run on two streams (separate workspaces).  Not a bench number; it sizes the auxiliary-stream design."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cmda_b200

dev = torch.device("cuda:0")
S, B = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 5
t, x, y, p, rmap, starts, fins = bench.make_workload(S, 5_000_000, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev)
out = torch.empty((S, B, bench.H, bench.W), dtype=torch.float32, device=dev)

def timed(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

def single():
    cmda_b200.events_vg_batch(store, starts, fins, B, out=out)

print("one call, one stream           : %.3f ms" % timed(single))
for parts in (2, 4, 8):
    for prio in (False, True):
        streams = [torch.cuda.Stream(dev, priority=(-1 if (prio and k % 2 == 1) else 0)) for k in range(parts)]
        n = S // parts
        def multi():
            cur = torch.cuda.current_stream(dev)
            for k, st in enumerate(streams):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    cmda_b200.events_vg_batch(store, starts[k * n:(k + 1) * n], fins[k * n:(k + 1) * n], B, out=out[k * n:(k + 1) * n])
            for st in streams:
                cur.wait_stream(st)
        print("%d calls on %d streams%s : %.3f ms" % (parts, parts, " (odd streams high priority)" if prio else "                            ", timed(multi)))
