#!/usr/bin/env python
"""Probe (GPU box): the C2 step with its windows split across two streams -- the first k windows through FACTORED
(stage A bound by L2 atomics), the rest through BANDED (stage A bound by instruction issue) -- against either mode
alone.  Windows are independent and each (device, stream) pair has its own workspace, so the split is a pure host
composition of two validated paths; the result must be bit-identical to the single-mode output.  Prints ms per step
for every split.  Not a bench number: a design probe (stage A next steps)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import cmda_b200

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=5)
ap.add_argument("--events", type=int, default=bench.EVENTS_PER_WINDOW)
ap.add_argument("--windows", type=int, default=bench.WINDOWS_PER_GPU)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--splits", default="0,4,6,8,10,12,16", help="windows given to FACTORED, comma separated")
ap.add_argument("--second", default="banded", choices=["banded", "banded2"], help="mode of the second stream")
a = ap.parse_args()
dev = torch.device("cuda:0")
t, x, y, p, rmap, starts, fins = bench.make_workload(a.windows, a.events, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev)
S = a.windows
ref = cmda_b200.events_vg_batch(store, starts, fins, a.bins, mode="factored")
out = torch.empty_like(ref)
side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
fork, join = torch.cuda.Event(), [torch.cuda.Event(), torch.cuda.Event()]


def step(k):
    main = torch.cuda.current_stream(dev)
    fork.record(main)
    parts = ((side[0], "factored", slice(0, k)), (side[1], a.second, slice(k, S)))
    for i, (st, mode, sl) in enumerate(parts):
        if sl.stop - sl.start <= 0:
            continue
        st.wait_event(fork)
        with torch.cuda.stream(st):
            cmda_b200.events_vg_batch(store, starts[sl], fins[sl], a.bins, mode=mode, out=out[sl])
            join[i].record(st)
        main.wait_event(join[i])


for k in [int(v) for v in a.splits.split(",")]:
    k = max(0, min(S, k))
    out.zero_()
    for _ in range(3):
        step(k)
    torch.cuda.synchronize(dev)
    same = bool(torch.equal(out, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize(dev)
    print(f"B={a.bins} factored windows {k:2d} | {a.second} windows {S - k:2d}: {e0.elapsed_time(e1) / a.steps:.3f} ms per step, "
          f"bit-identical to FACTORED alone: {same}", flush=True)
