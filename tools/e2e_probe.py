#!/usr/bin/env python
"""Probe: end-to-end (pinned host -> device -> pinned host) time of the C2 step for different pipeline group sizes,
next to the raw PCIe copy time of the same bytes.  Sizes the host-buffer front door; not a bench number."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cmda_b200.pipeline import HostEventsPipeline

dev = torch.device("cuda:0")
S, B = 16, 5
t, x, y, p, rmap, starts, fins = bench.make_workload(S, 5_000_000, seed_base=0)
host_out = torch.empty((S, B, bench.H, bench.W), dtype=torch.float32).pin_memory()
for g in (1, 2, 4, 8):
    pipe = HostEventsPipeline(t, x, y, p, rmap, B, bench.H, bench.W, device=dev, windows_per_group=g, max_window_events=5_000_000)
    for _ in range(2):
        pipe(starts, fins, out=host_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pipe(starts, fins, out=host_out)
    torch.cuda.synchronize()
    print("group=%d : %.2f ms per step" % (g, (time.perf_counter() - t0) / 5 * 1e3))
    del pipe
# raw copies
hp = [torch.empty(80_000_000 * k, dtype=torch.uint8).pin_memory() for k in (4, 2, 2, 1)]
dp = [torch.empty_like(h, device=dev) for h in hp]
do = torch.empty((S, B, bench.H, bench.W), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for both in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            for h, d in zip(hp, dp):
                d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                host_out.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    print("raw H2D 720 MB%s : %.2f ms" % (" + D2H 98 MB concurrently" if both else "", (time.perf_counter() - t0) / 5 * 1e3))
