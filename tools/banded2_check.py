#!/usr/bin/env python
"""BANDED2 (the second cut of the BANDED stage A, mode "banded2") in a process of its own: bit-identity against FACTORED
-- also with polarity bytes beyond {0, 1} -- and the time per step of bench.py's C2 workload, as one JSON line.
Its logic was verified on the CPU emulation (tests/test_emu_*.py) but it had not run on hardware when round 1 ended,
so the GPU test and the bench leg that exercise it do so through this script: whatever a first hardware run does to
the CUDA context stays in this process."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cmda_b200

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=bench.EVENTS_PER_WINDOW)
ap.add_argument("--windows", type=int, default=bench.WINDOWS_PER_GPU)
ap.add_argument("--bins", type=int, nargs="+", default=[5, 1])
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
t, x, y, p, rmap, starts, fins = bench.make_workload(a.windows, a.events, seed_base=0)
store = cmda_b200.EventStore(t, x, y, p, rmap, height=bench.H, width=bench.W, device=dev)
p_odd = np.array(p, copy=True)
p_odd[::997] = 3
p_odd[5::4001] = 255
store_odd = cmda_b200.EventStore(store.t, store.x, store.y, p_odd, rmap, height=bench.H, width=bench.W, device=dev)
res = {}
k = min(2, a.windows)
for b in a.bins:
    ref, rc = cmda_b200.events_vg_batch(store, starts[:k], fins[:k], b, mode="factored", return_bin_counts=True)
    got, gc = cmda_b200.events_vg_batch(store, starts[:k], fins[:k], b, mode="banded2", return_bin_counts=True)
    same = bool(torch.equal(ref, got) and torch.equal(rc, gc))
    # polarity bytes beyond {0, 1}: the banded forms flag the window and the fallback recomputes it with the reference's
    # per-event float32 weights; raw grids agree within the raw-grid bar (1e-5 of the largest magnitudes at play)
    ref = cmda_b200.events_vg_batch(store_odd, starts[:k], fins[:k], b, mode="factored", normalize=False)
    got = cmda_b200.events_vg_batch(store_odd, starts[:k], fins[:k], b, mode="banded2", normalize=False)
    same_odd = bool((ref - got).abs().max() <= 1e-5 * max(1.0, float(ref.abs().max())))
    del ref, got
    out = torch.empty((a.windows, b, bench.H, bench.W), dtype=torch.float32, device=dev)
    ms = {}
    for mode in ("factored", "banded", "banded2"):
        for _ in range(3):
            cmda_b200.events_vg_batch(store, starts, fins, b, mode=mode, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(a.steps):
            cmda_b200.events_vg_batch(store, starts, fins, b, mode=mode, out=out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms[mode] = e0.elapsed_time(e1) / a.steps
    n = int((np.asarray(fins) - np.asarray(starts) + 1).sum())
    res[f"bins_{b}"] = {"bit_identical_to_factored": same, "within_1e-5_with_polarity_bytes_beyond_0_1": same_odd,
                        "ms_per_step": ms, "Mevents_per_s_banded2": n / (ms["banded2"] * 1e-3) / 1e6,
                        "windows": a.windows, "events_per_window": a.events}
print(json.dumps(res), flush=True)
