#!/usr/bin/env python
"""Differential fuzzing ON THE GPU: the randomised differential test of tests/test_gpu_parity.py (every voxel mode against
the oracle on adversarial maps / timestamps / grid sizes) for seeds beyond the suite's, plus, per case, the sensor-space
forms against each other (FACTORED == BANDED == BANDED2 bit for bit, raw and normalised) and the packed P4 store against
the SoA store.  usage: gpu_fuzz.py [first_seed] [seconds]   (one B200; prints the number of cases and failures)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch

import cmda_b200 as cm
import test_gpu_parity as T

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 120.0)
cases = fails = 0
seed = first
while time.time() < t_end:
    try:
        T.test_events_vg_randomised_differential(cm, seed)
        # the sensor-space forms against each other, and the packed store against the SoA store
        rng = np.random.default_rng(5000 + seed)
        H, W = int(rng.integers(8, 200)), int(rng.integers(8, 300))
        n = int(rng.integers(50, 60000))
        B = int(rng.integers(1, 6))
        t = np.sort(rng.integers(0, int(rng.choice([3, 1000, 50000, 4_000_000])), size=n)).astype(np.uint32) + np.uint32(rng.integers(0, 1 << 24))
        x = rng.integers(0, W, size=n).astype(np.uint16)
        y = rng.integers(0, H, size=n).astype(np.uint16)
        p = rng.integers(0, 2, size=n).astype(np.uint8)
        rmap = cm.synth.make_rectify_map(H, W, seed=int(rng.integers(1 << 30)))
        store = cm.EventStore(t, x, y, p, rmap, height=H, width=W, device="cuda:0", plan=bool(seed % 2))
        pst = cm.PackedEventStore.from_event_store(store, plan=bool(seed % 3))
        a, b = sorted(int(v) for v in rng.integers(0, n, size=2))
        starts, fins = [0, a, b], [n - 1, b, b - 1 if seed % 4 == 0 else b]
        base = None
        for st_ in (store, pst):
            for mode in ("factored", "banded", "banded2", "auto"):
                try:
                    out, raw, cnt = cm.events_vg_batch(st_, starts, fins, B, mode=mode, return_raw=True, return_bin_counts=True)
                except cm.CmdaError:
                    continue                      # a grid the banded geometry does not cover
                cur = (T.bits(out), T.bits(raw), cnt.cpu().numpy())
                if base is None:
                    base = cur
                assert all(np.array_equal(u, v) for u, v in zip(cur, base)), (seed, mode, type(st_).__name__)
    except Exception as e:                        # noqa: BLE001
        fails += 1
        print("FAIL seed", seed, repr(e)[:300], flush=True)
    cases += 1
    seed += 1
print(f"gpu_fuzz: seeds {first}..{seed - 1}: {cases} cases, {fails} failures")
