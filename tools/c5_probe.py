"""Where the C5 (train-step input path) latency goes: host enqueue time vs device time per part, and a cProfile of
the host side.  python tools/c5_probe.py  (on a B200)"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cmda_b200                                   # noqa: E402
from cmda_b200 import synth                        # noqa: E402

H, W = 480, 640
dev = torch.device("cuda", 0)
n, S = 330_000, 2
ts, xs, ys, ps, starts, fins = [], [], [], [], [], []
for k in range(S):
    t, x, y, p = synth.make_events(n, H, W, seed=synth.seed_for(5, k))
    ts.append(t); xs.append(x); ys.append(y); ps.append(p)
    starts.append(k * n); fins.append((k + 1) * n - 1)
rmap = synth.make_rectify_map(H, W, seed=synth.seed_for(5, 99))
store = cmda_b200.EventStore(np.concatenate(ts), np.concatenate(xs), np.concatenate(ys), np.concatenate(ps), rmap,
                             height=H, width=W, device=dev)
parms = dict(shift_pixel=1, val_range=(0.01, 1.01), _threshold=0.005, _clip_range=0.1)
means = torch.tensor([123.675, 116.28, 103.53], device=dev).view(1, 3, 1, 1)
stds = torch.tensor([58.395, 57.12, 57.375], device=dev).view(1, 3, 1, 1)
ow, oh = 512, 512
g = torch.Generator(device="cpu").manual_seed(7)
mixed = torch.randn((S, 3, oh, ow), generator=g).to(dev)
warp = (torch.rand((S, oh, ow), generator=g) * 255).to(torch.uint8).to(dev)

parts = {
    "events": lambda: cmda_b200.events_vg_augmented_batch(store, starts, fins, 1, crop_xy=[(37, 61), (140, 20)],
                                                          crop_size=(400, 400), out_size=(ow, oh), flips=[1, 0], repeat=3),
    "target_isr": lambda: cmda_b200.isr_batch(warp, parms["shift_pixel"], parms["val_range"], parms["_threshold"],
                                              parms["_clip_range"], "rightdown"),
    "mixed_isr": lambda: cmda_b200.mixed_image_isr(mixed, means, stds, shift_direction="leftup", **parms),
}


def step():
    for f in parts.values():
        f()


for _ in range(5):
    step()
K = 200
for name, f in list(parts.items()) + [("all", step)]:
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        f()
    e1.record()
    w1 = time.perf_counter()          # host enqueue time
    torch.cuda.synchronize()
    print(f"{name:12s} device {e0.elapsed_time(e1) / K * 1e3:8.1f} us/step   host enqueue {(w1 - w0) / K * 1e6:8.1f} us/step")

pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue())
