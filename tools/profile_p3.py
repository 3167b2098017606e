#!/usr/bin/env python
"""The command ncu wraps for the P3 wire's unpack kernel: 10 M events (one copy group of the e2e pipeline) of 3-byte records
-> P4 records on the device, three times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import cmda_b200
from cmda_b200 import _lib, packed

t, x, y, p, rmap, starts, fins = bench.make_workload(2, 5_000_000, seed_base=0)
n = len(t)
rec3, sub, _ = packed.pack_p3(t, x, y, p)
rec4, _, _ = packed.pack_p4(t, x, y, p)
dev = torch.device("cuda:0")
d3, dsub = torch.from_numpy(rec3).to(dev), torch.from_numpy(sub).to(dev)
out = torch.empty((n,), dtype=torch.int32, device=dev)
L = cmda_b200.lib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(3):
    e0.record()
    _lib.check(L.cmda_unpack_p3_to_p4(_lib.ptr(d3), _lib.ptr(dsub), 0, len(sub) - 2, 0, n, _lib.ptr(out), _lib.stream_ptr(dev)), "unpack")
    e1.record()
    torch.cuda.synchronize()
print(f"{n} events: {e0.elapsed_time(e1) * 1e3:.1f} us, {7 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.0f} GB/s of 3 + 4 bytes per event",
      "bit-identical to pack_p4:", bool(np.array_equal(out.cpu().numpy().view(np.uint32), rec4)))
