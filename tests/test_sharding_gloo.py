"""N > 1 host logic on CPU: two gloo ranks shard the independent units (event windows), process
their own share (the CPU oracle stands in for the kernels here -- this test is about the plumbing:
no collective on the data path, only the gather of checksums / max-over-ranks timing that bench.py
uses), and together cover every window exactly once."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmda_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cmda_oracle as O
        H, W, B, n_windows, n = 24, 32, 3, 6, 400
        rmap = synth.make_rectify_map(H, W, seed=3)
        sizes = [n + 37 * w for w in range(n_windows)]           # ragged windows
        mine_rr = sharding.shard_round_robin(n_windows, world, rank)
        mine_lpt = sharding.shard_lpt(sizes, world, rank)
        sums = {}
        for w in mine_rr:
            t, x, y, p = synth.make_events(sizes[w], H, W, seed=synth.seed_for(9, w))
            vg = O.get_events_vg(t, x, y, p, rmap, W, H, B, sizes[w] - 1, 0)
            sums[w] = float(np.abs(vg).sum())
        everyone = sharding.gather_objects({"rr": mine_rr, "lpt": mine_lpt, "sums": sums})
        # the timing reduction of bench.py: max over ranks
        tm = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert sorted(u for e in everyone for u in e["rr"]) == list(range(n_windows))
            assert sorted(u for e in everyone for u in e["lpt"]) == list(range(n_windows))
            merged = {}
            for e in everyone:
                assert not set(merged) & set(e["sums"]), "a window was processed twice"
                merged.update(e["sums"])
            # same numbers as a single process doing all windows
            for w in range(n_windows):
                t, x, y, p = synth.make_events(sizes[w], H, W, seed=synth.seed_for(9, w))
                ref = float(np.abs(O.get_events_vg(t, x, y, p, rmap, W, H, B, sizes[w] - 1, 0)).sum())
                assert merged[w] == ref
            assert float(tm) == float(world)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def test_gather_objects_without_group():
    assert sharding.gather_objects({"a": 1}) == [{"a": 1}]
