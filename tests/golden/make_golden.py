"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Every ``*.npz`` written here holds seeded synthetic inputs together with the outputs
of the reference's own functions on them (loaded by ``ref_loader`` without copying
any source into the repo; 1 intra-op thread + deterministic algorithms).  The
fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from cmda_b200 import synth  # noqa: E402

# ISR parameter sets present in the reference tree
ISR_PARAM_SETS = {
    "dsec_default": dict(val_range=(1, 10 ** 2), _threshold=0.04, _clip_range=0.2, shift_pixel=3),      # dsec.py:177
    "cs_day": dict(val_range=(1, 10), _threshold=0.03, _clip_range=0.2, shift_pixel=3),                 # cityscapes_ic.py:99-115
    "cs_new_day": dict(val_range=(1e-5, 255 + 1e-5), _threshold=0, _clip_range=0.040, shift_pixel=3),   # cityscapes_ic.py:99-115
    "dz_new_night": dict(val_range=(500, 1000), _threshold=0.02, _clip_range=0.12, shift_pixel=3),      # dark_zurich_ic.py:110-126
    "cs2dsec_shipped": dict(val_range=[0.01, 1.01], _threshold=0.005, _clip_range=0.1, shift_pixel=1),  # config:46-49
    "cs2dz_shipped": dict(val_range=[1, 100], _threshold=0.01, _clip_range=0.1, shift_pixel=3),         # cs2dz config:42-45
}
DIRECTIONS = ["rightdown", "rightup", "leftdown", "leftup", "all"]


def float_events(n, width, height, seed, same_t=False, oob=True):
    """float32 (time, x, y, pol) with coordinates that spill outside the grid."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.random(n)).astype(np.float32) * np.float32(0.05) + np.float32(3.0)
    if same_t:
        t[:] = t[0]
    lo, hi = (-1.75, 1.5) if oob else (0.0, -1.0)
    x = rng.uniform(lo, width + hi, size=n).astype(np.float32)
    y = rng.uniform(lo, height + hi, size=n).astype(np.float32)
    # a handful of exactly-integer and exactly-negative-fraction coordinates (Q1)
    if n >= 8:
        x[:4] = [0.0, -0.3, width - 1.0, width - 0.5]
        y[:4] = [-1.4, 0.0, height - 1.0, -0.999]
    pol = rng.integers(0, 2, size=n).astype(np.float32)
    return t, x, y, pol


def gen_voxel(out):
    fn = ref_loader.dsec_functions()
    cases = {
        "b5": dict(n=6000, width=64, height=48, bins=5, seed=11),
        "b1": dict(n=4000, width=64, height=48, bins=1, seed=12),
        "b3": dict(n=3000, width=40, height=24, bins=3, seed=13),
        "b2_dense": dict(n=20000, width=16, height=12, bins=2, seed=14),
        "one_event": dict(n=1, width=32, height=24, bins=5, seed=15),
        "same_t": dict(n=500, width=32, height=24, bins=5, seed=16, same_t=True),
        "same_t_b1": dict(n=500, width=32, height=24, bins=1, seed=17, same_t=True),
    }
    for name, c in cases.items():
        t, x, y, pol = float_events(c["n"], c["width"], c["height"], c["seed"], same_t=c.get("same_t", False))
        if name == "one_event":
            x[:] = 5.25
            y[:] = 7.5
        grid = fn["events_to_voxel_grid"](torch.from_numpy(t), torch.from_numpy(x), torch.from_numpy(y),
                                          torch.from_numpy(pol), c["width"], c["height"], c["bins"]).numpy()
        out[f"voxel_{name}"] = dict(time=t, x=x, y=y, pol=pol, width=c["width"], height=c["height"],
                                    bins=c["bins"], grid=grid)


def gen_norm(out):
    fn = ref_loader.dsec_functions()
    rng = np.random.default_rng(21)
    t, x, y, pol = float_events(6000, 64, 48, 11)
    raw = fn["events_to_voxel_grid"](torch.from_numpy(t), torch.from_numpy(x), torch.from_numpy(y),
                                     torch.from_numpy(pol), 64, 48, 5).numpy()
    grids = {
        "raw_b2": raw[1:3].copy(),
        "zeros": np.zeros((1, 24, 32), np.float32),
        "all_pos": np.abs(raw[:1]) + np.float32(0.0),
        "all_neg": -np.abs(raw[:2]),
        "gauss": (rng.normal(0, 3, size=(2, 20, 30)) * (rng.random((2, 20, 30)) < 0.4)).astype(np.float32),
        "single_nonzero": np.where(np.arange(24 * 32).reshape(1, 24, 32) == 100, np.float32(2.5), np.float32(0)).astype(np.float32),
    }
    k = 0
    for gname, g in grids.items():
        out[f"normgrid_{gname}"] = dict(events=g)      # each input grid stored once
        for clip in (0.018, 0.75, 15.0):
            for enforce in (True, False):
                res = fn["events_norm"](torch.from_numpy(g.copy()), clip_range=clip, final_range=1.0,
                                        enforce_no_events_zero=enforce).numpy()
                out[f"norm_{k:02d}_{gname}"] = dict(grid=gname, clip_range=clip, final_range=1.0,
                                                    enforce=int(enforce), result=res)
                k += 1


def gen_events_vg(out):
    cases = {
        "w64_b5": dict(n=8000, width=64, height=48, bins=5, seed=31, start=100, finish=7400),
        "w64_b1": dict(n=8000, width=64, height=48, bins=1, seed=32, start=0, finish=7999),
        "w64_b1_clip": dict(n=3000, width=64, height=48, bins=1, seed=33, start=5, finish=2500, clip=(0.6, 0.6)),
        "one_event": dict(n=100, width=64, height=48, bins=5, seed=34, start=50, finish=50),
        "skew_b5": dict(n=20000, width=64, height=48, bins=5, seed=35, start=0, finish=19999, skew=0.3),
        "dsec_b1": dict(n=30000, width=640, height=480, bins=1, seed=36, start=10, finish=29990),
    }
    for name, c in cases.items():
        t, x, y, p = synth.make_events(c["n"], c["height"], c["width"], seed=c["seed"], skew=c.get("skew", 0.0))
        rmap = synth.make_rectify_map(c["height"], c["width"], seed=c["seed"] + 500)
        res = ref_loader.get_events_vg(t, x, y, p, rmap, c["width"], c["height"], c["bins"], c["finish"], c["start"],
                                       clip_range=c.get("clip"))
        if c["width"] == 640:
            # the full-size map is 2.4 MB of incompressible floats: store its seed and a
            # digest instead; the test regenerates it with synth and checks the digest
            import hashlib
            stored_map = np.frombuffer(hashlib.sha256(rmap.tobytes()).digest(), dtype=np.uint8)
        else:
            stored_map = rmap
        out[f"vg_{name}"] = dict(t=t, x=x, y=y, p=p, rectify_map=stored_map, map_seed=c["seed"] + 500,
                                 width=c["width"], height=c["height"],
                                 bins=c["bins"], start=c["start"], finish=c["finish"],
                                 clip=np.array(c["clip"] if c.get("clip") else [], dtype=np.float64),
                                 result=res.numpy())


def gen_isr(out):
    utils = ref_loader.utils_module()
    rgb = synth.make_rgb_image(72, 104, seed=41)
    pil = Image.fromarray(rgb, mode="RGB")
    out["isr_input"] = dict(rgb=rgb, gray=np.array(pil.convert("L")))
    for pname, parms in ISR_PARAM_SETS.items():
        lut = np.log(np.arange(256, dtype=np.float32) / 255 * (parms["val_range"][1] - parms["val_range"][0])
                     + parms["val_range"][0]).astype(np.float32)
        for d in DIRECTIONS:
            res = utils.get_image_change_from_pil(pil, width=104, height=72, shift_direction=d, **parms).numpy()
            out[f"isr_{pname}_{d}"] = dict(result=res, lut=lut, val_range=np.array(parms["val_range"], np.float64),
                                           threshold=parms["_threshold"], clip_range=parms["_clip_range"],
                                           shift_pixel=parms["shift_pixel"], direction=d)
    # a flat image: every difference is in the dead zone -> 0/1e-8 paths
    flat = np.full((16, 24, 3), 77, np.uint8)
    res = utils.get_image_change_from_pil(Image.fromarray(flat, mode="RGB"), width=24, height=16,
                                          **ISR_PARAM_SETS["dsec_default"]).numpy()
    out["isr_flat"] = dict(rgb=flat, result=res)
    # get_ic called directly on two different gray images
    now, front = synth.make_frame_pair(40, 56, seed=43)
    res = utils.get_ic(front, now, val_range=(1, 100), threshold=0.04, clip_range=0.2).numpy()
    out["get_ic_direct"] = dict(front=front, now=now, result=res)


def gen_image_change(out):
    fn = ref_loader.image_change_functions(log_add=50, threshold=0.1, clip_range=0.8)
    for k, (h, w) in enumerate([(64, 96), (33, 47)]):
        now, front = synth.make_frame_pair(h, w, seed=51 + k)
        # stretch contrast so that some differences exceed threshold and clip
        now = np.clip((now.astype(np.int32) - 128) * 3 + 128, 0, 255).astype(np.uint8)
        img = fn["get_image_change"](Image.fromarray(now, mode="L"), Image.fromarray(front, mode="L"))
        out[f"ic_pair_{k}"] = dict(now=now, front=front, result=np.array(img),
                                   lut=np.log(np.arange(256, dtype=np.float32) + 50).astype(np.float32))


def gen_index(out):
    t, x, y, p, ms_to_idx, t_offset = synth.make_event_store(200_000, 2_000_000, seed=61)
    rng = np.random.default_rng(62)
    ts = np.sort(rng.integers(-5000, 2_010_000, size=64)).astype(np.int64)
    ts[5] = int(t[1234])          # exactly on an event timestamp
    ts[6] = int(t[-1])            # exactly the last event
    ts[7] = int(t[-1]) + 1        # just past the end
    ts = ts[(ts <= 0) | (ts > int(t[-1])) | (ts // 1000 + 1 < len(ms_to_idx))]
    ts = ts + t_offset
    with tempfile.TemporaryDirectory() as tmp:
        res = ref_loader.images_to_events_index(t, t_offset, ms_to_idx, ts, tmp)
    out["index_table"] = dict(t=t, ms_to_idx=ms_to_idx, t_offset=t_offset, timestamps=ts,
                              result=np.array(res, np.int64))


def main():
    assert ref_loader.available(), "needs /root/reference"
    ref_loader.pin_deterministic()
    groups = {"voxel": gen_voxel, "norm": gen_norm, "events_vg": gen_events_vg, "isr": gen_isr,
              "image_change": gen_image_change, "index": gen_index}
    for gname, gen in groups.items():
        cases = {}
        gen(cases)
        flat = {}
        for cname, d in cases.items():
            for k, v in d.items():
                flat[f"{cname}/{k}"] = np.asarray(v)
        path = os.path.join(HERE, f"{gname}.npz")
        np.savez_compressed(path, **flat)
        print(f"{path}: {len(cases)} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
