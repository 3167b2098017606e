"""Decoded-sequence cache: the on-disk boundary of the event store.

The reference re-opens ``events.h5`` (blosc-compressed HDF5), ``rectify_map.h5`` and
``images_to_events_index.txt`` for every sample (mmseg/datasets/dsec.py:287-293) and slices the compressed
datasets per window (dsec.py:342-345).  Decoding is file I/O with no arithmetic; what the CUDA path wants
is the decoded SoA arrays in their native dtypes, resident on the device.  This module defines the format
in between: one directory per sequence holding plain ``.npy`` files (memory-mappable, no decompression on
the hot path) and a ``meta.json``::

    <dir>/t.npy            uint32 [N]      events/t   (microseconds, sorted)
    <dir>/x.npy            uint16 [N]      events/x
    <dir>/y.npy            uint16 [N]      events/y
    <dir>/p.npy            uint8  [N]      events/p   (0 / 1)
    <dir>/ms_to_idx.npy    int64  [M]      ms_to_idx
    <dir>/rectify_map.npy  float32 [H,W,2] rectify_map (channel 0 = x, 1 = y; dsec.py:351-353)
    <dir>/images_timestamps.npy  int64 [I] images/timestamps.txt        (optional)
    <dir>/meta.json        {"format": "cmda_b200.sequence", "version": 1, "n_events": N, "t_offset": ..., "height": H, "width": W}

``convert_dsec_h5`` writes it from a DSEC sequence with h5py + hdf5plugin where those are installed (they are
not in this image: the function is exercised here only up to the import, see tests/test_store_io.py);
``load_sequence`` / ``DSECEvents.from_cache`` read it back.
"""
from __future__ import annotations

import json
import os

import numpy as np

__all__ = ["save_sequence", "load_sequence", "convert_dsec_h5", "upload", "FORMAT", "VERSION"]

FORMAT = "cmda_b200.sequence"
VERSION = 1
_DTYPES = {"t": np.uint32, "x": np.uint16, "y": np.uint16, "p": np.uint8, "ms_to_idx": np.int64,
           "rectify_map": np.float32, "images_timestamps": np.int64}


def save_sequence(path, t, x, y, p, ms_to_idx, t_offset, rectify_map, images_timestamps=None) -> str:
    """Write one decoded sequence.  Arrays are converted to the DSEC dtypes; values that do not fit raise."""
    os.makedirs(path, exist_ok=True)
    arrays = {"t": t, "x": x, "y": y, "p": p, "ms_to_idx": ms_to_idx, "rectify_map": rectify_map}
    if images_timestamps is not None:
        arrays["images_timestamps"] = images_timestamps
    out = {}
    for name, a in arrays.items():
        a = np.asarray(a)
        b = np.ascontiguousarray(a, dtype=_DTYPES[name])
        if a.dtype != b.dtype and not np.array_equal(a, b.astype(a.dtype)):
            raise ValueError(f"{name}: values do not fit {np.dtype(_DTYPES[name]).name}")
        out[name] = b
    n = out["t"].shape[0]
    if not (out["t"].ndim == 1 and out["x"].shape == out["y"].shape == out["p"].shape == (n,)):
        raise ValueError("t, x, y, p must be 1-D arrays of one length (dsec.py:28-29)")
    if out["rectify_map"].ndim != 3 or out["rectify_map"].shape[2] != 2:
        raise ValueError("rectify_map is [H, W, 2] (dsec.py:351-353)")
    for name, b in out.items():
        np.save(os.path.join(path, name + ".npy"), b)
    meta = {"format": FORMAT, "version": VERSION, "n_events": int(n), "t_offset": int(t_offset),
            "height": int(out["rectify_map"].shape[0]), "width": int(out["rectify_map"].shape[1]),
            "has_images_timestamps": images_timestamps is not None}
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return path


def load_sequence(path, mmap=True) -> dict:
    """Read a sequence back: a dict with the arrays (memory-mapped unless ``mmap=False``), ``t_offset`` and the
    grid size.  Raises ``ValueError`` for a foreign or newer format, or arrays that disagree with ``meta.json``."""
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT or int(meta.get("version", -1)) > VERSION:
        raise ValueError(f"{path}: not a {FORMAT} v<={VERSION} directory")
    seq = {"t_offset": int(meta["t_offset"]), "height": int(meta["height"]), "width": int(meta["width"])}
    names = ["t", "x", "y", "p", "ms_to_idx", "rectify_map"] + (["images_timestamps"] if meta.get("has_images_timestamps") else [])
    for name in names:
        a = np.load(os.path.join(path, name + ".npy"), mmap_mode="r" if mmap else None)
        if a.dtype != _DTYPES[name]:
            raise ValueError(f"{path}/{name}.npy: dtype {a.dtype}, expected {np.dtype(_DTYPES[name]).name}")
        seq[name] = a
    n = int(meta["n_events"])
    if not (seq["t"].shape == seq["x"].shape == seq["y"].shape == seq["p"].shape == (n,)):
        raise ValueError(f"{path}: event arrays disagree with meta.json (n_events = {n})")
    if seq["rectify_map"].shape != (seq["height"], seq["width"], 2):
        raise ValueError(f"{path}: rectify_map shape {seq['rectify_map'].shape}")
    return seq


def convert_dsec_h5(events_h5_path, rectify_map_h5_path, out_dir, images_timestamps_path=None) -> str:
    """Decode one DSEC sequence (the files of dsec.py:287-291 and create_dsec_dataset_txt.py:14-18) into the cache
    format, once.  Needs h5py and hdf5plugin (blosc filter), like the reference; raises ``ImportError`` naming
    them when they are missing."""
    try:
        import hdf5plugin  # noqa: F401  (registers the blosc filter, dsec.py:3)
        import h5py
    except ImportError as e:
        raise ImportError("convert_dsec_h5 needs h5py and hdf5plugin to decode events.h5 "
                          "(reference mmseg/datasets/dsec.py:3-4); install them or decode elsewhere and call "
                          "save_sequence") from e
    with h5py.File(events_h5_path, "r") as ev, h5py.File(rectify_map_h5_path, "r") as rm:
        ts = None if images_timestamps_path is None else np.loadtxt(images_timestamps_path, dtype="int64")
        return save_sequence(out_dir, ev["events/t"][()], ev["events/x"][()], ev["events/y"][()], ev["events/p"][()],
                             np.asarray(ev["ms_to_idx"], dtype="int64"), int(ev["t_offset"][()]),
                             np.asarray(rm["rectify_map"]), ts)


def upload(a, device, chunk_bytes=256 << 20):
    """Host array (typically a read-only memory map of the cache) -> device tensor of the same dtype, streamed through
    two pinned staging buffers so that the page-cache read of one chunk overlaps the DMA of the previous one and no
    pageable full-size copy is made."""
    import torch
    a = np.asarray(a)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    tdt = {np.dtype(np.uint32): torch.uint32, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint8): torch.uint8,
           np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32}[a.dtype]
    raw = a.reshape(-1).view(np.uint8)
    nbytes = raw.shape[0]
    dst = torch.empty((nbytes,), dtype=torch.uint8, device=device)
    step = max(1, min(int(chunk_bytes), nbytes))
    stages = [torch.empty((step,), dtype=torch.uint8).pin_memory() for _ in range(2 if nbytes > step else 1)]
    done = [None, None]
    with torch.cuda.device(device):
        for k, lo in enumerate(range(0, nbytes, step)):
            hi = min(lo + step, nbytes)
            st = stages[k % len(stages)]
            if done[k % len(stages)] is not None:
                done[k % len(stages)].synchronize()              # the DMA that last read this staging buffer
            np.copyto(st.numpy()[: hi - lo], raw[lo:hi])
            dst[lo:hi].copy_(st[: hi - lo], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            done[k % len(stages)] = ev
        torch.cuda.current_stream().synchronize()
    return dst.view(tdt).reshape(a.shape)
