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

    <dir>/rec.npy, rec_ms_to_idx.npy, rec_meta.json   the packed P4 stream (optional; cmda_b200.packed: 4 B/event)
    <dir>/rec3.npy, rec3_sub_to_idx.npy               its 3-byte wire form (optional; what HostEventsPipeline(wire="p3") ships)

``convert_dsec_h5`` writes it from a DSEC sequence's ``events.h5`` + ``rectify_map.h5``: through h5py + hdf5plugin
where those are installed, else through ``cmda_b200.h5lite`` (neither package is in this image; the reader is
exercised on HDF5 files written by ``tests/h5_writer.py``, blosc / deflate / shuffle chunks included);
``load_sequence`` / ``DSECEvents.from_cache`` read it back.
"""
from __future__ import annotations

import json
import os

import numpy as np

__all__ = ["save_sequence", "load_sequence", "convert_dsec_h5", "write_packed", "load_packed", "load_packed_wire", "upload", "FORMAT",
           "VERSION"]

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


def _open_h5(path):
    """h5py + hdf5plugin where they are installed (what the reference uses, dsec.py:3-4), else the reader of
    ``cmda_b200.h5lite`` (the subset of HDF5 + the Blosc filter that DSEC's files use)."""
    try:
        import hdf5plugin  # noqa: F401  (registers the blosc filter, dsec.py:3)
        import h5py
        return h5py.File(path, "r")
    except ImportError:
        from . import h5lite
        return h5lite.File(path)


def convert_dsec_h5(events_h5_path, rectify_map_h5_path, out_dir, images_timestamps_path=None, chunk_events=1 << 24,
                    packed=False) -> str:
    """Decode one DSEC sequence (the files of dsec.py:287-291 and create_dsec_dataset_txt.py:14-18) into the cache
    format, once: ``events/{t,x,y,p}``, ``ms_to_idx`` and ``t_offset`` of events.h5, ``rectify_map`` of
    rectify_map.h5, optionally images/timestamps.txt.  The event datasets are streamed ``chunk_events`` at a time
    into memory-mapped ``.npy`` files (a sequence holds ~4 x 10^8 events).  ``packed=True`` also writes the P4
    stream of ``cmda_b200.packed`` (``rec.npy``, 4 bytes per event; its bucket table is DSEC's own ``ms_to_idx``),
    ``packed="p3"`` additionally its 3-byte wire form (``rec3.npy`` + the 16-microsecond bucket table)."""
    os.makedirs(out_dir, exist_ok=True)
    ev, rm = _open_h5(events_h5_path), _open_h5(rectify_map_h5_path)
    try:
        n = int(ev["events/t"].shape[0])
        maps = {name: np.lib.format.open_memmap(os.path.join(out_dir, name + ".npy"), mode="w+", dtype=_DTYPES[name], shape=(n,))
                for name in ("t", "x", "y", "p")}
        for a in range(0, n, int(chunk_events)):
            b = min(a + int(chunk_events), n)
            for name in ("t", "x", "y", "p"):
                src = np.asarray(ev["events/" + name][a:b])
                dst = src.astype(_DTYPES[name])
                if not np.array_equal(src, dst.astype(src.dtype)):
                    raise ValueError(f"events/{name}: values do not fit {np.dtype(_DTYPES[name]).name}")
                maps[name][a:b] = dst
        for m in maps.values():
            m.flush()
        ms_to_idx = np.asarray(ev["ms_to_idx"]).astype(np.int64)
        t_offset = int(np.asarray(ev["t_offset"][()]).reshape(-1)[0])
        rmap = np.ascontiguousarray(np.asarray(rm["rectify_map"]), dtype=np.float32)
    finally:
        for f in (ev, rm):
            if hasattr(f, "close"):
                f.close()
    if rmap.ndim != 3 or rmap.shape[2] != 2:
        raise ValueError("rectify_map is [H, W, 2] (dsec.py:351-353)")
    np.save(os.path.join(out_dir, "ms_to_idx.npy"), ms_to_idx)
    np.save(os.path.join(out_dir, "rectify_map.npy"), rmap)
    ts = None if images_timestamps_path is None else np.loadtxt(images_timestamps_path, dtype="int64")
    if ts is not None:
        np.save(os.path.join(out_dir, "images_timestamps.npy"), np.atleast_1d(ts).astype(np.int64))
    meta = {"format": FORMAT, "version": VERSION, "n_events": n, "t_offset": t_offset, "height": int(rmap.shape[0]),
            "width": int(rmap.shape[1]), "has_images_timestamps": ts is not None, "has_packed": bool(packed)}
    if packed:
        write_packed(out_dir, maps["t"], maps["x"], maps["y"], maps["p"], chunk_events=chunk_events, wire3=packed == "p3")
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return out_dir


def write_packed(path, t, x, y, p, chunk_events=1 << 24, wire3=False) -> None:
    """Add the packed (P4) stream to a cache directory: ``rec.npy`` uint32 [N] and ``rec_ms_to_idx.npy`` int64 (the
    bucket table of the records: first event of every millisecond since ``rec_t_base``, stored in ``rec_meta.json``).
    DSEC's own ``ms_to_idx`` is that table for ``t_base = 0``; it is rebuilt from ``t`` here so that the cache does not
    depend on the file's copy being consistent.  ``wire3``: also ``rec3.npy`` uint8 [3 N] and ``rec3_sub_to_idx.npy``, the
    3-byte wire form of the same stream (same ``t_base``)."""
    from . import packed as _packed
    n = int(t.shape[0])
    t_base = int(t[0]) // 1000 * 1000 if n else 0
    rec = np.lib.format.open_memmap(os.path.join(path, "rec.npy"), mode="w+", dtype=np.uint32, shape=(n,))
    for a in range(0, n, int(chunk_events)):
        b = min(a + int(chunk_events), n)
        rec[a:b] = _packed.pack_p4(t[a:b], x[a:b], y[a:b], p[a:b], t_base=t_base, check=True)[0]
        if a and int(t[a]) < int(t[a - 1]):
            raise ValueError("P4 needs ascending timestamps")
    rec.flush()
    np.save(os.path.join(path, "rec_ms_to_idx.npy"), _packed.ms_table(t, t_base))
    if wire3:
        rec3 = np.lib.format.open_memmap(os.path.join(path, "rec3.npy"), mode="w+", dtype=np.uint8, shape=(3 * n,))
        for a in range(0, n, int(chunk_events)):
            b = min(a + int(chunk_events), n)
            rec3[3 * a:3 * b] = _packed.pack_p3(t[a:b], x[a:b], y[a:b], p[a:b], t_base=t_base, check=True)[0]
        rec3.flush()
        np.save(os.path.join(path, "rec3_sub_to_idx.npy"), _packed.sub_table(t, t_base))
    with open(os.path.join(path, "rec_meta.json"), "w") as f:
        json.dump({"t_base": t_base, "n_events": n, "wire3": bool(wire3)}, f)


def load_packed(path, mmap=True):
    """``(rec, ms_to_idx, t_base)`` of a cache directory written with ``packed=True`` / ``write_packed``."""
    with open(os.path.join(path, "rec_meta.json")) as f:
        meta = json.load(f)
    rec = np.load(os.path.join(path, "rec.npy"), mmap_mode="r" if mmap else None)
    table = np.load(os.path.join(path, "rec_ms_to_idx.npy"))
    if rec.dtype != np.uint32 or rec.shape != (int(meta["n_events"]),) or table[-1] != rec.shape[0]:
        raise ValueError(f"{path}: packed stream disagrees with rec_meta.json")
    return rec, table, int(meta["t_base"])


def load_packed_wire(path, mmap=True):
    """``(rec3, sub_to_idx, ms_to_idx, t_base)`` of a cache directory written with ``packed="p3"``: the ``packed=`` argument
    of ``HostEventsPipeline(wire="p3")`` is its first three items."""
    with open(os.path.join(path, "rec_meta.json")) as f:
        meta = json.load(f)
    if not meta.get("wire3"):
        raise ValueError(f"{path}: no 3-byte wire form in this cache (convert with packed='p3')")
    rec3 = np.load(os.path.join(path, "rec3.npy"), mmap_mode="r" if mmap else None)
    sub = np.load(os.path.join(path, "rec3_sub_to_idx.npy"))
    table = np.load(os.path.join(path, "rec_ms_to_idx.npy"))
    if rec3.dtype != np.uint8 or rec3.shape != (3 * int(meta["n_events"]),) or sub[-1] != meta["n_events"]:
        raise ValueError(f"{path}: wire stream disagrees with rec_meta.json")
    return rec3, sub, table, int(meta["t_base"])


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
