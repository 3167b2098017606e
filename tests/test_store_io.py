"""Decoded-sequence cache: format round trip and error behaviour on the CPU."""
import json
import os

import numpy as np
import pytest

from cmda_b200 import store_io, synth


def _sequence(n=5000, H=48, W=64, seed=3):
    t, x, y, p = synth.make_events(n, H, W, seed=seed)
    rmap = synth.make_rectify_map(H, W, seed=seed + 1)
    t_ms = (t.astype(np.int64)) // 1000
    ms_to_idx = np.searchsorted(t_ms, np.arange(int(t_ms[-1]) + 2), side="left").astype(np.int64)
    stamps = np.linspace(int(t[0]) + 2000, int(t[-1]) - 10, 7).astype(np.int64) + 123456
    return t, x, y, p, rmap, ms_to_idx, 123456, stamps


def test_round_trip(tmp_path):
    t, x, y, p, rmap, ms, off, stamps = _sequence()
    d = store_io.save_sequence(str(tmp_path / "seq"), t, x, y, p, ms, off, rmap, stamps)
    assert sorted(os.listdir(d)) == ["images_timestamps.npy", "meta.json", "ms_to_idx.npy", "p.npy", "rectify_map.npy",
                                     "t.npy", "x.npy", "y.npy"]
    for mmap in (True, False):
        seq = store_io.load_sequence(d, mmap=mmap)
        assert seq["t_offset"] == off and (seq["height"], seq["width"]) == rmap.shape[:2]
        for name, ref in (("t", t), ("x", x), ("y", y), ("p", p), ("ms_to_idx", ms), ("rectify_map", rmap),
                          ("images_timestamps", stamps)):
            assert seq[name].dtype == ref.dtype and np.array_equal(np.asarray(seq[name]), ref)
    assert isinstance(store_io.load_sequence(d)["t"], np.memmap)


def test_values_that_do_not_fit_raise(tmp_path):
    t, x, y, p, rmap, ms, off, _ = _sequence(200)
    bad_x = x.astype(np.int64)
    bad_x[5] = 70000
    with pytest.raises(ValueError):
        store_io.save_sequence(str(tmp_path / "a"), t, bad_x, y, p, ms, off, rmap)
    with pytest.raises(ValueError):
        store_io.save_sequence(str(tmp_path / "b"), t, x[:-1], y, p, ms, off, rmap)
    with pytest.raises(ValueError):
        store_io.save_sequence(str(tmp_path / "c"), t, x, y, p, ms, off, rmap[..., 0])
    # int64 inputs that do fit are accepted and stored in the DSEC dtypes
    d = store_io.save_sequence(str(tmp_path / "d"), t.astype(np.int64), x.astype(np.int64), y, p, ms, off, rmap)
    seq = store_io.load_sequence(d)
    assert seq["t"].dtype == np.uint32 and "images_timestamps" not in seq


def test_foreign_or_damaged_directory_raises(tmp_path):
    t, x, y, p, rmap, ms, off, _ = _sequence(200)
    d = store_io.save_sequence(str(tmp_path / "seq"), t, x, y, p, ms, off, rmap)
    meta = json.load(open(os.path.join(d, "meta.json")))
    json.dump(dict(meta, version=99), open(os.path.join(d, "meta.json"), "w"))
    with pytest.raises(ValueError):
        store_io.load_sequence(d)
    json.dump(dict(meta, n_events=meta["n_events"] + 1), open(os.path.join(d, "meta.json"), "w"))
    with pytest.raises(ValueError):
        store_io.load_sequence(d)
    json.dump(meta, open(os.path.join(d, "meta.json"), "w"))
    np.save(os.path.join(d, "x.npy"), x.astype(np.int32))
    with pytest.raises(ValueError):
        store_io.load_sequence(d)


def test_convert_needs_h5py(tmp_path):
    try:
        import h5py  # noqa: F401
        import hdf5plugin  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            store_io.convert_dsec_h5("events.h5", "rectify_map.h5", str(tmp_path / "o"))
    else:
        pytest.skip("h5py present: conversion needs a real DSEC file")
