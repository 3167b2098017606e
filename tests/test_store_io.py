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


def _write_dsec_like_files(tmp_path, cname="zstd", n=60_000, H=48, W=64):
    """events.h5 / rectify_map.h5 / timestamps.txt laid out like a DSEC sequence (create_dsec_dataset_txt.py:14-18,
    dsec.py:287-291): chunked datasets behind the Blosc filter, ms_to_idx as uint64, t_offset as an int64 scalar."""
    import h5_writer
    t, x, y, p = synth.make_events(n, H, W, window_us=1_500_000, t_base=0, seed=11)
    rmap = synth.make_rectify_map(H, W, seed=12)
    ms_to_idx = np.searchsorted(t, np.arange(1502) * 1000, side="left").astype(np.uint64)
    opt = dict(chunks=(8192,), filters=["blosc"], cname=cname)
    ev_path, rm_path, ts_path = (str(tmp_path / f) for f in ("events.h5", "rectify_map.h5", "timestamps.txt"))
    h5_writer.write_h5(ev_path, {"events": {"t": (t, opt), "x": (x, opt), "y": (y, opt), "p": (p, opt)}, "ms_to_idx": ms_to_idx,
                                 "t_offset": np.int64(7_654_321)})
    h5_writer.write_h5(rm_path, {"rectify_map": (rmap, dict(chunks=(16, W, 2), filters=["blosc"], cname=cname))})
    stamps = np.linspace(int(t[0]) + 60_000, int(t[-1]) - 10, 9).astype(np.int64) + 7_654_321
    np.savetxt(ts_path, stamps, fmt="%d")
    return ev_path, rm_path, ts_path, (t, x, y, p, rmap, ms_to_idx.astype(np.int64), 7_654_321, stamps)


@pytest.mark.parametrize("cname", ["zstd", "lz4", "zlib"])
def test_convert_dsec_h5_reads_the_reference_file_layout(tmp_path, cname):
    """f-3: the cache builder on HDF5 files laid out like DSEC's (Blosc-filtered chunked datasets in an `events` group,
    `ms_to_idx`, scalar `t_offset`, `rectify_map`): every array comes back bit for bit, streamed in pieces smaller than
    the datasets; the packed P4 stream of the cache unpacks to the same events."""
    from cmda_b200 import packed
    ev_path, rm_path, ts_path, (t, x, y, p, rmap, ms, off, stamps) = _write_dsec_like_files(tmp_path, cname)
    d = store_io.convert_dsec_h5(ev_path, rm_path, str(tmp_path / "seq"), ts_path, chunk_events=10_007, packed="p3")
    seq = store_io.load_sequence(d)
    assert seq["t_offset"] == off and (seq["height"], seq["width"]) == rmap.shape[:2]
    for name, ref in (("t", t), ("x", x), ("y", y), ("p", p), ("ms_to_idx", ms), ("rectify_map", rmap), ("images_timestamps", stamps)):
        assert seq[name].dtype == ref.dtype and np.array_equal(np.asarray(seq[name]), ref), name
    rec, table, t_base = store_io.load_packed(d)
    t2, x2, y2, p2 = packed.unpack_p4(np.asarray(rec), table, t_base)
    assert np.array_equal(t2, t) and np.array_equal(x2, x) and np.array_equal(y2, y) and np.array_equal(p2, p)
    k = min(len(table), len(ms)) - 1
    assert t_base == 0 and np.array_equal(table[:k], ms[:k])                # the records' bucket table is DSEC's own ms_to_idx
    rec3, sub, table3, t_base3 = store_io.load_packed_wire(d)                # the 3-byte wire form of the same stream
    assert t_base3 == t_base and np.array_equal(table3, table) and np.array_equal(packed.p3_to_p4(np.asarray(rec3), sub), np.asarray(rec))
    assert all(np.array_equal(a, b) for a, b in zip(packed.unpack_p3(np.asarray(rec3), sub, t_base), (t, x, y, p)))


def test_h5lite_slices_groups_and_layouts(tmp_path):
    """The reader itself: group listing, window slices that touch only some chunks (dsec.py:342-345), scalar and
    contiguous datasets, HDF5's own shuffle + deflate pipeline, a stored (memcpy) Blosc frame, and a blosclz stream."""
    import h5_writer
    from cmda_b200 import h5lite
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2 ** 32, size=30_001, dtype=np.uint64).astype(np.uint32)
    b = rng.normal(size=(33, 7)).astype(np.float64)
    path = str(tmp_path / "x.h5")
    h5_writer.write_h5(path, {"g": {"a": (a, dict(chunks=(4096,), filters=["shuffle", "deflate"])),
                                    "m": (a[:5000], dict(chunks=(2048,), filters=["blosc"], cname="memcpy"))},
                              "b": (b, dict(chunks=(8, 4))), "c": np.arange(12, dtype=np.int16).reshape(3, 4), "s": np.float32(2.5)})
    with h5lite.File(path) as f:
        assert f.keys() == ["b", "c", "g", "s"] and f["g"].keys() == ["a", "m"] and "g/a" in f and "g/zz" not in f
        d = f["g/a"]
        assert d.shape == a.shape and d.dtype == np.uint32 and len(d) == a.size
        for lo, hi in ((0, 1), (4095, 4097), (12_345, 29_999), (30_000, 30_001), (7, 7)):
            assert np.array_equal(d[lo:hi], a[lo:hi])
        assert np.array_equal(d[()], a) and np.array_equal(f["g"]["m"][100:4200], a[100:4200])
        assert np.array_equal(np.asarray(f["b"]), b) and np.array_equal(f["b"][5:20], b[5:20])
        assert np.array_equal(f["c"][()], np.arange(12, dtype=np.int16).reshape(3, 4)) and float(f["s"][()]) == 2.5
        with pytest.raises(KeyError):
            f["nope"]
    # blosclz: 3 literals, a 6-byte match at distance 3, 1 literal
    assert h5lite._blosclz_decompress(bytes([0x02]) + b"abc" + bytes([0x80, 0x02, 0x00]) + b"X", 10) == b"abcabcabcX"
    with pytest.raises(ValueError):
        h5lite.File(str(tmp_path / "x.h5"), "w")
    open(str(tmp_path / "junk.h5"), "wb").write(b"not hdf5" * 100)
    with pytest.raises(ValueError):
        h5lite.File(str(tmp_path / "junk.h5"))
