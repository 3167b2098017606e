"""The voxel path's real host launch code and kernels (launch_factored: plans, stage A, fused gather; launch_norm_apply)
executed on the CPU through the fiber emulation of tests/emu/ and checked against the oracle: raw grids within the
GPU tests' bound of the float64 sum of the reference's weights, normalised grids within 1e-5 of the C port of the
reference, and the BANDED stage A (both cuts) bit-identical to the L2-RED stage A through the whole path.  Kernel and
host LOGIC only -- test infrastructure, not a product path; the `-m gpu` tests remain the gate for the CUDA build."""
import ctypes
import os
import platform
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="tests/emu switches fibers with x86-64 assembly")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from oracle import cmda_oracle as O  # noqa: E402


def _load(path):
    lib = ctypes.CDLL(path)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.emu_events_vg.restype = ci
    lib.emu_events_vg.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, vp, ci, ci, vp, vp, vp]
    return lib


@pytest.fixture(scope="module")
def lib():
    return _load(build_emu.build_vg())


def events_vg(lib, t, x, y, p, starts, fins, maps, mids, H, W, B, banded, normalize=True):
    S = len(starts)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(fins, dtype=np.int64) + 1
    clips = np.array([(int(f) - int(s)) / 500000 * 1.5 for s, f in zip(starts, fins)], dtype=np.float32)   # dsec.py:362
    out = np.full((S, B, H, W), np.nan, dtype=np.float32)
    raw = np.full((S, B, H, W), np.nan, dtype=np.float32)
    bins = np.full((S, B), -1, dtype=np.int64)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    m = None if maps is None else np.ascontiguousarray(maps, dtype=np.float32)
    ids = None if mids is None else np.ascontiguousarray(mids, dtype=np.int32)
    rc = lib.emu_events_vg(ptr(t), ptr(x), ptr(y), ptr(p), ptr(starts), ptr(ends), S, ptr(m), ptr(ids), H, W, B, ptr(clips),
                           int(banded), int(normalize), ptr(out), ptr(raw), ptr(bins))
    assert rc == 0, rc
    return out, raw, bins


def check_normalised(out, raw, t, x, y, p, start, fin, rmap, W, H, bins):
    """The rule of tests/test_gpu_parity.py::check_normalised: events_norm counts voxels with `events != 0`
    (dsec.py:88-93); where ON / OFF events cancel exactly, the reference's sequential float32 sum may leave a residue
    (|v| ~ 1e-8) that it counts and an exact sum does not.  So: within 1e-5 of events_norm applied to OUR raw grid
    always, and within 1e-5 of the reference's output whenever no voxel's zero / non-zero status differs."""
    clip = np.float32((fin - start) / 500000 * 1.5)
    np.testing.assert_allclose(out, O.events_norm(raw.copy(), clip, 1.0, True), rtol=0, atol=1e-5)
    ref_out, ref_raw = C.get_events_vg_batch(t, x, y, p, [start], [fin], rmap, W, H, bins, return_raw=True)
    flipped = (raw == 0) != (ref_raw[0] == 0)
    if not flipped.any():
        np.testing.assert_allclose(out, ref_out[0], rtol=0, atol=1e-5)
    else:
        assert flipped.mean() < 0.01
        assert np.all(np.abs(ref_raw[0][flipped]) < 1e-6) and np.all(np.abs(raw[flipped]) < 1e-6)


@pytest.mark.parametrize("bins", [1, 5])
def test_emulated_voxel_path_against_the_oracle(lib, bins):
    from cmda_b200 import synth
    H, W, n = 480, 640, 150_000
    t, x, y, p = synth.make_events(n, H, W, seed=synth.seed_for(3, bins), skew=0.2)
    t[-2000:] = t[-2000]
    maps = np.stack([synth.make_rectify_map(H, W, seed=5), synth.make_rectify_map(H, W, seed=6, k1=0.03)])
    starts, fins, mids = [0, 1001, 600, n - 1500, 40_000], [n - 2500, 70_000, 599, n - 1, 40_000], [0, 1, 0, 1, 1]
    out, raw, counts = events_vg(lib, t, x, y, p, starts, fins, maps, mids, H, W, bins, banded=0)
    for s in range(len(starts)):
        if fins[s] < starts[s]:                                        # empty window: zero raw grid, events_norm of zeros
            assert not raw[s].any()
            clip = (fins[s] - starts[s]) / 500000 * 1.5
            np.testing.assert_allclose(out[s], O.events_norm(raw[s].copy(), np.float32(clip), 1.0, True), rtol=0, atol=1e-5)
            continue
        sl = slice(starts[s], fins[s] + 1)
        tf, xf, yf, pf = O.rectify_events(t[sl], x[sl], y[sl], p[sl], maps[mids[s]])
        truth, abs_w, n_contrib = O.voxel_grid_f64(tf, xf, yf, pf, W, H, bins, return_aux=True)
        err = np.abs(raw[s].astype(np.float64) - truth)
        tol = 1e-5 * np.maximum(np.abs(truth), abs_w) + n_contrib * 2.0 ** -31
        assert np.all(err <= tol), f"window {s}: raw grid out of tolerance by {np.max(err - tol):.3e}"
        assert np.all(raw[s][n_contrib == 0] == 0.0)
    assert not raw[3].any() and int(counts[3].sum()) == 0              # single timestamp: NaN t_norm, nothing lands (Q3)
    live = [s for s in range(len(starts)) if fins[s] >= starts[s]]
    for s in live:                                                      # one map per oracle call
        check_normalised(out[s], raw[s], t, x, y, p, starts[s], fins[s], maps[mids[s]], W, H, bins)
    # the BANDED stage A, first and second cut: the same R, hence the same bits all the way down
    for cut in (1, 2):
        o2, r2, c2 = events_vg(lib, t, x, y, p, starts, fins, maps, mids, H, W, bins, banded=cut)
        assert np.array_equal(r2.view(np.uint32), raw.view(np.uint32))
        assert np.array_equal(o2.view(np.uint32), out.view(np.uint32))
        assert np.array_equal(c2, counts)


def test_emulated_voxel_path_small_grids_and_no_map(lib):
    from cmda_b200 import synth
    for (H, W), bins in (((37, 53), 3), ((24, 700), 2)):
        n = 20_000
        t, x, y, p = synth.make_events(n, H, W, seed=synth.seed_for(4, H))
        rmap = synth.make_rectify_map(H, W, seed=9)
        starts, fins = [0, 33], [n - 1, 9000]
        for maps in (rmap[None], None):
            out, raw, counts = events_vg(lib, t, x, y, p, starts, fins, maps, None, H, W, bins, banded=0)
            for s in range(2):
                check_normalised(out[s], raw[s], t, x, y, p, starts[s], fins[s], None if maps is None else rmap, W, H, bins)
            for cut in (1, 2):
                o2, r2, c2 = events_vg(lib, t, x, y, p, starts, fins, maps, None, H, W, bins, banded=cut)
                assert np.array_equal(o2.view(np.uint32), out.view(np.uint32)) and np.array_equal(c2, counts)
