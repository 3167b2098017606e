"""Kernel-logic checks without a GPU: the stage-A kernels of cmda_b200/csrc/voxel_factored.cu -- the L2-RED kernel,
the BANDED pair and its second cut (mode BANDED2) -- are compiled for the host against a fiber-based stand-in for
CUDA (tests/emu/) and run on small windows.  All three must produce the same sensor-space grid R and the same
per-bin event counts, bit for bit, and R must equal a direct numpy restatement of its definition
(R[t0][y][x] += sign * (2^44 + round(f * 2^24)), voxel_factored.cu header).  This is test infrastructure: it shares
no code path with the product (which has no CPU fallback) and proves nothing about speed; the GPU parity tests
(`-m gpu`) remain the gate for the CUDA build."""
import ctypes
import os
import platform
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="tests/emu switches fibers with x86-64 assembly")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    lib = ctypes.CDLL(build_emu.build())
    vp = ctypes.c_void_p
    lib.emu_stage_a.restype = ctypes.c_int
    lib.emu_stage_a.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    return lib


def stage_a(lib, t, x, y, p, starts, ends, H, W, B, variant):
    S = len(starts)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(ends, dtype=np.int64)
    if B == 1:
        R = np.full((S, H, W), -123456, dtype=np.int32)           # garbage: the BANDED kernels must store every cell
    else:
        R = np.full((S, B, H, W), 0x5a5a5a5a5a5a5a5a, dtype=np.int64)
    bins = np.zeros((S, B), dtype=np.uint64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emu_stage_a(ptr(t), ptr(x), ptr(y), ptr(p), ptr(starts), ptr(ends), S, H, W, B, variant, ptr(R), ptr(bins))
    assert rc == 0
    return R, bins


def direct_R(t, x, y, p, start, end, H, W, B):
    """The definition of R for one window (float32 arithmetic of dsec.py:347-348 and 38-43 for the time)."""
    R = np.zeros((B, H, W), dtype=np.int64)
    cnt = np.zeros(B, dtype=np.uint64)
    if end <= start:
        return R, cnt
    tt, xx, yy, pp = t[start:end], x[start:end].astype(np.int64), y[start:end].astype(np.int64), p[start:end]
    dT = np.float32(np.uint32(tt[-1] - tt[0]))
    if not dT > 0:
        return R, cnt                                              # 0 / 0: every t_norm is NaN
    with np.errstate(over="ignore"):
        dt = (tt - tt[0]).astype(np.uint32).astype(np.float32)
    tn = np.float32(B - 1) * (dt / dT)
    tb = np.trunc(tn).astype(np.int64)
    ok = (xx < W) & (yy < H) & (tb >= 0) & (tb < B)
    f = (tn - tb.astype(np.float32)).astype(np.float32)
    fq = np.rint(f * np.float32(16777216.0)).astype(np.int64)
    sign = 2 * pp.astype(np.int64) - 1                              # dsec.py:45 on whatever the polarity byte holds
    val = sign * ((1 << 44) + fq) if B > 1 else sign
    np.add.at(R, (tb[ok], yy[ok], xx[ok]), val[ok])
    cnt = np.bincount(tb[ok], minlength=B).astype(np.uint64)
    return R, cnt


def make_events(n, H, W, seed, span=50_000):
    rng = np.random.default_rng(seed)
    t = (np.sort(rng.integers(0, span, size=n)) + 10_000_000).astype(np.uint32)
    x = rng.integers(0, W, size=n).astype(np.uint16)
    y = rng.integers(0, H, size=n).astype(np.uint16)
    hot = rng.random(n) < 0.2                                      # a hot spot: many events per cell, carries in the low word
    x[hot] = W // 3
    y[hot] = rng.integers(0, min(H, 3), size=int(hot.sum())).astype(np.uint16)
    p = rng.integers(0, 2, size=n).astype(np.uint8)
    return t, x, y, p


@pytest.mark.parametrize("H,W,bins", [(37, 53, 5), (37, 53, 1), (24, 1500, 3), (480, 640, 2), (301, 7, 4)])
def test_banded_kernels_match_red_kernel_and_definition(emu, H, W, bins):
    n = 30_000
    t, x, y, p = make_events(n, H, W, seed=H * 1000 + W + bins)
    t[-700:] = t[-700]                                             # identical timestamps at the end
    x[100:104] = W + 3                                             # outside the sensor: dropped
    y[200:203] = H
    t[4000:4100] = t[4000:4100][::-1].copy()                       # locally unsorted
    starts = [0, 1001, 500, 777, n - 600, 3]
    ends = [n - 900, 9193 + 1001, 500, 778, n, 8192 + 8192 + 11]   # large, one-and-a-bit chunks, empty, one event, one timestamp, two chunks + tail
    red, red_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, 0)
    for s in range(len(starts)):
        want, cnt = direct_R(t, x, y, p, starts[s], ends[s], H, W, bins)
        got = red[s].astype(np.int64).reshape(bins, H, W)
        assert np.array_equal(got, want), f"window {s}: the RED kernel differs from the definition of R"
        assert np.array_equal(red_bins[s], cnt)
    for variant in (1, 2):
        got, got_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, variant)
        assert np.array_equal(got, red), f"BANDED cut {variant}: R differs from the RED kernel's"
        assert np.array_equal(got_bins, red_bins), f"BANDED cut {variant}: per-bin counts differ"


def test_banded_kernels_scalar_load_path(emu):
    """Arrays that start off the 16-byte grid take the scalar loads everywhere."""
    H, W, bins, n = 37, 53, 3, 20_000
    t, x, y, p = make_events(n + 1, H, W, seed=5)
    t, x, y, p = t[1:], x[1:], y[1:], p[1:]                         # views: base + one element
    assert x.ctypes.data % 16 != 0
    starts, ends = [0, 37], [n, 9000]
    red, red_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, 0)
    for variant in (1, 2):
        got, got_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, variant)
        assert np.array_equal(got, red) and np.array_equal(got_bins, red_bins)


@pytest.mark.parametrize("bins", [1, 3])
def test_window_longer_than_one_round_of_run_bounds(emu, bins):
    """A 4.4 M-event window is 538 chunks: the accumulate pass loads the run bounds of 512 chunks per round (32 lanes x
    16 warps), so its outer loop takes a second turn."""
    H, W, n = 24, 40, 4_400_000
    t, x, y, p = make_events(n, H, W, seed=13, span=2_000_000)
    starts, ends = [3], [n - 2]
    red, red_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, 0)
    for variant in (1, 2):
        got, got_bins = stage_a(emu, t, x, y, p, starts, ends, H, W, bins, variant)
        assert np.array_equal(got, red) and np.array_equal(got_bins, red_bins), variant
