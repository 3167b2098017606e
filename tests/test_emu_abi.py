"""The C ABI of include/cmda_b200.h on the CPU: every translation unit of libcmda_b200 is compiled for the host against the fiber emulation of tests/emu/ ("device" pointers are numpy buffers) and
run against the committed golden fixtures -- outputs of the reference's own functions -- and the oracle, with the GPU
suite's rules: bit-exact integers, indices and pseudo-events; stated tolerances for float voxel sums.  Both cuts of
the BANDED stage A (modes BANDED and BANDED2) are covered.  This checks the LOGIC of the kernel source and of the
host code around it where no GPU exists; it is test infrastructure, shares no path with the product (which has no
CPU fallback), and the `-m gpu` tests remain the gate for the CUDA build."""
import ctypes
import os
import platform
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="tests/emu switches fibers with x86-64 assembly")

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402
import golden_io  # noqa: E402
from oracle import cmda_oracle as O  # noqa: E402

VOXEL = golden_io.load("voxel")
NORM = golden_io.load("norm")
VG = golden_io.load("events_vg")
ISR = golden_io.load("isr")
IC = golden_io.load("image_change")
INDEX = golden_io.load("index")
GLOBAL, AUTO, FACTORED, BANDED, BANDED2 = 0, 2, 4, 5, 6
DIRECTIONS = {"rightdown": 0, "rightup": 1, "leftdown": 2, "leftup": 3, "all": 4}


def _bind(path):
    from cmda_b200 import _lib
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)          # the emulated build exports the whole ABI
        fn.restype, fn.argtypes = restype, argtypes
    return lib


@pytest.fixture(scope="module")
def L():
    return _bind(build_emu.build_abi())


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def workspace(nbytes):
    raw = np.full(nbytes + 512, 0xA5, dtype=np.uint8)              # garbage: nothing may rely on a zeroed workspace
    off = (-raw.ctypes.data) % 256
    return raw[off:off + nbytes]


def events_vg(L, c, mode, clip):
    t, x, y, p = (np.ascontiguousarray(c[k]) for k in ("t", "x", "y", "p"))
    W, H, B = int(c["width"]), int(c["height"]), int(c["bins"])
    rmap = np.ascontiguousarray(c["rectify_map"], dtype=np.float32)
    starts = np.array([int(c["start"])], dtype=np.int64)
    ends = np.array([int(c["finish"]) + 1], dtype=np.int64)
    clips = np.array([clip], dtype=np.float32)
    out = np.full((1, B, H, W), np.nan, dtype=np.float32)
    raw = np.full((1, B, H, W), np.nan, dtype=np.float32)
    counts = np.full((1, B), -1, dtype=np.int64)
    need = L.cmda_events_vg_workspace_bytes(int(ends[0] - starts[0]), 1, H, W, B, mode)
    ws = workspace(need)
    rc = L.cmda_events_vg_batch(ptr(t), ptr(x), ptr(y), ptr(p), ptr(starts), ptr(ends), 1, ptr(rmap), None, H, W, B, ptr(clips),
                                1.0, 1, 1, ptr(out), ptr(raw), ptr(counts), ptr(ws), need, mode, None)
    assert rc == 0, L.cmda_strerror(rc)
    return out[0], raw[0], counts[0], rmap


def test_emulated_build_exports_the_whole_abi(L):
    assert L.cmda_version() == 100
    assert L.cmda_events_vg_resolved_mode(1000, 1, 480, 640, 5, AUTO) == FACTORED


@pytest.mark.parametrize("name", sorted(k for k in VG if VG[k]["rectify_map"].ndim == 3))
@pytest.mark.parametrize("mode", [GLOBAL, FACTORED, BANDED, BANDED2])
def test_events_vg_golden(L, name, mode):
    c = VG[name]
    W, H, B = int(c["width"]), int(c["height"]), int(c["bins"])
    start, finish = int(c["start"]), int(c["finish"])
    clip = float(c["clip"][0]) if c["clip"].size else O.default_clip_range(finish, start)
    sl = slice(start, finish + 1)
    if True:
        out, raw, counts, rmap = events_vg(L, c, mode, clip)
        tf, xf, yf, pf = O.rectify_events(c["t"][sl], c["x"][sl], c["y"][sl], c["p"][sl], rmap)
        ref_raw, aux = O.events_to_voxel_grid(tf, xf, yf, pf, W, H, B, return_aux=True)
        tol = 1e-5 * np.maximum(np.abs(ref_raw), aux["abs_weight_sum"]) + aux["n_contrib"] * 2.0 ** -31
        assert np.all(np.abs(raw.astype(np.float64) - ref_raw.astype(np.float64)) <= tol)
        assert np.all(raw[aux["n_contrib"] == 0] == 0.0)
        assert np.array_equal(counts, aux["bin_counts"])
        # events_norm: within 1e-5 of the oracle on our raw grid; of the reference's output unless a voxel whose
        # exact sum is zero kept a float32 residue in the reference (tests/test_gpu_parity.py::check_normalised)
        np.testing.assert_allclose(out, O.events_norm(raw.copy(), clip, 1.0, True), rtol=0, atol=1e-5)
        if not ((raw == 0) != (ref_raw == 0)).any():
            np.testing.assert_allclose(out, c["result"], rtol=0, atol=1e-5)
    # integer side outputs: bit-exact
    n = finish + 1 - start
    xr, yr, tn = (np.empty(n, np.float32) for _ in range(3))
    x0, y0, t0 = (np.empty(n, np.int32) for _ in range(3))
    t, x, y, p = (np.ascontiguousarray(c[k]) for k in ("t", "x", "y", "p"))
    assert L.cmda_remap_events(ptr(t), ptr(x), ptr(y), ptr(p), start, finish + 1, ptr(rmap), H, W, B, ptr(xr), ptr(yr), ptr(tn),
                               ptr(x0), ptr(y0), ptr(t0), None) == 0
    assert np.array_equal(bits(xr), bits(xf)) and np.array_equal(bits(yr), bits(yf))
    assert np.array_equal(x0, np.trunc(xf).astype(np.int32)) and np.array_equal(y0, np.trunc(yf).astype(np.int32))


@pytest.mark.parametrize("name", sorted(VOXEL))
def test_voxel_grid_f32_golden(L, name):
    c = VOXEL[name]
    W, H, B, n = int(c["width"]), int(c["height"]), int(c["bins"]), int(c["time"].shape[0])
    tm, x, y, pol = (np.ascontiguousarray(c[k], dtype=np.float32) for k in ("time", "x", "y", "pol"))
    grid = np.full((B, H, W), np.nan, dtype=np.float32)
    need = L.cmda_events_vg_workspace_bytes(n, 1, H, W, B, GLOBAL)
    ws = workspace(need)
    assert L.cmda_voxel_grid_f32(ptr(tm), ptr(x), ptr(y), ptr(pol), n, W, H, B, ptr(grid), None, ptr(ws), need, GLOBAL, None) == 0
    ref, aux = O.events_to_voxel_grid(tm, x, y, pol, W, H, B, return_aux=True)
    assert np.array_equal(bits(ref), bits(c["grid"])), "the oracle reproduces the reference's grid bit for bit"
    tol = 1e-5 * np.maximum(np.abs(ref), aux["abs_weight_sum"]) + aux["n_contrib"] * 2.0 ** -31
    assert np.all(np.abs(grid.astype(np.float64) - ref.astype(np.float64)) <= tol)


@pytest.mark.parametrize("name", sorted(k for k in NORM if k.startswith("norm_")))
def test_events_norm_golden(L, name):
    c = NORM[name]
    grid = np.ascontiguousarray(NORM["normgrid_" + str(c["grid"])]["events"], dtype=np.float32).copy()
    clips = np.array([float(c["clip_range"])], dtype=np.float32)
    need = L.cmda_events_norm_workspace_bytes(1)
    ws = workspace(need)
    assert L.cmda_events_norm_batch(ptr(grid), 1, grid.size, ptr(clips), float(c["final_range"]), int(c["enforce"]), ptr(ws), need,
                                    None) == 0
    np.testing.assert_allclose(grid, c["result"], rtol=0, atol=1e-5)


def test_images_to_events_index_golden(L):
    c = INDEX["index_table"]
    t = np.ascontiguousarray(c["t"], dtype=np.uint32)
    ms = np.ascontiguousarray(c["ms_to_idx"], dtype=np.int64)
    ts = np.ascontiguousarray(c["timestamps"], dtype=np.int64)
    idx = np.full(ts.shape[0], -7, dtype=np.int64)
    status = np.full(ts.shape[0], -7, dtype=np.int32)
    assert L.cmda_images_to_events_index(ptr(t), t.shape[0], ptr(ms), ms.shape[0], int(c["t_offset"]), ptr(ts), ts.shape[0], ptr(idx),
                                         ptr(status), None) == 0
    assert np.array_equal(idx, np.asarray(c["result"], dtype=np.int64)) and not status.any()
    q = np.array([0, int(t[0]), int(t[777]), int(t[-1]), 2 ** 33], dtype=np.int64)
    got = np.empty(q.shape[0], dtype=np.int64)
    assert L.cmda_searchsorted_right_u32(ptr(t), t.shape[0], ptr(q), q.shape[0], ptr(got), None) == 0
    assert np.array_equal(got, np.searchsorted(t.astype(np.int64), q, side="right"))


def _isr(L, img, channels, c):
    from cmda_b200 import image_change as ic
    vr = tuple(float(v) for v in c["val_range"])
    lut = ic.log_lut_val_range(vr)
    span = np.log(vr[1]) - np.log(vr[0])                               # utils.py:93-94 (float64), compared in float32
    thr, clip = np.float32(span * float(c["threshold"])), np.float32(span * float(c["clip_range"]))
    H, W = img.shape[0], img.shape[1]
    out = np.full((1, H, W), np.nan, dtype=np.float32)
    need = L.cmda_image_workspace_bytes(1, H, W, channels)
    ws = workspace(need)
    src = np.ascontiguousarray(img)
    rc = L.cmda_isr_shift_u8(ptr(src), channels, 1, H, W, int(c["shift_pixel"]), DIRECTIONS[str(c["direction"])], ptr(lut), float(thr),
                             float(clip), ptr(out), ptr(ws), need, None)
    assert rc == 0
    return out, lut


@pytest.mark.parametrize("name", sorted(k for k in ISR if "lut" in ISR[k]))
def test_isr_golden_bit_exact(L, name):
    c = ISR[name]
    out, lut = _isr(L, ISR["isr_input"]["rgb"], 3, c)
    assert np.array_equal(bits(lut), bits(c["lut"])), "host LUT == the reference's np.log values"
    assert np.array_equal(bits(out), bits(c["result"]))
    out, _ = _isr(L, ISR["isr_input"]["gray"], 1, c)
    assert np.array_equal(bits(out), bits(c["result"]))


def test_rgb_to_gray_bit_exact(L):
    rgb = np.ascontiguousarray(ISR["isr_input"]["rgb"])
    gray = np.zeros(rgb.shape[:2], dtype=np.uint8)
    assert L.cmda_rgb_to_gray_u8(ptr(rgb), gray.size, ptr(gray), None) == 0
    assert np.array_equal(gray, ISR["isr_input"]["gray"])


@pytest.mark.parametrize("name", sorted(IC))
def test_image_change_pair_golden(L, name):
    from cmda_b200 import image_change as ic
    c = IC[name]
    now, front = np.ascontiguousarray(c["now"]), np.ascontiguousarray(c["front"])
    H, W = now.shape
    lut = ic.log_lut_log_add(ic.log_add)
    assert np.array_equal(bits(lut), bits(c["lut"]))
    f32 = np.full((1, H, W), np.nan, dtype=np.float32)
    u8 = np.zeros((1, H, W), dtype=np.uint8)
    need = L.cmda_image_workspace_bytes(1, H, W, 1)
    ws = workspace(need)
    assert L.cmda_logdiff_pair_u8(ptr(now), ptr(front), 1, H, W, ptr(lut), float(np.float32(ic.threshold)),
                                  float(np.float32(ic.clip_range)), ptr(f32), ptr(u8), ptr(ws), need, None) == 0
    assert np.array_equal(u8[0], c["result"])
    assert np.array_equal(bits(f32[0]), bits(O.get_image_change(now, front, return_float=True)))


@pytest.mark.parametrize("shape,size,channels", [((67, 131), (50, 40), 1), ((64, 96), (48, 32), 3), ((33, 47), (47, 60), 1)])
def test_resize_bilinear_matches_pillow(L, shape, size, channels):
    from PIL import Image
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, size=shape + ((3,) if channels == 3 else ()), dtype=np.uint8)
    out_w, out_h = size
    dst = np.zeros((out_h, out_w) + ((3,) if channels == 3 else ()), dtype=np.uint8)
    need = L.cmda_resize_bilinear_workspace_bytes(1, shape[0], shape[1], channels, out_h, out_w)
    ws = workspace(need)
    assert L.cmda_resize_bilinear_u8(ptr(img), channels, 1, shape[0], shape[1], out_h, out_w, ptr(dst), ptr(ws), need, None) == 0
    want = np.asarray(Image.fromarray(img, mode="RGB" if channels == 3 else "L").resize((out_w, out_h), Image.BILINEAR))
    assert np.array_equal(dst, want)


def _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, mode, aug=None, normalize=1):
    """cmda_events_vg_batch (or, with aug = (xy, crop_size, out_size, flips, repeat), cmda_events_vg_augmented_batch)
    with the reference's default clip; returns (out, bin_counts)."""
    S = len(starts)
    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(fins, dtype=np.int64) + 1
    clips = np.array([O.default_clip_range(int(f), int(s)) for s, f in zip(starts, fins)], dtype=np.float32)
    ids = None if mids is None else np.ascontiguousarray(mids, dtype=np.int32)
    m = np.ascontiguousarray(maps, dtype=np.float32)
    total = int(np.clip(en - st, 0, None).sum())
    counts = np.full((S, B), -1, dtype=np.int64)
    if aug is None:
        out = np.full((S, B, H, W), np.nan, dtype=np.float32)
        need = L.cmda_events_vg_workspace_bytes(total, S, H, W, B, mode)
        ws = workspace(need)
        rc = L.cmda_events_vg_batch(ptr(t), ptr(x), ptr(y), ptr(p), ptr(st), ptr(en), S, ptr(m), ptr(ids), H, W, B, ptr(clips), 1.0, 1,
                                    normalize, ptr(out), None, ptr(counts), ptr(ws), need, mode, None)
    else:
        xy, crop_size, out_size, flips, repeat = aug
        table = np.array([[cx, cy, fl] for (cx, cy), fl in zip(xy, flips)], dtype=np.int32)
        out = np.full((S, repeat * B, out_size[1], out_size[0]), np.nan, dtype=np.float32)
        need = L.cmda_events_vg_augmented_workspace_bytes(total, S, H, W, B, mode)
        ws = workspace(need)
        rc = L.cmda_events_vg_augmented_batch(ptr(t), ptr(x), ptr(y), ptr(p), ptr(st), ptr(en), S, ptr(m), ptr(ids), H, W, B, ptr(clips),
                                              1.0, 1, ptr(table), crop_size[0], crop_size[1], out_size[0], out_size[1], 0, repeat,
                                              ptr(out), None, ptr(counts), ptr(ws), need, mode, None, None)
    assert rc == 0, L.cmda_strerror(rc)
    return out, counts


def test_many_windows_and_maps_through_the_group_loop(L):
    """More windows than one launch group holds (64) and more distinct maps than one group builds plans for (8):
    the offsets of api.cu's group loop, with the FACTORED stage A and both BANDED cuts, window by window against the
    oracle."""
    from cmda_b200 import synth
    H, W, B, S, n_maps = 40, 56, 3, 150, 11
    rng = np.random.default_rng(99)
    n = 30_000
    t, x, y, p = synth.make_events(n, H, W, seed=31)
    maps = np.stack([synth.make_rectify_map(H, W, seed=200 + k) for k in range(n_maps)])
    starts = np.sort(rng.integers(0, n - 400, size=S))
    fins = starts + rng.integers(0, 400, size=S)
    fins[7] = starts[7] - 1                                  # an empty window in the first group
    fins[100] = starts[100]                                  # a single-event window in the second
    mids = rng.integers(0, n_maps, size=S)
    mids[:12] = np.arange(12) % n_maps                       # > 8 distinct maps inside the first windows
    base, base_counts = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, FACTORED)
    raw, _ = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, FACTORED, normalize=0)
    for s in range(S):
        if fins[s] < starts[s]:
            assert not raw[s].any() and int(base_counts[s].sum()) == 0
            continue
        sl = slice(int(starts[s]), int(fins[s]) + 1)
        tf, xf, yf, pf = O.rectify_events(t[sl], x[sl], y[sl], p[sl], maps[mids[s]])
        g, aux = O.events_to_voxel_grid(tf, xf, yf, pf, W, H, B, return_aux=True)
        tol = 1e-5 * np.maximum(np.abs(g), aux["abs_weight_sum"]) + aux["n_contrib"] * 2.0 ** -31
        assert np.all(np.abs(raw[s].astype(np.float64) - g.astype(np.float64)) <= tol), f"window {s}"
        assert np.array_equal(base_counts[s], aux["bin_counts"]), f"window {s}"
        clip = O.default_clip_range(int(fins[s]), int(starts[s]))
        np.testing.assert_allclose(base[s], O.events_norm(raw[s].copy(), clip, 1.0, True), rtol=0, atol=1e-5, err_msg=f"window {s}")
    for mode in (BANDED, BANDED2):
        out, counts = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, mode)
        assert np.array_equal(bits(out), bits(base)) and np.array_equal(counts, base_counts), mode


def test_fused_augmentation_many_windows(L):
    """dsec.py:304-319 fused into the normaliser (crop / flip / bilinear resize / repeat) past one launch group, on
    top of the FACTORED and both BANDED stage A: against the unfused path + the oracle's post-voxel stage."""
    from cmda_b200 import synth
    H, W, B, S = 40, 56, 2, 70
    rng = np.random.default_rng(123)
    n = 20_000
    t, x, y, p = synth.make_events(n, H, W, seed=41)
    maps = np.stack([synth.make_rectify_map(H, W, seed=300 + k) for k in range(3)])
    starts = np.sort(rng.integers(0, n - 500, size=S))
    fins = starts + rng.integers(50, 500, size=S)
    mids = rng.integers(0, 3, size=S)
    crop_size, out_size = (24, 20), (32, 28)                    # (w, h)
    xy = [(int(rng.integers(0, W - 24 + 1)), int(rng.integers(0, H - 20 + 1))) for _ in range(S)]
    flips = [int(v) for v in rng.integers(0, 2, size=S)]
    grid, _ = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, FACTORED)
    first = None
    for mode in (FACTORED, BANDED, BANDED2):
        got, _ = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, mode, aug=(xy, crop_size, out_size, flips, 3))
        assert got.shape == (S, 3 * B, 28, 32)
        if first is None:
            first = got
            for s in range(S):
                exp = O.events_vg_post(grid[s], crop_xy=xy[s], crop_size=crop_size, out_size=out_size, flip_flag=bool(flips[s]),
                                       avg_bins=False, enforce_3_channels=True, test_mode=False)
                np.testing.assert_allclose(got[s], exp, rtol=0, atol=1e-5, err_msg=f"window {s}")
        else:
            assert np.array_equal(bits(got), bits(first)), mode


def test_denorm_to_gray_and_crop_to_centered(L):
    """f-2 / f-4 hand-offs: normalised float image -> uint8 RGB -> PIL 'L' (dacs.py:730-733), and crop -> flip ->
    (x / 255 - 0.5) / 0.5 -> repeat (cityscapes_ic.py:177-183, 207-209): bit-exact against the oracle."""
    rng = np.random.default_rng(21)
    means, stds = np.array([123.675, 116.28, 103.53], np.float32), np.array([58.395, 57.12, 57.375], np.float32)
    S, H, W = 2, 48, 64
    img = rng.normal(0.0, 1.4, size=(S, 3, H, W)).astype(np.float32)
    gray = np.zeros((S, H, W), dtype=np.uint8)
    rgb = np.zeros((S, H, W, 3), dtype=np.uint8)
    assert L.cmda_denorm_rgb_to_gray_u8(ptr(img), S, H, W, ptr(means), ptr(stds), ptr(gray), ptr(rgb), None) == 0
    for s in range(S):
        g_ref, rgb_ref = O.mixed_image_to_gray(img[s], means, stds, return_rgb=True)
        assert np.array_equal(gray[s], g_ref) and np.array_equal(rgb[s], rgb_ref)
    table = np.array([[5, 3, 0], [17, 9, 1]], dtype=np.int32)              # crop_x, crop_y, flip
    cw, ch, rep = 40, 32, 3
    out = np.full((S, rep, ch, cw), np.nan, dtype=np.float32)
    assert L.cmda_u8_crop_to_centered_f32(ptr(gray), S, H, W, ptr(table), cw, ch, rep, ptr(out), None) == 0
    for s in range(S):
        want = O.u8_crop_to_centered(gray[s], (int(table[s, 0]), int(table[s, 1])), (cw, ch), flip_flag=bool(table[s, 2]), repeat=rep)
        assert np.array_equal(bits(out[s]), bits(want))


@pytest.mark.parametrize("avg,test_mode", [(True, False), (False, True)])
def test_fused_augmentation_options(L, avg, test_mode):
    """The other options of dsec.py:304-319 in the fused normaliser: mean over the bins (events_bins_5_avg_1) and the
    test-mode crop [:, :440, :] without resize, against the oracle's post-voxel stage applied to the unfused grid."""
    from cmda_b200 import synth
    H, W, B, n = 480, 640, 5, 60_000
    t, x, y, p = synth.make_events(n, H, W, seed=synth.seed_for(5, 5))
    rmap = synth.make_rectify_map(H, W, seed=12)[None]
    starts, fins = [0, 5000], [n - 1, 40_000]
    crop_size, out_size = ((W, 440), (W, 440)) if test_mode else ((400, 400), (512, 512))
    xy = [(0, 0), (0, 0)] if test_mode else [(37, 61), (240, 80)]
    flips = [0, 0] if test_mode else [1, 0]
    grid, _ = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, B, FACTORED)
    S = 2
    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(fins, dtype=np.int64) + 1
    clips = np.array([O.default_clip_range(int(f), int(s)) for s, f in zip(starts, fins)], dtype=np.float32)
    table = np.array([[cx, cy, fl] for (cx, cy), fl in zip(xy, flips)], dtype=np.int32)
    Bo = 1 if avg else B
    for mode in (FACTORED, BANDED2):
        out = np.full((S, 3 * Bo, out_size[1], out_size[0]), np.nan, dtype=np.float32)
        need = L.cmda_events_vg_augmented_workspace_bytes(int((en - st).sum()), S, H, W, B, mode)
        ws = workspace(need)
        rc = L.cmda_events_vg_augmented_batch(ptr(t), ptr(x), ptr(y), ptr(p), ptr(st), ptr(en), S, ptr(rmap), None, H, W, B, ptr(clips),
                                              1.0, 1, ptr(table), crop_size[0], crop_size[1], out_size[0], out_size[1], int(avg), 3,
                                              ptr(out), None, None, ptr(ws), need, mode, None, None)
        assert rc == 0, L.cmda_strerror(rc)
        for s in range(S):
            exp = O.events_vg_post(grid[s], crop_xy=xy[s], crop_size=crop_size, out_size=out_size, flip_flag=bool(flips[s]),
                                   avg_bins=avg, enforce_3_channels=True, test_mode=test_mode)
            np.testing.assert_allclose(out[s], exp, rtol=0, atol=1e-5)


def test_prebuilt_plans_equal_per_call_plans(L):
    """cmda_rectify_plan_build once per sequence == the plans built inside every call: same bits."""
    from cmda_b200 import synth
    H, W, B, n = 120, 160, 3, 30_000
    t, x, y, p = synth.make_events(n, H, W, seed=8)
    maps = np.stack([synth.make_rectify_map(H, W, seed=40 + k) for k in range(3)]).astype(np.float32)
    starts, fins, mids = [0, 100, 7000], [n - 1, 9000, 20_000], np.array([2, 0, 1], dtype=np.int32)
    base, counts = _vg_batch(L, t, x, y, p, starts, fins, maps, mids, H, W, B, FACTORED)
    nb = L.cmda_rectify_plan_bytes(H, W)
    plans = workspace(3 * nb)
    assert L.cmda_rectify_plan_build(ptr(maps), 3, H, W, ptr(plans), None) == 0
    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(fins, dtype=np.int64) + 1
    clips = np.array([O.default_clip_range(int(f), int(s)) for s, f in zip(starts, fins)], dtype=np.float32)
    for mode in (FACTORED, BANDED, BANDED2):
        out = np.full((3, B, H, W), np.nan, dtype=np.float32)
        need = L.cmda_events_vg_workspace_bytes(int((en - st).sum()), 3, H, W, B, mode)
        ws = workspace(need)
        rc = L.cmda_events_vg_batch_planned(ptr(t), ptr(x), ptr(y), ptr(p), ptr(st), ptr(en), 3, ptr(maps), ptr(mids), H, W, B,
                                            ptr(clips), 1.0, 1, 1, ptr(out), None, None, ptr(ws), need, mode, ptr(plans), None)
        assert rc == 0, L.cmda_strerror(rc)
        assert np.array_equal(bits(out), bits(base)), mode


def test_randomised_geometry_all_stage_a_forms_agree(L):
    """Seeded random grids (odd sizes, 1-9 bins, widths that leave one row per band), ragged windows with duplicate
    timestamps and out-of-sensor events: FACTORED, BANDED and BANDED2 produce the same raw grids and per-bin counts,
    bit for bit.  With polarity bytes beyond {0, 1} the BANDED cuts flag the window (one sign bit per record) and the
    fallback recomputes it with the reference's own weights: same counts, the GLOBAL mode's grid bit for bit."""
    from cmda_b200 import synth
    rng = np.random.default_rng(20260117)
    for it in range(10):
        H = int(rng.integers(3, 70))
        W = int(rng.choice([int(rng.integers(3, 90)), int(rng.integers(1200, 1700)), 24_000]))
        if W > 2000:
            H = int(rng.integers(1, 4))
        B = int(rng.integers(1, 10))
        n = int(rng.integers(200, 12_000))
        t, x, y, p = synth.make_events(n, H, W, seed=int(rng.integers(1 << 30)))
        t = np.sort((t - t.min()) // int(rng.choice([1, 7, 400]))).astype(np.uint32) + 5        # duplicate timestamps
        x[rng.random(n) < 0.01] = W + int(rng.integers(0, 3))
        y[rng.random(n) < 0.01] = H
        rmap = synth.make_rectify_map(H, W, seed=it)[None]
        S = int(rng.integers(1, 6))
        starts = np.sort(rng.integers(0, n, size=S))
        fins = np.minimum(starts + rng.integers(-1, n, size=S), n - 1)
        for odd in (False, True):
            if odd:
                p = p.copy()
                p[rng.random(n) < 0.02] = int(rng.integers(2, 256))
            base, counts = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, B, FACTORED, normalize=0)
            for mode in (BANDED, BANDED2):
                got, c2 = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, B, mode, normalize=0)
                assert np.array_equal(c2, counts), (it, H, W, B, mode, odd)
                if not odd:
                    assert np.array_equal(bits(got), bits(base)), (it, H, W, B, mode)
                    continue
                # a window that holds such a byte (and is alive: two distinct timestamps) is recomputed with the GLOBAL
                # formulation, bit for bit; the others keep the sensor-space sums
                glob, _ = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, B, GLOBAL, normalize=0)
                for s in range(S):
                    if fins[s] < starts[s]:
                        continue
                    sl = slice(int(starts[s]), int(fins[s]) + 1)
                    flagged = bool((p[sl] > 1).any()) and t[sl][0] != t[sl][-1]
                    assert np.array_equal(bits(got[s]), bits(glob[s] if flagged else base[s])), (it, H, W, B, mode, s, flagged)
            for s in range(S):
                if fins[s] < starts[s]:
                    assert not base[s].any() and int(counts[s].sum()) == 0


@pytest.mark.parametrize("shape", [(67, 131), (40, 56), (9, 5), (33, 260)])
def test_pseudo_events_odd_sizes_every_direction(L, shape):
    """Sizes that leave the 4-pixel vector paths, tiny images, shifts up to the image size, constant and two-level
    images: both generators bit-exact against the oracle, every direction, batches of several images."""
    from cmda_b200 import image_change as ic
    H, W = shape
    rng = np.random.default_rng(H * 1000 + W)
    imgs = np.stack([rng.integers(0, 256, size=(H, W), dtype=np.uint8),
                     np.full((H, W), int(rng.integers(0, 256)), dtype=np.uint8),
                     (rng.integers(0, 2, size=(H, W)) * 255).astype(np.uint8)])
    S = imgs.shape[0]
    for vr, thr_f, clip_f, shift in (((1, 100), 0.04, 0.2, 3), ((0.01, 1.01), 0.005, 0.1, 1), ((1e-5, 255 + 1e-5), 0.0, 0.04, min(H, W))):
        lut = ic.log_lut_val_range(tuple(float(v) for v in vr))
        span = np.log(vr[1]) - np.log(vr[0])
        thr, clip = np.float32(span * thr_f), np.float32(span * clip_f)
        for name, code in DIRECTIONS.items():
            out = np.full((S, 1, H, W), np.nan, dtype=np.float32)
            need = L.cmda_image_workspace_bytes(S, H, W, 1)
            ws = workspace(need)
            assert L.cmda_isr_shift_u8(ptr(imgs), 1, S, H, W, shift, code, ptr(lut), float(thr), float(clip), ptr(out), ptr(ws), need,
                                       None) == 0
            for s in range(S):
                want = O.get_image_change_from_pil(imgs[s], W, H, shift_pixel=shift, val_range=vr, _threshold=thr_f,
                                                   _clip_range=clip_f, shift_direction=name)
                assert np.array_equal(bits(out[s]), bits(want)), (vr, name, s)
    front = np.roll(imgs, 1, axis=2)
    lut = ic.log_lut_log_add(ic.log_add)
    f32 = np.full((S, H, W), np.nan, dtype=np.float32)
    u8 = np.zeros((S, H, W), dtype=np.uint8)
    need = L.cmda_image_workspace_bytes(S, H, W, 1)
    ws = workspace(need)
    assert L.cmda_logdiff_pair_u8(ptr(imgs), ptr(front), S, H, W, ptr(lut), float(np.float32(ic.threshold)),
                                  float(np.float32(ic.clip_range)), ptr(f32), ptr(u8), ptr(ws), need, None) == 0
    for s in range(S):
        assert np.array_equal(u8[s], O.get_image_change(imgs[s], front[s]))
        assert np.array_equal(bits(f32[s]), bits(O.get_image_change(imgs[s], front[s], return_float=True)))


TILED, EXACT = 1, 3


@pytest.mark.parametrize("name", sorted(k for k in VG if VG[k]["rectify_map"].ndim == 3))
def test_events_vg_golden_tiled_and_exact(L, name):
    """TILED sums the same quantised weights as GLOBAL: the same bits.  EXACT performs the reference's own float32
    additions in the reference's own order: its raw grid is BIT-IDENTICAL to the oracle's (which reproduces the
    reference's grid bit for bit, test_voxel_grid_f32_golden) and its normalised grid is within 1e-5 of the fixture."""
    c = VG[name]
    W, H, B = int(c["width"]), int(c["height"]), int(c["bins"])
    start, finish = int(c["start"]), int(c["finish"])
    clip = float(c["clip"][0]) if c["clip"].size else O.default_clip_range(finish, start)
    sl = slice(start, finish + 1)
    g_out, g_raw, g_counts, rmap = events_vg(L, c, GLOBAL, clip)
    t_out, t_raw, t_counts, _ = events_vg(L, c, TILED, clip)
    assert np.array_equal(bits(t_raw), bits(g_raw)) and np.array_equal(bits(t_out), bits(g_out)) and np.array_equal(t_counts, g_counts)
    e_out, e_raw, e_counts, _ = events_vg(L, c, EXACT, clip)
    tf, xf, yf, pf = O.rectify_events(c["t"][sl], c["x"][sl], c["y"][sl], c["p"][sl], rmap)
    ref_raw, aux = O.events_to_voxel_grid(tf, xf, yf, pf, W, H, B, return_aux=True)
    assert np.array_equal(bits(e_raw), bits(ref_raw)), "EXACT must reproduce the reference's raw grid bit for bit"
    assert np.array_equal(e_counts, aux["bin_counts"])
    np.testing.assert_allclose(e_out, c["result"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("name", sorted(VOXEL))
def test_voxel_grid_f32_golden_exact_mode(L, name):
    """events_to_voxel_grid through the EXACT mode against the fixture the reference itself produced: the same bits."""
    c = VOXEL[name]
    W, H, B, n = int(c["width"]), int(c["height"]), int(c["bins"]), int(c["time"].shape[0])
    tm, x, y, pol = (np.ascontiguousarray(c[k], dtype=np.float32) for k in ("time", "x", "y", "pol"))
    grid = np.full((B, H, W), np.nan, dtype=np.float32)
    need = L.cmda_events_vg_workspace_bytes(n, 1, H, W, B, EXACT)
    ws = workspace(need)
    assert L.cmda_voxel_grid_f32(ptr(tm), ptr(x), ptr(y), ptr(pol), n, W, H, B, ptr(grid), None, ptr(ws), need, EXACT, None) == 0
    assert np.array_equal(bits(grid), bits(c["grid"]))


def test_capacity_guard_routes_overflowing_windows_to_the_fallback(L):
    """An R cell holds |C| < 2^19 in units of |2 * pol - 1|.  A hot pixel with polarity byte 255 (value 509) reaches
    that with ~1 100 events per temporal interval: without the guard the sensor-space sums wrap and the grid is wrong.
    The sketch of the RED path / the flags of the BANDED cut must route such a window to the fallback (the GLOBAL
    formulation) and leave its neighbours on the fast path: every window within the raw-grid bound of the oracle."""
    from cmda_b200 import synth
    H, W, B = 24, 40, 3
    rng = np.random.default_rng(5)
    n_hot, n_bg = 2600, 3000
    t, x, y, p = synth.make_events(n_hot + n_bg, H, W, seed=77)
    # window 0: a single hot pixel, every polarity byte 255 (1 300 events per interval x 509 > 2^19: a real overflow);
    # window 1: ordinary DSEC events
    x[:n_hot], y[:n_hot], p[:n_hot] = 7, 5, 255
    t[:n_hot] = np.sort(t[:n_hot])
    t[n_hot:] = np.sort(t[n_hot:])
    rmap = synth.make_rectify_map(H, W, seed=3)[None]
    starts, fins = np.array([0, n_hot]), np.array([n_hot - 1, n_hot + n_bg - 1])
    for mode in (FACTORED, BANDED, BANDED2, AUTO):
        raw, counts = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, B, mode, normalize=0)
        for s in range(2):
            sl = slice(int(starts[s]), int(fins[s]) + 1)
            tf, xf, yf, pf = O.rectify_events(t[sl], x[sl], y[sl], p[sl], rmap[0])
            g, aux = O.events_to_voxel_grid(tf, xf, yf, pf, W, H, B, return_aux=True)
            tol = 1e-5 * np.maximum(np.abs(g), aux["abs_weight_sum"]) + aux["n_contrib"] * 2.0 ** -31
            assert np.all(np.abs(raw[s].astype(np.float64) - g.astype(np.float64)) <= tol), (mode, s)
            assert np.array_equal(counts[s], aux["bin_counts"]), (mode, s)
    # B == 1 through BANDED: the record keeps one sign bit, so the polarity byte alone flags the window
    raw1, _ = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, 1, BANDED, normalize=0)
    base1, _ = _vg_batch(L, t, x, y, p, starts, fins, rmap, None, H, W, 1, FACTORED, normalize=0)
    np.testing.assert_allclose(raw1[0], base1[0], rtol=1e-6, atol=1e-3)
    assert np.array_equal(bits(raw1[1]), bits(base1[1]))


def _vg_batch_p4(L, rec, table, starts, fins, maps, mids, H, W, B, mode, src=None, normalize=1):
    S = len(starts)
    st = np.ascontiguousarray(starts, dtype=np.int64)
    en = np.ascontiguousarray(fins, dtype=np.int64) + 1
    base = st if src is None else np.ascontiguousarray(src, dtype=np.int64)
    clips = np.array([O.default_clip_range(int(f), int(s)) for s, f in zip(starts, fins)], dtype=np.float32)
    ids = None if mids is None else np.ascontiguousarray(mids, dtype=np.int32)
    m = np.ascontiguousarray(maps, dtype=np.float32)
    total = int(np.clip(en - st, 0, None).sum())
    counts = np.full((S, B), -1, dtype=np.int64)
    out = np.full((S, B, H, W), np.nan, dtype=np.float32)
    need = L.cmda_events_vg_workspace_bytes(total, S, H, W, B, mode)
    ws = workspace(need)
    rc = L.cmda_events_vg_batch_p4(ptr(rec), ptr(table), ptr(table), len(table) - 1, ptr(st), ptr(en), None if src is None else ptr(base),
                                   S, ptr(m), ptr(ids), H, W, B, ptr(clips), 1.0, 1, normalize, ptr(out), None, ptr(counts), ptr(ws),
                                   need, mode, None, None)
    assert rc == 0, L.cmda_strerror(rc)
    return out, counts


def test_packed_p4_source_is_bit_identical_to_soa(L):
    """The packed event stream (4 bytes per event, millisecond bucket from ms_to_idx) through cmda_events_vg_batch_p4
    against the SoA entry point on the stream it was packed from: raw grids, normalised grids and per-bin counts bit
    for bit, for the RED kernel and the BANDED cut, B = 1 / 3 / 5; ragged windows (empty, one event, unaligned, one
    timestamp), a sparse stream (a chunk spans far more than the 32 buckets a CTA keeps in shared memory) and a
    staging buffer that holds only the windows' events (h_win_src).  The device packer against the numpy packer."""
    from cmda_b200 import packed, synth
    H, W = 33, 47
    rng = np.random.default_rng(17)
    for density, n in (("dense", 30_000), ("sparse", 3_000)):
        t, x, y, p = synth.make_events(n, H, W, window_us=50_000 if density == "dense" else 40_000_000, seed=23)
        rmap = synth.make_rectify_map(H, W, seed=9)[None]
        rec, table, t_base = packed.pack_p4(t, x, y, p)
        t2, x2, y2, p2 = packed.unpack_p4(rec, table, t_base)
        assert np.array_equal(t, t2) and np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(p, p2)
        # the device packer
        drec = np.zeros(n, dtype=np.uint32)
        dtab = np.zeros(len(table), dtype=np.int64)
        status = np.full(1, 77, dtype=np.int32)
        assert L.cmda_pack_events_p4(ptr(t), ptr(x), ptr(y), ptr(p), n, t_base, len(table) - 1, ptr(drec), ptr(dtab), ptr(status), None) == 0
        assert status[0] == 0 and np.array_equal(drec, rec) and np.array_equal(dtab, table)
        bad_p = p.copy(); bad_p[5] = 2
        assert L.cmda_pack_events_p4(ptr(t), ptr(x), ptr(y), ptr(bad_p), n, t_base, len(table) - 1, ptr(drec), ptr(dtab), ptr(status), None) == 0
        assert status[0] == 1
        starts = np.array([0, 1001, 17, 5, n // 2, 123])
        fins = np.array([n - 1, n - 7, 9000 if n > 9000 else n // 3, 4, n // 2, 2500])
        dup = t.copy(); dup[starts[5]:fins[5] + 1] = dup[starts[5]]          # a single-timestamp window (Q3): all-zero grid
        for tt in (t, dup):
            rec_t, table_t, _ = packed.pack_p4(tt, x, y, p)
            for B in (1, 3, 5):
                for mode in (FACTORED, BANDED, BANDED2):
                    base, counts = _vg_batch(L, tt, x, y, p, starts, fins, rmap, None, H, W, B, mode)
                    got, c2 = _vg_batch_p4(L, rec_t, table_t, starts, fins, rmap, None, H, W, B, mode)
                    assert np.array_equal(bits(got), bits(base)) and np.array_equal(c2, counts), (density, B, mode)
        # a staging buffer holding two windows back to back, each at a 64-event boundary
        B = 5
        a0, a1, b0, b1 = 1001, 8000 if n > 9000 else 1500, 123, 2500
        pos_b = (a1 - a0 + 1 + 63) // 64 * 64 + 64
        stage = np.zeros(pos_b + (b1 - b0 + 1), dtype=np.uint32)
        stage[:a1 - a0 + 1] = rec[a0:a1 + 1]
        stage[pos_b:] = rec[b0:b1 + 1]
        base, counts = _vg_batch(L, t, x, y, p, [a0, b0], [a1, b1], rmap, None, H, W, B, FACTORED)
        got, c2 = _vg_batch_p4(L, stage, table, [0, pos_b], [a1 - a0, pos_b + b1 - b0], rmap, None, H, W, B, FACTORED, src=[a0, b0])
        assert np.array_equal(bits(got), bits(base)) and np.array_equal(c2, counts), density


def test_p3_wire_unpacks_to_the_p4_records(L):
    """The 3-byte wire form: cmda_unpack_p3_to_p4 writes exactly the records cmda_pack_events_p4 / pack_p4 make of the
    same stream -- the whole store, a range that starts and ends inside 16-microsecond buckets (staging-buffer use),
    an empty range; a dense stream (many events per bucket) and a sparse one (most buckets empty)."""
    from cmda_b200 import packed, synth
    H, W = 400, 1000
    for n, window_us in ((30_000, 2_000), (3_000, 40_000_000)):
        t, x, y, p = synth.make_events(n, H, W, window_us=window_us, seed=29)
        rec4, _, t_base = packed.pack_p4(t, x, y, p)
        rec3, sub, t_base3 = packed.pack_p3(t, x, y, p)
        assert t_base3 == t_base and rec3.shape == (3 * n,) and sub[0] == 0 and sub[-1] == n
        assert all(np.array_equal(a, b) for a, b in zip(packed.unpack_p3(rec3, sub, t_base), (t, x, y, p)))
        assert np.array_equal(packed.p3_to_p4(rec3, sub), rec4)
        for first, last in ((0, n), (1234, 2777), (7, 8), (n - 1, n)):
            j_lo = int(np.searchsorted(sub, first, side="right")) - 1
            j_hi = int(np.searchsorted(sub, last - 1, side="right")) - 1
            out = np.full(last - first + 2, 0xdeadbeef, dtype=np.uint32)          # one guard word on either side
            stage = np.ascontiguousarray(rec3[3 * first:3 * last])
            assert L.cmda_unpack_p3_to_p4(ptr(stage), ptr(sub), j_lo, j_hi, first, last, out[1:].ctypes.data, None) == 0
            assert np.array_equal(out[1:-1], rec4[first:last]) and out[0] == 0xdeadbeef and out[-1] == 0xdeadbeef, (n, first, last)
        assert L.cmda_unpack_p3_to_p4(ptr(rec3), ptr(sub), 0, 0, 5, 5, None, None) == 0          # empty range: nothing touched
        assert L.cmda_unpack_p3_to_p4(ptr(rec3), ptr(sub), 0, 0, 5, 4, None, None) != 0
    with pytest.raises(ValueError):
        packed.pack_p3(t, np.where(np.arange(n) == 3, 1024, x).astype(np.uint16), y, p)
    with pytest.raises(ValueError):
        packed.pack_p3(t, x, y, p, t_base=int(t[0]) // 1000 * 1000 - 16)


@pytest.mark.parametrize("seed", list(range(12)))
def test_p3_wire_randomised(L, seed):
    """Adversarial streams through the P3 packer and the unpack kernel: runs of equal timestamps, gaps of seconds (empty
    buckets by the thousand), a first event well after t_base, timestamps up to the last 16 microseconds before 2^32,
    1 - 3 events, coordinates at the format's limits."""
    from cmda_b200 import packed
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.choice([1, 2, 3, 17, 500, 5000]))
    kind = seed % 4
    if kind == 0:
        t = np.sort(rng.integers(0, 40, size=n)) + 123_456
    elif kind == 1:
        t = np.cumsum(rng.choice([0, 0, 1, 15, 16, 17, 999, 1000, 1001, 2_000_000], size=n)) + 5_000
    elif kind == 2:
        t = (1 << 32) - 1 - np.sort(rng.integers(0, 3_000_000, size=n))[::-1]
    else:
        t = np.sort(rng.integers(0, 50_000, size=n)) + 1_000_000 * int(rng.integers(0, 4000))
    t = t.astype(np.uint32)
    x = rng.choice([0, 1, 639, 1023], size=n).astype(np.uint16)
    y = rng.choice([0, 479, 511], size=n).astype(np.uint16)
    p = rng.integers(0, 2, size=n).astype(np.uint8)
    t_base = None if seed % 3 else int(t[0]) // 1000 * 1000 - 1000 * int(rng.integers(0, 3)) if int(t[0]) >= 3000 else None
    rec4, _, tb = packed.pack_p4(t, x, y, p, t_base=t_base)
    rec3, sub, tb3 = packed.pack_p3(t, x, y, p, t_base=t_base)
    assert tb3 == tb and all(np.array_equal(a, b) for a, b in zip(packed.unpack_p3(rec3, sub, tb), (t, x, y, p)))
    assert np.array_equal(packed.p3_to_p4(rec3, sub), rec4)
    first = int(rng.integers(0, n))
    last = int(rng.integers(first + 1, n + 1))
    for a, b in ((0, n), (first, last)):
        j_lo = int(np.searchsorted(sub, a, side="right")) - 1
        j_hi = int(np.searchsorted(sub, b - 1, side="right")) - 1
        out = np.zeros(b - a, dtype=np.uint32)
        assert L.cmda_unpack_p3_to_p4(ptr(np.ascontiguousarray(rec3[3 * a:3 * b])), ptr(sub), j_lo, j_hi, a, b, ptr(out), None) == 0
        assert np.array_equal(out, rec4[a:b]), (seed, a, b)


def test_table_driven_frame_pair_path(L):
    """Images of >= 2^17 pixels take the table-driven uint8 apply pass when only the uint8 output is asked for (one
    evaluation per (now, front) byte pair and image, then a gather from a 64 KB table in shared memory): bit-exact
    against the oracle for the three output selections (the float32 ones use the arithmetic kernels), two different
    images per batch (a persistent CTA crosses from one image's table to the next)."""
    from cmda_b200 import image_change as ic
    from cmda_b200 import synth
    H, W, S = 256, 512, 2
    rng = np.random.default_rng(41)
    now = np.stack([synth.make_frame_pair(H, W, seed=s)[0] for s in range(S)])
    front = np.stack([synth.make_frame_pair(H, W, seed=s)[1] for s in range(S)])
    now[0, :3, :] = rng.integers(0, 256, size=(3, W))              # |now - front| beyond the shared-memory band
    front[1, 7, :] = 255 - now[1, 7, :]
    lut = ic.log_lut_log_add(ic.log_add)
    need = L.cmda_image_workspace_bytes(S, H, W, 1)
    assert need >= S * 65536                                        # the tables are part of the workspace at this size
    ref_u8 = np.stack([O.get_image_change(now[s], front[s]) for s in range(S)])
    ref_f32 = np.stack([O.get_image_change(now[s], front[s], return_float=True) for s in range(S)])
    for want_f32, want_u8 in ((True, True), (False, True), (True, False)):
        f32 = np.full((S, H, W), np.nan, dtype=np.float32)
        u8 = np.full((S, H, W), 7, dtype=np.uint8)
        ws = workspace(need)
        assert L.cmda_logdiff_pair_u8(ptr(now), ptr(front), S, H, W, ptr(lut), float(np.float32(ic.threshold)),
                                      float(np.float32(ic.clip_range)), ptr(f32) if want_f32 else None,
                                      ptr(u8) if want_u8 else None, ptr(ws), need, None) == 0
        if want_u8:
            assert np.array_equal(u8, ref_u8)
        if want_f32:
            assert np.array_equal(bits(f32), bits(ref_f32))
