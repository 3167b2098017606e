"""CPU oracle for the CMDA event-representation path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference's algorithm, statement by
statement, each function citing the reference file:line it follows.  It is the
checker used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  It is NEVER imported by the product
package ``cmda_b200`` -- the product path is CUDA only and fails loudly when the
extension is missing.

Pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference's OWN functions executed in
the build container: ``tests/golden/make_golden.py`` runs them (AST-extracted from
/root/reference, 1 thread + deterministic) on seeded synthetic inputs and commits
inputs+outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this file against every fixture (raw voxel grids, integer coordinates and the whole
pseudo-event path bit-exact; normalised grids to <= 1e-5 absolute because the
reference's ``torch.sum`` reduction order is a third-party detail).

All arithmetic is float32 with one rounding per operation, evaluated left to
right, exactly as the torch/numpy expressions of the reference do.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
_INT_MIN = -(2 ** 31)


# --------------------------------------------------------------------------- helpers
def trunc_to_int(v: np.ndarray) -> np.ndarray:
    """``tensor.int()`` of a float32 tensor on x86 (reference dsec.py:41-43):
    truncation toward zero; NaN, +-inf and |v| >= 2^31 give INT_MIN
    (cvttss2si's "integer indefinite"), which the bounds mask then drops
    (SURVEY.md Q1, Q3).  Returned as int64 so that ``+ 1`` cannot wrap."""
    v = np.asarray(v, dtype=F32)
    ok = np.isfinite(v) & (np.abs(v) < F32(2147483648.0))
    out = np.full(v.shape, _INT_MIN, dtype=np.int64)
    out[ok] = np.trunc(v[ok]).astype(np.int64)
    return out


def t_norm_of(time: np.ndarray, num_bins: int) -> np.ndarray:
    """dsec.py:38-39: ``(C-1) * (t - t[0]) / (t[-1] - t[0])`` in float32, left to right."""
    time = np.asarray(time, dtype=F32)
    with np.errstate(all="ignore"):
        a = time - time[0]
        b = F32(num_bins - 1) * a
        return (b / F32(time[-1] - time[0])).astype(F32)


# --------------------------------------------------------------------------- a4
def events_to_voxel_grid(time, x, y, pol, width, height, num_bins, return_aux=False):
    """Trilinear (x, y, t) scatter-add (reference mmseg/datasets/dsec.py:26-58,
    ``normalize_flag=False`` which is the only way the reference calls it,
    dsec.py:356-357).

    Accumulation order is the deterministic reference's: corner pass major
    (x outer, y, t inner: dsec.py:47-49), event index minor (single-thread
    ``put_(accumulate=True)``, dsec.py:58); ``np.add.at`` applies the float32
    adds one at a time in exactly that order.
    """
    time = np.ascontiguousarray(time, dtype=F32)
    x = np.ascontiguousarray(x, dtype=F32)
    y = np.ascontiguousarray(y, dtype=F32)
    pol = np.ascontiguousarray(pol, dtype=F32)
    assert x.shape == y.shape == pol.shape == time.shape  # dsec.py:28
    assert x.ndim == 1                                    # dsec.py:29
    C, H, W = int(num_bins), int(height), int(width)
    grid = np.zeros(C * H * W, dtype=F32)                 # dsec.py:31
    aux = {"abs_weight_sum": np.zeros(C * H * W, dtype=np.float64),
           "n_contrib": np.zeros(C * H * W, dtype=np.int64)} if return_aux else None

    t_norm = t_norm_of(time, C)                           # dsec.py:38-39
    x0 = trunc_to_int(x)                                  # dsec.py:41
    y0 = trunc_to_int(y)                                  # dsec.py:42
    t0 = trunc_to_int(t_norm)                             # dsec.py:43
    value = F32(2) * pol - F32(1)                         # dsec.py:45

    with np.errstate(all="ignore"):
        for xlim in (x0, x0 + 1):                         # dsec.py:47
            for ylim in (y0, y0 + 1):                     # dsec.py:48
                for tlim in (t0, t0 + 1):                 # dsec.py:49
                    mask = ((xlim < W) & (xlim >= 0) & (ylim < H) & (ylim >= 0)
                            & (tlim >= 0) & (tlim < C))   # dsec.py:50
                    wx = F32(1) - np.abs(xlim.astype(F32) - x)
                    wy = F32(1) - np.abs(ylim.astype(F32) - y)
                    wt = F32(1) - np.abs(tlim.astype(F32) - t_norm)
                    w = ((value * wx) * wy) * wt          # dsec.py:51-52, each op rounded
                    index = H * W * tlim + W * ylim + xlim  # dsec.py:54-56
                    np.add.at(grid, index[mask], w[mask].astype(F32))  # dsec.py:58
                    if return_aux:
                        np.add.at(aux["abs_weight_sum"], index[mask], np.abs(w[mask]).astype(np.float64))
                        np.add.at(aux["n_contrib"], index[mask], 1)
    grid = grid.reshape(C, H, W)
    if return_aux:
        aux["abs_weight_sum"] = aux["abs_weight_sum"].reshape(C, H, W)
        aux["n_contrib"] = aux["n_contrib"].reshape(C, H, W)
        aux["x0"], aux["y0"], aux["t0"] = x0, y0, t0
        in_t = (t0 >= 0) & (t0 < C)
        aux["bin_counts"] = np.bincount(t0[in_t], minlength=C).astype(np.int64)
        return grid, aux
    return grid


def voxel_grid_f64(time, x, y, pol, width, height, num_bins, return_aux=False):
    """Independent float64 'truth' of the same scatter: the float32 per-corner
    weights of the reference (bit-exact per contribution) summed exactly in
    float64.  Used to show which of two float32 summation orders is closer."""
    time = np.ascontiguousarray(time, dtype=F32)
    x = np.ascontiguousarray(x, dtype=F32)
    y = np.ascontiguousarray(y, dtype=F32)
    pol = np.ascontiguousarray(pol, dtype=F32)
    C, H, W = int(num_bins), int(height), int(width)
    grid = np.zeros(C * H * W, dtype=np.float64)
    abs_w = np.zeros(C * H * W, dtype=np.float64)
    n_contrib = np.zeros(C * H * W, dtype=np.int64)
    t_norm = t_norm_of(time, C)
    x0, y0, t0 = trunc_to_int(x), trunc_to_int(y), trunc_to_int(t_norm)
    value = F32(2) * pol - F32(1)
    with np.errstate(all="ignore"):
        for xlim in (x0, x0 + 1):
            for ylim in (y0, y0 + 1):
                for tlim in (t0, t0 + 1):
                    mask = ((xlim < W) & (xlim >= 0) & (ylim < H) & (ylim >= 0) & (tlim >= 0) & (tlim < C))
                    w = ((value * (F32(1) - np.abs(xlim.astype(F32) - x)))
                         * (F32(1) - np.abs(ylim.astype(F32) - y))) * (F32(1) - np.abs(tlim.astype(F32) - t_norm))
                    idx = (H * W * tlim + W * ylim + xlim)[mask]
                    grid += np.bincount(idx, weights=w[mask].astype(np.float64), minlength=C * H * W)
                    if return_aux:
                        abs_w += np.bincount(idx, weights=np.abs(w[mask]).astype(np.float64), minlength=C * H * W)
                        n_contrib += np.bincount(idx, minlength=C * H * W)
    if return_aux:
        return grid.reshape(C, H, W), abs_w.reshape(C, H, W), n_contrib.reshape(C, H, W)
    return grid.reshape(C, H, W)


# --------------------------------------------------------------------------- a5
def tensor_normalize_to_range(tensor, min_val, max_val):
    """dsec.py:73-77 / utils.py:10-14 / create_cityscapes_image_change.py:9-13."""
    tensor = np.asarray(tensor, dtype=F32)
    with np.errstate(all="ignore"):
        tmin = F32(tensor.min())
        tmax = F32(tensor.max())
        den = F32(F32(tmax - tmin) + F32(1e-8))
        out = (tensor - tmin) / den
        out = out * F32(max_val - min_val)
        out = out + F32(min_val)
    return out.astype(F32)


def events_norm(events, clip_range=1.0, final_range=1.0, enforce_no_events_zero=False):
    """Global z-score over the non-zero voxels, split by sign, clamp, min-max
    (reference mmseg/datasets/dsec.py:80-121).  Sums are accumulated in float64
    and rounded to float32 once (torch's own float32 reduction order is a
    third-party detail; the effect on the output is <= 2e-7)."""
    events = np.array(events, dtype=F32, copy=True)
    with np.errstate(all="ignore"):
        if isinstance(clip_range, str):
            assert clip_range == "auto"
            n_mean = F32(F32(events[events < 0].astype(np.float64).mean()) * F32(1.5))   # dsec.py:85
            p_mean = F32(F32(events[events > 0].astype(np.float64).mean()) * F32(1.5))   # dsec.py:86
        else:
            nz = events != 0                                                    # dsec.py:88
            n = int(nz.sum())                                                   # dsec.py:89
            if n > 0:                                                           # dsec.py:90
                mean = F32(F32(events.astype(np.float64).sum()) / F32(n))       # dsec.py:91
                sq = F32((events * events).astype(np.float64).sum())
                std = F32(np.sqrt(F32(F32(sq / F32(n)) - F32(mean * mean))))    # dsec.py:92
                events = (nz.astype(F32) * (events - mean)) / F32(std + F32(1e-8))  # dsec.py:93-94
            n_mean = F32(-clip_range)                                           # dsec.py:95
            p_mean = F32(clip_range)                                            # dsec.py:96
        if enforce_no_events_zero:                                              # dsec.py:106
            neg = events.copy()                                                 # dsec.py:107
            events[events < 0] = 0                                              # dsec.py:108
            events = np.clip(events, F32(0), p_mean)                            # dsec.py:110
            events = tensor_normalize_to_range(events, 0, final_range)          # dsec.py:111
            neg[neg > 0] = 0                                                    # dsec.py:112
            neg = np.clip(neg, n_mean, F32(0))                                  # dsec.py:114
            neg = tensor_normalize_to_range(neg, -final_range, 0)               # dsec.py:116
            events = events + neg                                               # dsec.py:117
        else:
            events = np.clip(events, F32(-clip_range), F32(clip_range)) * F32(final_range)  # dsec.py:119
            events = events / F32(clip_range) * F32(final_range)                # dsec.py:120
    return events.astype(F32)


# --------------------------------------------------------------------------- a3
def default_clip_range(finish: int, start: int) -> float:
    """dsec.py:362 -- a Python float (float64)."""
    return (finish - start) / 500000 * 1.5


def rectify_events(t, x, y, p, rectify_map):
    """dsec.py:347-355: window-relative float32 time in [0, 1], float32 polarity,
    and the ``rectify_map[y, x]`` gather (channel 0 = x, 1 = y)."""
    t = np.asarray(t)
    with np.errstate(all="ignore"):
        tf = (t - t[0]).astype(F32)                  # dsec.py:347 (integer subtraction first)
        tf = (tf / tf[-1]).astype(F32)               # dsec.py:348
    pf = np.asarray(p).astype(F32)                   # dsec.py:349
    if rectify_map is not None:
        xy = np.asarray(rectify_map)[np.asarray(y), np.asarray(x)]  # dsec.py:351
        xf, yf = xy[:, 0], xy[:, 1]                  # dsec.py:352-353
    else:
        xf, yf = np.asarray(x), np.asarray(y)
    return tf, xf.astype(F32), yf.astype(F32), pf    # dsec.py:354-355


def get_events_vg(t, x, y, p, rectify_map, width, height, bins, finish, start, clip_range=None,
                  return_raw=False):
    """``DSECDataset.get_events_vg`` (dsec.py:341-366) on in-memory SoA arrays:
    inclusive slice ``[start, finish]``, rectify, voxelize, ``events_norm`` with
    ``final_range=1.0, enforce_no_events_zero=True``."""
    sl = slice(start, finish + 1)                    # dsec.py:342-345
    tf, xf, yf, pf = rectify_events(t[sl], x[sl], y[sl], p[sl], rectify_map)
    raw = events_to_voxel_grid(tf, xf, yf, pf, width, height, bins)          # dsec.py:356
    if clip_range is None:
        clip_range = default_clip_range(finish, start)                        # dsec.py:362
    out = events_norm(raw, clip_range=clip_range, final_range=1.0, enforce_no_events_zero=True)  # dsec.py:365
    return (out, raw) if return_raw else out


# --------------------------------------------------------------------------- a1
def images_to_events_index(t, t_offset, ms_to_idx, timestamps):
    """create_dsec_dataset_txt.py:19-42: index of the last event with
    ``t <= ts - t_offset``, or -1 outside the stream; ``ValueError('range error!')``
    when the ms bracket does not contain the timestamp."""
    t = np.asarray(t)
    ms_to_idx = np.asarray(ms_to_idx, dtype=np.int64)
    n_total = t.shape[0]
    out = []
    for ts in np.asarray(timestamps, dtype=np.int64):
        ts_us = int(ts) - int(t_offset)                                  # :20
        if ts_us <= 0 or ts_us > int(t[-1]):                             # :21
            out.append(-1)                                               # :22
            continue
        ms = max(math.floor(ts_us / 1000) - 1, 0)                        # :24-25
        left = int(ms_to_idx[ms])                                        # :26
        right = int(ms_to_idx[ms + 2])                                   # :33
        if right > n_total - 1:                                          # :34-35
            right = n_total - 1
        if not int(t[left]) <= ts_us <= int(t[right]):                   # :37-39
            raise ValueError("range error!")
        win = np.asarray(t[left:right + 1], dtype=np.int64)              # :40
        idx = int(np.searchsorted(win, ts_us, "right"))                  # :41
        out.append(left + idx - 1)                                       # :42
    return out


def events_vg_post(events_vg, crop_xy=None, crop_size=None, out_size=None, flip_flag=False, avg_bins=False,
                   enforce_3_channels=True, test_mode=False):
    """dsec.py:304-319 on one window's normalised ``[B, H, W]`` grid, statement by statement, with the
    reference's own torch calls (``torch.mean``, slicing, ``flip``, ``F.interpolate(bilinear,
    align_corners=False)``, ``repeat``): third-party arithmetic pinned by execution.  ``crop_size`` and
    ``out_size`` are ``(w, h)`` as ``self.crop_size`` / ``self.after_crop_resize_size`` are after the swap of
    dsec.py:150-152."""
    import torch
    import torch.nn.functional as F
    ev = torch.from_numpy(np.ascontiguousarray(events_vg, dtype=F32))[None]          # [output_num = 1, B, H, W]
    if avg_bins:
        ev = torch.mean(ev, dim=1, keepdim=True)                                      # :304-305
    ev = ev[0]                                                                        # :306-307
    if not test_mode:
        x, y = crop_xy
        ev = ev[:, y: y + crop_size[1], x: x + crop_size[0]]                           # :310
        if flip_flag:
            ev = torch.flip(ev, dims=[-1])                                            # :311-312 (transforms.RandomHorizontalFlip(p=1))
        ev = F.interpolate(ev[None], size=(out_size[1], out_size[0]), mode='bilinear', align_corners=False)[0]   # :313-315
    else:
        ev = ev[:, :440, :]                                                           # :316-317
    if enforce_3_channels:
        ev = ev.repeat(3, 1, 1)                                                       # :318-319
    return ev.numpy()


def window_bounds(index_table, now_image_index, image_change_range=1, events_num=-1, i=0):
    """dsec.py:296-302: inclusive ``(start, finish)`` of output window ``i`` or
    ``None`` when ``start > finish``."""
    finish = int(index_table[now_image_index - i])
    if events_num != -1:
        start = finish - events_num + 1
    else:
        start = int(index_table[now_image_index - image_change_range - i])
    if start > finish:
        return None
    return start, finish


# --------------------------------------------------------------------------- a6-a8
def pil_gray_L(rgb: np.ndarray) -> np.ndarray:
    """``PIL.Image.convert('L')`` of an RGB uint8 image: ITU-R 601 in 16.16 fixed
    point with rounding, ``(19595 R + 38470 G + 7471 B + 32768) >> 16`` (third-party
    arithmetic, Pillow's ``rgb2l``; pinned by the golden fixtures)."""
    rgb = np.asarray(rgb, dtype=np.uint8)
    if rgb.ndim == 2:
        return rgb
    r, g, b = (rgb[..., c].astype(np.uint32) for c in range(3))
    return ((19595 * r + 38470 * g + 7471 * b + 32768) >> 16).astype(np.uint8)


def mixed_image_to_gray(img, means, stds, return_rgb=False, cuda_division=True):
    """dacs.py:730-733 on one normalised float32 image ``[3, H, W]``:
    ``clamp(denorm(img, means, stds), 0, 1) * 255`` -> HWC -> ``np.uint8`` (truncation) ->
    ``Image.fromarray`` -> ``convert('L')`` (the first statement of get_image_change_from_pil,
    utils.py:126).  ``denorm`` is ``img.mul(std).add(mean) / 255.0``
    (mmseg/models/utils/dacs_transforms.py:52-53); float32, one rounding per operation.

    The reference evaluates denorm on CUDA tensors (dacs.py:729), where torch divides a tensor by a Python scalar
    as ``a * fl(1 / 255)`` (ATen div kernel, CPU-scalar fast path, torch 1.7 through 2.x), not as a true division;
    the two differ by one ulp on some inputs, which the truncation to uint8 turns into a gray level.
    ``cuda_division=True`` (default) restates the arithmetic the reference actually runs; ``False`` is torch's
    CPU arithmetic (a true division), used to pin this function against torch + PIL where there is no GPU."""
    img = np.asarray(img, dtype=F32)
    m = np.asarray(means, dtype=F32).reshape(3, 1, 1)
    sd = np.asarray(stds, dtype=F32).reshape(3, 1, 1)
    v = ((img * sd).astype(F32) + m).astype(F32)
    v = (v * (F32(1.0) / F32(255.0))).astype(F32) if cuda_division else v / F32(255.0)
    v = np.clip(v.astype(F32), F32(0.0), F32(1.0)) * F32(255.0)
    rgb = np.uint8(np.transpose(v.astype(F32), (1, 2, 0)))
    gray = pil_gray_L(rgb)
    return (gray, rgb) if return_rgb else gray


def _pil_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` (src/libImaging/Resample.c, Pillow 7+; the
    library is absent from /root/reference, its algorithm is restated here and pinned against the installed
    Pillow by tests/test_oracle_golden.py): bounds ``(xmin, count)`` and 22-bit fixed-point taps per output index."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(1.0 - abs((x + xmin - center + 0.5) * ss), 0.0) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear(img: np.ndarray, size) -> np.ndarray:
    """``Image.resize(size, Image.BILINEAR)`` of a uint8 ``[H, W]`` or ``[H, W, 3]`` image (cityscapes_ic.py:153, 176):
    horizontal pass, then vertical pass on the 8-bit intermediate, int32 accumulation from 2^21, clip to 0..255."""
    img = np.asarray(img, dtype=np.uint8)
    ow, oh = int(size[0]), int(size[1])
    a = img.astype(np.int64)
    if ow != a.shape[1]:
        b, kk = _pil_bilinear_coeffs(a.shape[1], ow)
        out = np.empty((a.shape[0], ow) + a.shape[2:], np.int64)
        for xx in range(ow):
            acc = np.full((a.shape[0],) + a.shape[2:], 1 << 21, np.int64)
            for x in range(int(b[xx, 1])):
                acc += a[:, b[xx, 0] + x] * kk[xx, x]
            out[:, xx] = np.clip(acc >> 22, 0, 255)
        a = out
    if oh != a.shape[0]:
        b, kk = _pil_bilinear_coeffs(a.shape[0], oh)
        out = np.empty((oh,) + a.shape[1:], np.int64)
        for yy in range(oh):
            acc = np.full(a.shape[1:], 1 << 21, np.int64)
            for y in range(int(b[yy, 1])):
                acc += a[b[yy, 0] + y] * kk[yy, y]
            out[yy] = np.clip(acc >> 22, 0, 255)
        a = out
    return a.astype(np.uint8)


def u8_crop_to_centered(gray, crop_xy, crop_size, flip_flag=False, repeat=3):
    """cityscapes_ic.py:177-183, 207-209: crop -> HorizontalFlip -> float32 -> ``(x / 255.0 - 0.5) / 0.5`` ->
    ``repeat(3, 1, 1)`` with the reference's torch float32 arithmetic."""
    import torch
    x, y = crop_xy
    g = np.asarray(gray, dtype=np.uint8)[y: y + crop_size[1], x: x + crop_size[0]]
    if flip_flag:
        g = g[:, ::-1]
    t = (torch.from_numpy(np.ascontiguousarray(g, dtype=np.float32))[None] / 255.0 - 0.5) / 0.5
    return t.repeat(repeat, 1, 1).numpy()


def log_lut_val_range(val_range) -> np.ndarray:
    """The 256 possible values of ``np.log(img/255*(v1-v0)+v0)`` (utils.py:88-91):
    float32 numpy arithmetic with weak Python scalars, evaluated by numpy itself."""
    g = np.arange(256, dtype=F32)
    with np.errstate(all="ignore"):
        return np.log(g / 255 * (val_range[1] - val_range[0]) + val_range[0]).astype(F32)


def log_lut_log_add(log_add) -> np.ndarray:
    """``np.log(img + log_add)`` (create_cityscapes_image_change.py:17-20)."""
    g = np.arange(256, dtype=F32)
    with np.errstate(all="ignore"):
        return np.log(g + log_add).astype(F32)


def _dead_zone_split_norm(d: np.ndarray, thr, clip) -> np.ndarray:
    """utils.py:95-104 == create_cityscapes_image_change.py:22-31 on a float32
    difference image: dead zone, sign split, clamp, min-max, sum."""
    d = np.array(d, dtype=F32, copy=True)
    thr, clip = F32(thr), F32(clip)                      # compared / clamped in float32
    d[np.abs(d) <= thr] = 0
    neg = d.copy()
    d[d < 0] = 0
    d = np.clip(d, F32(0), clip)
    d = tensor_normalize_to_range(d, 0, 1)
    neg[neg > 0] = 0
    neg = np.clip(neg, -clip, F32(0))
    neg = tensor_normalize_to_range(neg, -1, 0)
    return (d + neg).astype(F32)


def get_ic(image_front, image_now, val_range, threshold, clip_range):
    """utils.py:87-105 -> float32 ``[1, H, W]``."""
    front = np.asarray(image_front, dtype=F32)
    now = np.asarray(image_now, dtype=F32)
    with np.errstate(all="ignore"):
        front = np.log(front / 255 * (val_range[1] - val_range[0]) + val_range[0])   # :88-89
        now = np.log(now / 255 * (val_range[1] - val_range[0]) + val_range[0])       # :90-91
        d = (now - front).astype(F32)                                                # :92
        span = np.log(val_range[1]) - np.log(val_range[0])                           # float64
    return _dead_zone_split_norm(d, span * threshold, span * clip_range)[None]       # :93-104


def shifted(gray: np.ndarray, shift_pixel: int, direction: str) -> np.ndarray:
    """utils.py:129-132 / 140-148: the copy shifted by ``shift_pixel`` with the
    first (right/down) or last (left/up) columns/rows left unshifted."""
    h, w = gray.shape
    s = shift_pixel
    if direction == "left":
        return np.concatenate((gray[:, s:], gray[:, w - s:]), axis=1)
    if direction == "right":
        return np.concatenate((gray[:, :s], gray[:, :w - s]), axis=1)
    if direction == "up":
        return np.concatenate((gray[s:, :], gray[h - s:, :]), axis=0)
    if direction == "down":
        return np.concatenate((gray[:s, :], gray[:h - s, :]), axis=0)
    raise AssertionError(direction)


def get_image_change_from_pil(image, width, height, data_type=None, shift_pixel=4, val_range=None,
                              _threshold=None, _clip_range=None, auto_threshold=None,
                              shift_direction="rightdown"):
    """utils.py:108-152.  ``image`` is a PIL image or a uint8 array (``[H,W,3]`` RGB
    or ``[H,W]`` gray)."""
    if auto_threshold is not None:
        raise ValueError("auto_threshold function not implement！")       # :124-125
    if hasattr(image, "convert"):
        gray = np.array(image.convert("L"))                               # :126
    else:
        gray = pil_gray_L(np.asarray(image))
    kw = dict(val_range=val_range, threshold=_threshold, clip_range=_clip_range)
    if shift_direction == "all":                                          # :128-137
        terms = [get_ic(gray, shifted(gray, shift_pixel, d), **kw) for d in ("up", "left", "down", "right")]
        return (terms[0] / F32(4) + terms[1] / F32(4) + terms[2] / F32(4) + terms[3] / F32(4)).astype(F32)
    if "left" in shift_direction:                                         # :139-143
        row = shifted(gray, shift_pixel, "left")
    else:
        assert "right" in shift_direction
        row = shifted(gray, shift_pixel, "right")
    if "up" in shift_direction:                                           # :144-148
        col = shifted(gray, shift_pixel, "up")
    else:
        assert "down" in shift_direction
        col = shifted(gray, shift_pixel, "down")
    a = get_ic(gray, row, **kw)                                           # :149
    b = get_ic(gray, col, **kw)                                           # :150
    return (a / F32(2) + b / F32(2)).astype(F32)                          # :151


def get_image_change(image_now, image_front, log_add=50, threshold=0.1, clip_range=0.8,
                     return_float=False):
    """create_cityscapes_image_change.py:16-35 -> uint8 ``[H, W]`` (the 'L' PNG
    payload); module globals of lines 169-172 as defaults."""
    front = np.asarray(image_front, dtype=F32)
    now = np.asarray(image_now, dtype=F32)
    with np.errstate(all="ignore"):
        front = np.log(front + log_add)                                   # :17-18
        now = np.log(now + log_add)                                       # :19-20
        d = (now - front).astype(F32)                                     # :21
    d = _dead_zone_split_norm(d, threshold, clip_range)                   # :22-31
    if return_float:
        return d
    return np.uint8(np.around((d + 1) / 2 * 255))                         # :33
