"""Synthetic inputs for the event-representation path.

Everything here is host-side numpy: the generators feed the oracle, the golden
fixtures, the parity tests and ``bench.py`` with the same seeded arrays.  The
dtypes are the DSEC on-disk ones the reference slices out of ``events.h5``
(reference mmseg/datasets/dsec.py:342-345): ``t`` uint32 microseconds sorted
ascending, ``x``/``y`` uint16, ``p`` uint8 in {0, 1}; ``rectify_map`` is float32
``[H, W, 2]`` with channel 0 = x and channel 1 = y (dsec.py:351-353).
"""
from __future__ import annotations

import numpy as np

DSEC_H = 480  # reference mmseg/datasets/dsec.py:160
DSEC_W = 640  # reference mmseg/datasets/dsec.py:161


def seed_for(config: int, sample: int) -> int:
    """Seed convention of SURVEY.md §8(d): ``default_rng(1000*config + sample)``."""
    return 1000 * int(config) + int(sample)


def make_rectify_map(height: int = DSEC_H, width: int = DSEC_W, k1: float = -0.08,
                     jitter: float = 0.25, seed: int = 0) -> np.ndarray:
    """Identity + radial distortion about the image centre + Gaussian jitter.

    Produces non-integer coordinates and a border band that maps outside
    ``[0, W) x [0, H)`` so that the per-corner bounds mask and the
    truncate-toward-zero quirk are exercised.
    """
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64),
                         np.arange(width, dtype=np.float64), indexing="ij")
    cx, cy = width / 2.0, height / 2.0
    nx, ny = (xs - cx) / cx, (ys - cy) / cx
    r2 = nx * nx + ny * ny
    # k1 < 0 pushes border pixels outwards (away from the centre) here, so a
    # band of events lands at negative or >= W/H coordinates.
    scale = 1.0 - k1 * r2
    mx = cx + (xs - cx) * scale + rng.normal(0.0, jitter, size=xs.shape)
    my = cy + (ys - cy) * scale + rng.normal(0.0, jitter, size=ys.shape)
    return np.stack([mx, my], axis=-1).astype(np.float32)


def make_events(n: int, height: int = DSEC_H, width: int = DSEC_W, window_us: int = 50_000,
                t_base: int = 10_000_000, seed: int = 0, skew: float = 0.0):
    """One Poisson-process window conditioned on ``n`` events.

    ``t`` = ``t_base`` + order statistics of U[0, window_us); ``x``/``y`` uniform;
    ``p`` Bernoulli(0.5).  ``skew`` > 0 moves that fraction of the events onto
    1 % of the pixels (moving-edge hot spots, the contention stress variant).
    Returns ``(t uint32, x uint16, y uint16, p uint8)``.
    """
    rng = np.random.default_rng(seed)
    t = np.sort(rng.integers(0, window_us, size=n, dtype=np.int64)).astype(np.uint32)
    t += np.uint32(t_base)
    x = rng.integers(0, width, size=n, dtype=np.int64)
    y = rng.integers(0, height, size=n, dtype=np.int64)
    if skew > 0.0 and n > 0:
        n_hot_px = max(1, (height * width) // 100)
        hot = rng.choice(height * width, size=n_hot_px, replace=False)
        sel = rng.random(n) < skew
        pick = hot[rng.integers(0, n_hot_px, size=int(sel.sum()))]
        x[sel] = pick % width
        y[sel] = pick // width
    p = rng.integers(0, 2, size=n, dtype=np.int64).astype(np.uint8)
    return t, x.astype(np.uint16), y.astype(np.uint16), p


def make_event_store(n_total: int, duration_us: int, height: int = DSEC_H, width: int = DSEC_W,
                     seed: int = 0):
    """A long sorted event stream with DSEC's ``ms_to_idx`` table and ``t_offset``.

    ``ms_to_idx[ms]`` is the index of the first event with ``t >= ms*1000`` (the
    DSEC file-format definition the reference relies on in
    create_dsec_dataset_txt.py:26-35).
    """
    rng = np.random.default_rng(seed)
    t = np.sort(rng.integers(0, duration_us, size=n_total, dtype=np.int64))
    x = rng.integers(0, width, size=n_total, dtype=np.int64).astype(np.uint16)
    y = rng.integers(0, height, size=n_total, dtype=np.int64).astype(np.uint16)
    p = rng.integers(0, 2, size=n_total, dtype=np.int64).astype(np.uint8)
    n_ms = duration_us // 1000 + 1
    ms_to_idx = np.searchsorted(t, np.arange(n_ms, dtype=np.int64) * 1000, side="left").astype(np.int64)
    t_offset = int(rng.integers(1_000_000, 9_000_000))
    return t.astype(np.uint32), x, y, p, ms_to_idx, t_offset


def make_smooth_image(height: int, width: int, seed: int = 0, sigma: float = 8.0,
                      noise: float = 2.0) -> np.ndarray:
    """Gaussian-filtered noise scaled to 0..255 plus N(0, noise^2), uint8 ``[H, W]``."""
    rng = np.random.default_rng(seed)
    f = rng.random((height, width))
    # separable box blur repeated 3x approximates a Gaussian without scipy
    k = max(1, int(round(sigma)))
    for _ in range(3):
        c = np.cumsum(np.pad(f, ((0, 0), (k, k)), mode="wrap"), axis=1)
        f = (c[:, 2 * k:] - c[:, :-2 * k]) / (2 * k)
        c = np.cumsum(np.pad(f, ((k, k), (0, 0)), mode="wrap"), axis=0)
        f = (c[2 * k:, :] - c[:-2 * k, :]) / (2 * k)
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12) * 255.0
    f = f + rng.normal(0.0, noise, size=f.shape)
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def make_frame_pair(height: int, width: int, seed: int = 0):
    """Frame A (smooth field + noise) and frame B = A translated by an integer
    (dx, dy) in [-4, 4]^2 plus fresh noise.  Returns ``(now, front)`` uint8."""
    rng = np.random.default_rng(seed + 7919)
    a = make_smooth_image(height + 8, width + 8, seed=seed, noise=0.0).astype(np.float64)
    dx, dy = (int(v) for v in rng.integers(-4, 5, size=2))
    front = a[4:4 + height, 4:4 + width] + rng.normal(0, 2.0, size=(height, width))
    now = a[4 + dy:4 + dy + height, 4 + dx:4 + dx + width] + rng.normal(0, 2.0, size=(height, width))
    to_u8 = lambda v: np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return to_u8(now), to_u8(front)


def make_rgb_image(height: int, width: int, seed: int = 0) -> np.ndarray:
    """uint8 ``[H, W, 3]`` image made of three smooth fields."""
    return np.stack([make_smooth_image(height, width, seed=seed * 3 + c) for c in range(3)], axis=-1)
