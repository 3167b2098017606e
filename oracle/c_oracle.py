"""ctypes binding of oracle/cmda_oracle.c.  TEST INFRASTRUCTURE ONLY (see that file).

The reference is pure Python (no native sources to compile into ``oracle/_ref``), so
this C restatement is the fast, multi-window CPU checker and the ``"port"`` CPU
baseline of ``bench.py``.  LUTs are always computed by numpy on the caller's side,
exactly as the reference evaluates ``np.log``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import cmda_oracle as _py

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcmda_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cmda_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def voxel_grid(time, x, y, pol, width, height, bins) -> np.ndarray:
    time, x, y, pol = (np.ascontiguousarray(a, dtype=np.float32) for a in (time, x, y, pol))
    grid = np.empty((bins, height, width), np.float32)
    lib().oracle_voxel_grid(_p(time, _f32p), _p(x, _f32p), _p(y, _f32p), _p(pol, _f32p),
                            ctypes.c_int64(time.shape[0]), width, height, bins, _p(grid, _f32p))
    return grid


def events_norm(events, clip_range, final_range=1.0, enforce_no_events_zero=False) -> np.ndarray:
    ev = np.array(events, dtype=np.float32, copy=True, order="C")
    scratch = np.empty_like(ev)
    lib().oracle_events_norm(_p(ev, _f32p), ctypes.c_int64(ev.size), ctypes.c_float(clip_range),
                             ctypes.c_float(final_range), int(bool(enforce_no_events_zero)), _p(scratch, _f32p))
    return ev


def get_events_vg_batch(t, x, y, p, starts, finishes, rectify_map, width, height, bins, clips=None,
                        return_raw=False, nthreads=0):
    t = np.ascontiguousarray(t, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.uint16)
    y = np.ascontiguousarray(y, dtype=np.uint16)
    p = np.ascontiguousarray(p, dtype=np.uint8)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    finishes = np.ascontiguousarray(finishes, dtype=np.int64)
    S = starts.shape[0]
    rmap = np.ascontiguousarray(rectify_map, dtype=np.float32) if rectify_map is not None else None
    clip_arr = np.full(S, -1.0) if clips is None else np.array([-1.0 if c is None else c for c in clips], np.float64)
    out = np.empty((S, bins, height, width), np.float32)
    raw = np.empty_like(out) if return_raw else None
    i64p = ctypes.POINTER(ctypes.c_int64)
    err = lib().oracle_get_events_vg_batch(
        _p(t, ctypes.POINTER(ctypes.c_uint32)), _p(x, ctypes.POINTER(ctypes.c_uint16)),
        _p(y, ctypes.POINTER(ctypes.c_uint16)), _p(p, _u8p), _p(starts, i64p), _p(finishes, i64p), S,
        _p(rmap, _f32p), width, height, bins, _p(clip_arr, ctypes.POINTER(ctypes.c_double)),
        _p(out, _f32p), _p(raw, _f32p), int(nthreads))
    if err:
        raise MemoryError("oracle_get_events_vg_batch")
    return (out, raw) if return_raw else out


def voxel_aux_batch(t, x, y, p, starts, finishes, rectify_map, width, height, bins, nthreads=0):
    """Per voxel the sum of |w| (float64) and the number of contributions of each window: the tolerance inputs
    of the raw-grid comparisons (tests/test_gpu_parity.py), at sizes where numpy takes minutes."""
    t = np.ascontiguousarray(t, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.uint16)
    y = np.ascontiguousarray(y, dtype=np.uint16)
    p = np.ascontiguousarray(p, dtype=np.uint8)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    finishes = np.ascontiguousarray(finishes, dtype=np.int64)
    S = starts.shape[0]
    rmap = np.ascontiguousarray(rectify_map, dtype=np.float32) if rectify_map is not None else None
    abs_w = np.empty((S, bins, height, width), np.float64)
    n_contrib = np.empty((S, bins, height, width), np.int32)
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib().oracle_voxel_aux_batch(
        _p(t, ctypes.POINTER(ctypes.c_uint32)), _p(x, ctypes.POINTER(ctypes.c_uint16)),
        _p(y, ctypes.POINTER(ctypes.c_uint16)), _p(p, _u8p), _p(starts, i64p), _p(finishes, i64p), S,
        _p(rmap, _f32p), width, height, bins, _p(abs_w, ctypes.POINTER(ctypes.c_double)),
        _p(n_contrib, ctypes.POINTER(ctypes.c_int32)), int(nthreads))
    return abs_w, n_contrib


_DIR_MODE = {"rightdown": 0, "rightup": 1, "leftdown": 2, "leftup": 3, "all": 4}


def isr_batch(gray, shift_pixel, val_range, threshold, clip_range, shift_direction="rightdown", nthreads=0):
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    if gray.ndim == 2:
        gray = gray[None]
    S, H, W = gray.shape
    lut = _py.log_lut_val_range(val_range)
    span = np.log(val_range[1]) - np.log(val_range[0])
    out = np.empty((S, H, W), np.float32)
    err = lib().oracle_isr_batch(_p(gray, _u8p), S, H, W, int(shift_pixel), _DIR_MODE[shift_direction],
                                 _p(lut, _f32p), ctypes.c_float(np.float32(span * threshold)),
                                 ctypes.c_float(np.float32(span * clip_range)), _p(out, _f32p), int(nthreads))
    if err:
        raise MemoryError("oracle_isr_batch")
    return out


def image_change_batch(now, front, log_add=50, threshold=0.1, clip_range=0.8, want_u8=True, nthreads=0):
    now = np.ascontiguousarray(now, dtype=np.uint8)
    front = np.ascontiguousarray(front, dtype=np.uint8)
    if now.ndim == 2:
        now, front = now[None], front[None]
    S, H, W = now.shape
    lut = _py.log_lut_log_add(log_add)
    out_f = np.empty((S, H, W), np.float32)
    out_u = np.empty((S, H, W), np.uint8) if want_u8 else None
    err = lib().oracle_image_change_batch(_p(now, _u8p), _p(front, _u8p), S, H, W, _p(lut, _f32p),
                                          ctypes.c_float(np.float32(threshold)), ctypes.c_float(np.float32(clip_range)),
                                          _p(out_f, _f32p), _p(out_u, _u8p), int(nthreads))
    if err:
        raise MemoryError("oracle_image_change_batch")
    return (out_f, out_u) if want_u8 else out_f
