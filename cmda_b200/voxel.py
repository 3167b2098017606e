"""Event windows -> voxel grids: the reference's call signatures over the CUDA kernels.

Mirrors ``events_to_voxel_grid`` / ``events_norm`` / ``DSECDataset.get_events_vg``
(reference mmseg/datasets/dsec.py:26-121, 341-366).  Outputs live where the inputs
live: CUDA tensors in -> CUDA tensor out; host tensors / numpy arrays in -> they are
copied to the current CUDA device, the kernels run there, and the result is copied back
(the reference's CPU-tensor contract).  ``out_device`` overrides that.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["events_to_voxel_grid", "events_norm", "EventStore", "events_vg_batch", "events_vg_augmented_batch",
           "remap_events", "default_clip_range"]


def _cuda_device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.CmdaError("no CUDA device: the cmda_b200 path has no CPU implementation")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.CmdaError("cmda_b200 computes on CUDA devices only")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def _as_tensor(a, dtype, device) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def _home_of(t) -> torch.device:
    return t.device if isinstance(t, torch.Tensor) else torch.device("cpu")


def default_clip_range(events_finish_index: int, events_start_index: int) -> float:
    """reference dsec.py:362."""
    return (events_finish_index - events_start_index) / 500000 * 1.5


def events_to_voxel_grid(time, x, y, pol, width, height, num_bins, normalize_flag=False, *, mode="auto",
                         out_device=None, return_bin_counts=False):
    """Drop-in for ``events_to_voxel_grid`` (reference dsec.py:26-70).

    float32 1-D ``time, x, y, pol`` -> float32 ``[num_bins, height, width]`` on ``pol``'s
    device.  ``normalize_flag=True`` (dsec.py:60-68, never used by the reference) is applied
    with plain torch ops on the result.
    """
    assert x.shape == y.shape == pol.shape == time.shape      # dsec.py:28
    assert x.ndim == 1                                         # dsec.py:29
    home = _home_of(pol)
    dev = _cuda_device(home if home.type == "cuda" else None)
    t_, x_, y_, p_ = (_as_tensor(a, torch.float32, dev) for a in (time, x, y, pol))
    n = int(t_.shape[0])
    if n == 0:
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")   # t_norm[0], dsec.py:39
    mode_id = _lib.VOXEL_MODES[mode]
    L = _lib.lib()
    grid = torch.empty((num_bins, height, width), dtype=torch.float32, device=dev)
    counts = torch.empty((num_bins,), dtype=torch.int64, device=dev) if return_bin_counts else None
    with _lib.on_device(dev):
        nbytes = L.cmda_events_vg_workspace_bytes(n, 1, height, width, num_bins, mode_id)
        ws = _lib.workspace(dev, nbytes)
        _lib.check(L.cmda_voxel_grid_f32(_lib.ptr(t_), _lib.ptr(x_), _lib.ptr(y_), _lib.ptr(p_), n, width, height,
                                         num_bins, _lib.ptr(grid), _lib.ptr(counts), _lib.ptr(ws), ws.numel(),
                                         mode_id, _lib.stream_ptr(dev)), "cmda_voxel_grid_f32")
    if normalize_flag:                                         # dsec.py:60-68
        mask = torch.nonzero(grid, as_tuple=True)
        if mask[0].size()[0] > 0:
            mean, std = grid[mask].mean(), grid[mask].std()
            grid[mask] = (grid[mask] - mean) / std if std > 0 else grid[mask] - mean
    out = grid.to(out_device if out_device is not None else home)
    return (out, counts.to(out.device)) if return_bin_counts else out


def events_norm(events, clip_range=1.0, final_range=1.0, enforce_no_events_zero=False, *, out_device=None):
    """Drop-in for ``events_norm`` (reference dsec.py:80-121), numeric ``clip_range``.

    Accepts ``[..., H, W]`` grids of one window (statistics are global over the whole
    tensor, as in the reference).  Returns a new tensor on the input's device.
    """
    if isinstance(clip_range, str):
        # dsec.py:84-86 -- commented out at its only call site (dsec.py:363); not built
        raise NotImplementedError("clip_range='auto' is not used by the reference's hot path")
    home = _home_of(events)
    dev = _cuda_device(home if home.type == "cuda" else None)
    g = _as_tensor(events, torch.float32, dev).clone()
    L = _lib.lib()
    clip = np.array([clip_range], dtype=np.float32)
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_events_norm_workspace_bytes(1))
        _lib.check(L.cmda_events_norm_batch(_lib.ptr(g), 1, g.numel(), _lib.host_ptr(clip), float(final_range),
                                            int(bool(enforce_no_events_zero)), _lib.ptr(ws), ws.numel(),
                                            _lib.stream_ptr(dev)), "cmda_events_norm_batch")
    return g.to(out_device if out_device is not None else home)


class EventStore:
    """A device-resident DSEC event stream: SoA arrays in their on-disk dtypes plus the
    rectify map(s).  Stands in for the reference's ``self.events_h5`` + ``self.rectify_map``
    (dsec.py:287-291); decoding events.h5 itself is file I/O and out of scope."""

    def __init__(self, t, x, y, p, rectify_map=None, height=480, width=640, device=None, plan=True):
        self.device = _cuda_device(device)
        def put(a, np_dtype, torch_dtype):
            if isinstance(a, torch.Tensor):
                if a.dtype != torch_dtype:
                    # a same-size signed view (torch has few unsigned ops) is accepted when no value is negative:
                    # its bits are then the unsigned value; anything else would be silently reinterpreted
                    signed = {torch.uint32: torch.int32, torch.uint16: torch.int16}.get(torch_dtype)
                    if a.dtype != signed:
                        raise TypeError(f"DSEC dtype {torch_dtype} expected, got {a.dtype}")
                    if a.numel() and int(a.min()) < 0:
                        raise ValueError(f"negative values in a {a.dtype} array passed for {torch_dtype}")
                    a = a.view(torch_dtype)
                return a.to(self.device, non_blocking=True).contiguous()
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype)).to(self.device, non_blocking=True)

        self.t = put(t, np.uint32, torch.uint32)
        self.x = put(x, np.uint16, torch.uint16)
        self.y = put(y, np.uint16, torch.uint16)
        self.p = put(p, np.uint8, torch.uint8)
        for a, sz in ((self.t, 4), (self.x, 2), (self.y, 2), (self.p, 1)):
            assert a.element_size() == sz and a.ndim == 1 and a.is_contiguous()
        assert self.t.shape == self.x.shape == self.y.shape == self.p.shape
        self.height, self.width = int(height), int(width)
        self.rectify_map = None
        if rectify_map is not None:
            m = _as_tensor(rectify_map, torch.float32, self.device)
            if m.ndim == 3:
                m = m[None]
            assert m.shape[1:] == (self.height, self.width, 2), "rectify_map is [H, W, 2] (dsec.py:351-353)"
            self.rectify_map = m.contiguous()
        # the maps are static for a sequence: their gather plans (inverse index, stencil, tile boxes) are
        # built once here instead of on every voxel call
        self.plans = None
        if self.rectify_map is not None and plan:
            L = _lib.lib()
            nbytes = L.cmda_rectify_plan_bytes(self.height, self.width)
            if nbytes:
                n_maps = int(self.rectify_map.shape[0])
                self.plans = torch.empty((n_maps * nbytes,), dtype=torch.uint8, device=self.device)
                with torch.cuda.device(self.device):
                    _lib.check(L.cmda_rectify_plan_build(_lib.ptr(self.rectify_map), n_maps, self.height, self.width,
                                                         _lib.ptr(self.plans), _lib.stream_ptr(self.device)),
                               "cmda_rectify_plan_build")

    def __len__(self):
        return int(self.t.shape[0])


def _check_windows(store, starts, finishes, clip_ranges, map_ids):
    """Validation shared by the batched entry points: inclusive windows inside the store, per-window clip ranges
    (``None`` -> the reference's default, dsec.py:362), map ids inside the store's rectify maps (and its plans).
    The C ABI cannot know how many maps / plans the device arrays hold, so the range check lives here."""
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    ends = np.ascontiguousarray(finishes, dtype=np.int64) + 1
    S = int(starts.shape[0])
    if ends.shape != starts.shape or starts.ndim != 1:
        raise ValueError("starts / finishes must be 1-D and of equal length")
    if S and (starts.min() < 0 or ends.max() > len(store)):
        raise IndexError("event window outside the store")
    clips = np.empty(S, dtype=np.float32)
    for s in range(S):
        c = None if clip_ranges is None else clip_ranges[s]
        clips[s] = default_clip_range(int(ends[s]) - 1, int(starts[s])) if c is None else c
    mids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
    if mids is not None:
        if mids.shape != (S,):
            raise ValueError("map_ids must hold one id per window")
        n_maps = 0 if store.rectify_map is None else int(store.rectify_map.shape[0])
        if S and store.rectify_map is not None and (mids.min() < 0 or mids.max() >= n_maps):
            raise IndexError("map_id outside the store's rectify maps")
    return starts, ends, S, clips, mids


def events_vg_batch(store, starts, finishes, num_bins, clip_ranges=None, *, map_ids=None,
                    normalize=True, final_range=1.0, enforce_no_events_zero=True, mode="auto", out=None,
                    return_raw=False, return_bin_counts=False):
    """S windows ``[start, finish]`` (INCLUSIVE, as in dsec.py:342-345) of one store ->
    ``[S, num_bins, H, W]`` float32 on the store's device: ``get_events_vg`` batched.  ``store`` is an
    ``EventStore`` (SoA arrays in DSEC dtypes) or a ``packed.PackedEventStore`` (4 bytes per event).

    ``clip_ranges[s] is None`` (or ``clip_ranges is None``) selects the reference's default
    ``(finish - start) / 500000 * 1.5`` (dsec.py:362).

    An EMPTY window (``finish < start``) is legal in a batch: its raw grid is all zero and ``events_norm`` of an
    all-zero grid follows (with ``enforce_no_events_zero`` every voxel is the normalised value of 0).  The
    reference never gets there: ``__getitem__`` returns ``None`` for ``start > finish`` (dsec.py:301-302, mirrored by
    ``DSECEvents.events_vg_for_image``) and ``get_events_vg`` itself would raise on ``events_t[0]`` (mirrored by
    ``DSECEvents.get_events_vg``).
    """
    L = _lib.lib()
    dev = store.device
    starts, ends, S, clips, mids = _check_windows(store, starts, finishes, clip_ranges, map_ids)
    H, W, B = store.height, store.width, int(num_bins)
    mode_id = _lib.VOXEL_MODES[mode]
    if out is None:
        out = torch.empty((S, B, H, W), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.is_contiguous() and out.shape == (S, B, H, W) and out.dtype == torch.float32
    raw = torch.empty_like(out) if (return_raw and normalize) else None
    counts = torch.empty((S, B), dtype=torch.int64, device=dev) if return_bin_counts else None
    total = int(np.clip(ends - starts, 0, None).sum())
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_events_vg_workspace_bytes(total, S, H, W, B, mode_id))
        if hasattr(store, "rec"):        # packed.PackedEventStore: 4 bytes per event, same results bit for bit
            _lib.check(L.cmda_events_vg_batch_p4(
                _lib.ptr(store.rec), _lib.ptr(store.ms_to_idx), _lib.host_ptr(store.h_ms_to_idx), store.n_ms,
                _lib.host_ptr(starts), _lib.host_ptr(ends), None, S, _lib.ptr(store.rectify_map), _lib.host_ptr(mids), H, W, B,
                _lib.host_ptr(clips), float(final_range), int(bool(enforce_no_events_zero)), int(bool(normalize)), _lib.ptr(out),
                _lib.ptr(raw), _lib.ptr(counts), _lib.ptr(ws), ws.numel(), mode_id, _lib.ptr(store.plans), _lib.stream_ptr(dev)),
                "cmda_events_vg_batch_p4")
        else:
            _lib.check(L.cmda_events_vg_batch_planned(
                _lib.ptr(store.t), _lib.ptr(store.x), _lib.ptr(store.y), _lib.ptr(store.p), _lib.host_ptr(starts),
                _lib.host_ptr(ends), S, _lib.ptr(store.rectify_map), _lib.host_ptr(mids), H, W, B, _lib.host_ptr(clips),
                float(final_range), int(bool(enforce_no_events_zero)), int(bool(normalize)), _lib.ptr(out), _lib.ptr(raw),
                _lib.ptr(counts), _lib.ptr(ws), ws.numel(), mode_id, _lib.ptr(store.plans), _lib.stream_ptr(dev)),
                "cmda_events_vg_batch_planned")
    res = [out]
    if return_raw:
        res.append(raw if normalize else out)
    if return_bin_counts:
        res.append(counts)
    return res[0] if len(res) == 1 else tuple(res)


def events_vg_augmented_batch(store: EventStore, starts, finishes, num_bins, clip_ranges=None, *, crop_xy, crop_size,
                              out_size, flips=None, avg_bins=False, repeat=1, map_ids=None, final_range=1.0,
                              enforce_no_events_zero=True, mode="auto", out=None):
    """``get_events_vg`` + the post-voxel stage of ``DSECDataset.__getitem__`` (reference dsec.py:304-319) in
    one call: mean over bins (``events_bins_5_avg_1``) -> crop ``crop_size = (w, h)`` at ``crop_xy[s] = (x, y)``
    -> horizontal flip -> bilinear resize to ``out_size = (w, h)`` (``F.interpolate(align_corners=False)``) ->
    ``repeat(repeat, 1, 1)``.  The normaliser applies the augmentation itself: the normalised full grid is
    never written.  Test mode of the reference (dsec.py:316-317) is ``crop_xy=(0, 0)``,
    ``crop_size = out_size = (W, 440)``.  Returns ``[S, repeat * Bo, out_h, out_w]`` float32 on the device."""
    L = _lib.lib()
    dev = store.device
    starts, ends, S, clips, mids = _check_windows(store, starts, finishes, clip_ranges, map_ids)
    H, W, B = store.height, store.width, int(num_bins)
    cw, ch = (int(v) for v in crop_size)
    ow, oh = (int(v) for v in out_size)
    aug = np.zeros((S, 3), dtype=np.int32)
    xy = np.asarray(crop_xy, dtype=np.int64).reshape(-1, 2)
    aug[:, 0:2] = xy if xy.shape[0] == S else np.broadcast_to(xy, (S, 2))
    if flips is not None:
        aug[:, 2] = np.asarray(flips, dtype=np.int32).reshape(-1)
    if S and (aug[:, 0].min() < 0 or aug[:, 1].min() < 0 or (aug[:, 0] + cw).max() > W or (aug[:, 1] + ch).max() > H):
        raise IndexError("crop outside the voxel grid")
    Bo = 1 if avg_bins else B
    mode_id = _lib.VOXEL_MODES[mode]
    if out is None:
        out = torch.empty((S, int(repeat) * Bo, oh, ow), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.is_contiguous() and out.shape == (S, int(repeat) * Bo, oh, ow)
    total = int(np.clip(ends - starts, 0, None).sum())
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_events_vg_augmented_workspace_bytes(total, S, H, W, B, mode_id))
        _lib.check(L.cmda_events_vg_augmented_batch(
            _lib.ptr(store.t), _lib.ptr(store.x), _lib.ptr(store.y), _lib.ptr(store.p), _lib.host_ptr(starts),
            _lib.host_ptr(ends), S, _lib.ptr(store.rectify_map), _lib.host_ptr(mids), H, W, B, _lib.host_ptr(clips),
            float(final_range), int(bool(enforce_no_events_zero)), _lib.host_ptr(aug), cw, ch, ow, oh, int(bool(avg_bins)),
            int(repeat), _lib.ptr(out), None, None, _lib.ptr(ws), ws.numel(), mode_id, _lib.ptr(store.plans),
            _lib.stream_ptr(dev)), "cmda_events_vg_augmented_batch")
    return out


def remap_events(store: EventStore, start: int, finish: int, num_bins: int, map_id: int = 0):
    """Integer side outputs of one window ``[start, finish]``: rectified float coordinates,
    ``t_norm`` and the truncated corner origin (dsec.py:41-43, 347-355) per event."""
    L = _lib.lib()
    dev = store.device
    n = finish - start + 1
    f = lambda dt: torch.empty((max(n, 0),), dtype=dt, device=dev)
    xr, yr, tn = f(torch.float32), f(torch.float32), f(torch.float32)
    x0, y0, t0 = f(torch.int32), f(torch.int32), f(torch.int32)
    rmap = None if store.rectify_map is None else store.rectify_map[map_id]
    with _lib.on_device(dev):
        _lib.check(L.cmda_remap_events(_lib.ptr(store.t), _lib.ptr(store.x), _lib.ptr(store.y), _lib.ptr(store.p),
                                       start, finish + 1, _lib.ptr(rmap), store.height, store.width, num_bins,
                                       _lib.ptr(xr), _lib.ptr(yr), _lib.ptr(tn), _lib.ptr(x0), _lib.ptr(y0),
                                       _lib.ptr(t0), _lib.stream_ptr(dev)), "cmda_remap_events")
    return dict(x=xr, y=yr, t_norm=tn, x0=x0, y0=y0, t0=t0)
