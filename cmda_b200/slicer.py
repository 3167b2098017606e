"""Event-window slicer: image timestamps -> event indices on the device (K1).

Mirrors ``create_images_to_events_index`` (reference create_dsec_dataset_txt.py:10-47) on
in-memory arrays: the ``t`` column stays resident on the GPU, every image timestamp is one
bracketed binary search.  Reading events.h5 / writing the .txt table is file I/O and is
left to the caller (``write_index_txt`` reproduces the reference's file format).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .voxel import _cuda_device

__all__ = ["searchsorted_right", "images_to_events_index", "write_index_txt", "window_bounds"]


def _dev_tensor(a, np_dtype, dev):
    if isinstance(a, torch.Tensor):
        return a.to(dev, non_blocking=True).contiguous()
    a = np.ascontiguousarray(a, dtype=np_dtype)
    if not a.flags.writeable:          # e.g. a memory-mapped cache file: torch wants a writable buffer to wrap
        a = a.copy()
    return torch.from_numpy(a).to(dev, non_blocking=True)


def searchsorted_right(t, queries, *, device=None) -> torch.Tensor:
    """``np.searchsorted(t, q, 'right')`` for an ascending uint32 ``t`` (device resident)
    and int64 queries -> int64 tensor on the device."""
    dev = t.device if isinstance(t, torch.Tensor) and t.is_cuda else _cuda_device(device)
    t_d = _dev_tensor(t, np.uint32, dev)
    q_d = _dev_tensor(queries, np.int64, dev)
    out = torch.empty(q_d.shape, dtype=torch.int64, device=dev)
    with _lib.on_device(dev):
        _lib.check(_lib.lib().cmda_searchsorted_right_u32(_lib.ptr(t_d), t_d.numel(), _lib.ptr(q_d), q_d.numel(),
                                                          _lib.ptr(out), _lib.stream_ptr(dev)),
                   "cmda_searchsorted_right_u32")
    return out


def images_to_events_index(t, t_offset, ms_to_idx, images_timestamps, *, device=None) -> list:
    """Per image timestamp: index of the last event with ``t <= ts - t_offset`` or -1
    (create_dsec_dataset_txt.py:19-42).  Raises ``ValueError('range error!')`` where the
    reference does (line 37-39).  Returns a Python list of ints like the reference builds."""
    dev = t.device if isinstance(t, torch.Tensor) and t.is_cuda else _cuda_device(device)
    t_d = _dev_tensor(t, np.uint32, dev)
    ms_d = _dev_tensor(ms_to_idx, np.int64, dev)
    ts_d = _dev_tensor(images_timestamps, np.int64, dev)
    n_ts = int(ts_d.numel())
    index = torch.empty((n_ts,), dtype=torch.int64, device=dev)
    status = torch.empty((n_ts,), dtype=torch.int32, device=dev)
    with _lib.on_device(dev):
        _lib.check(_lib.lib().cmda_images_to_events_index(_lib.ptr(t_d), t_d.numel(), _lib.ptr(ms_d), ms_d.numel(),
                                                          int(t_offset), _lib.ptr(ts_d), n_ts, _lib.ptr(index),
                                                          _lib.ptr(status), _lib.stream_ptr(dev)),
                   "cmda_images_to_events_index")
    status_h = status.cpu().numpy()
    if (status_h == 1).any():
        raise ValueError('range error!')                        # create_dsec_dataset_txt.py:39
    if (status_h == 2).any():
        raise IndexError('index out of bounds for ms_to_idx')   # ms_to_idx[timestamps_ms + 2], line 33
    return [int(v) for v in index.cpu().numpy()]


def write_index_txt(index_list, output_txt_path: str) -> None:
    """The reference's output format (create_dsec_dataset_txt.py:44-47): one integer per line."""
    with open(output_txt_path, 'w', encoding='UTF-8') as f:
        for v in index_list:
            f.write(str(int(v)) + '\n')


def window_bounds(images_to_events_index, now_image_index, image_change_range=1, events_num=-1, i=0):
    """Inclusive ``(start, finish)`` of output window ``i`` (reference dsec.py:296-302), or
    ``None`` when ``start > finish`` (where ``__getitem__`` returns None)."""
    finish = int(images_to_events_index[now_image_index - i])
    if events_num != -1:
        start = finish - events_num + 1
    else:
        start = int(images_to_events_index[now_image_index - image_change_range - i])
    if start > finish:
        return None
    return start, finish
