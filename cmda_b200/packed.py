"""The packed event stream ("P4", include/cmda_b200.h): 4 bytes per event instead of the SoA arrays' 9.

The reference slices four HDF5 datasets per window (``events/{t,x,y,p}``, dsec.py:342-345: uint32 + uint16 + uint16
+ uint8).  On a B200 the kernels of this path outrun the PCIe link by an order of magnitude, so the bytes per event
on the wire (and in the decoded-sequence cache of ``store_io``) are what the end-to-end rate is made of.  A record is

    x | y << 11 | p << 21 | sub << 22          sub = t_us - t_base - 1000 * ms

and the millisecond bucket ``ms`` of an event is not stored: it follows from the event's index through
``ms_to_idx`` -- the table DSEC's own events.h5 carries (create_dsec_dataset_txt.py:16, 26-35).  Lossless for
x < 2048, y < 1024, p in {0, 1}, t ascending; the voxel path's results are bit-identical to the SoA entry points.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["pack_p4", "unpack_p4", "ms_table", "pack_p3", "unpack_p3", "sub_table", "p3_to_p4", "PackedEventStore"]

X_BITS, Y_BITS = 11, 10
# the 3-byte WIRE form ("P3", include/cmda_b200.h): x | y << 10 | p << 19 | d << 20 with d = (t - t_base) mod 16; the
# 16-microsecond bucket of an event follows from its index through sub_to_idx as the millisecond does through ms_to_idx
P3_X_BITS, P3_Y_BITS, P3_SUB_US = 10, 9, 16


def ms_table(t, t_base: int, n_ms: int | None = None) -> np.ndarray:
    """``ms_to_idx[k]`` = index of the first event with ``t - t_base >= 1000 k`` for k = 0 .. n_ms; the last entry
    is the number of events."""
    t = np.asarray(t)
    rel_last = int(t[-1]) - int(t_base) if t.size else 0
    if n_ms is None:
        n_ms = rel_last // 1000 + 1
    q = int(t_base) + 1000 * np.arange(n_ms, dtype=np.int64)
    table = np.empty(n_ms + 1, dtype=np.int64)
    table[:n_ms] = np.searchsorted(t.astype(np.int64, copy=False), q, side="left")
    table[n_ms] = t.size
    return table


def pack_p4(t, x, y, p, t_base: int | None = None, check: bool = True):
    """Host-side packer (numpy): ``(rec uint32 [n], ms_to_idx int64 [n_ms + 1], t_base)``.  ``t_base`` defaults to
    the millisecond that holds the first event.  Raises ``ValueError`` when the stream does not fit the format."""
    t = np.ascontiguousarray(t, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.uint16)
    y = np.ascontiguousarray(y, dtype=np.uint16)
    p = np.ascontiguousarray(p, dtype=np.uint8)
    if t_base is None:
        t_base = int(t[0]) // 1000 * 1000 if t.size else 0
    if check:
        if t.size and (int(t[0]) < t_base or np.any(t[1:] < t[:-1])):
            raise ValueError("P4 needs ascending timestamps at or after t_base")
        if x.size and (int(x.max()) >= 1 << X_BITS or int(y.max()) >= 1 << Y_BITS or int(p.max()) > 1):
            raise ValueError("P4 holds x < 2048, y < 1024 and polarity in {0, 1}")
    rel = t - np.uint32(t_base)
    sub = rel % np.uint32(1000)
    rec = x.astype(np.uint32) | (y.astype(np.uint32) << np.uint32(X_BITS)) | (p.astype(np.uint32) << np.uint32(X_BITS + Y_BITS)) | \
        (sub << np.uint32(X_BITS + Y_BITS + 1))
    return rec, ms_table(t, t_base), int(t_base)


def unpack_p4(rec, ms_to_idx, t_base: int):
    """Inverse of :func:`pack_p4` (numpy): ``(t uint32, x uint16, y uint16, p uint8)``."""
    rec = np.asarray(rec, dtype=np.uint32)
    ms_to_idx = np.asarray(ms_to_idx, dtype=np.int64)
    idx = np.arange(rec.size, dtype=np.int64)
    ms = np.searchsorted(ms_to_idx[:-1], idx, side="right") - 1
    t = (np.int64(t_base) + 1000 * ms + (rec >> np.uint32(X_BITS + Y_BITS + 1)).astype(np.int64)).astype(np.uint32)
    return (t, (rec & np.uint32((1 << X_BITS) - 1)).astype(np.uint16),
            ((rec >> np.uint32(X_BITS)) & np.uint32((1 << Y_BITS) - 1)).astype(np.uint16),
            ((rec >> np.uint32(X_BITS + Y_BITS)) & np.uint32(1)).astype(np.uint8))


def sub_table(t, t_base: int, n_sub: int | None = None) -> np.ndarray:
    """``sub_to_idx[k]`` = index of the first event with ``t - t_base >= 16 k`` for k = 0 .. n_sub; the last entry is
    the number of events."""
    t = np.asarray(t)
    rel_last = int(t[-1]) - int(t_base) if t.size else 0
    if n_sub is None:
        n_sub = rel_last // P3_SUB_US + 1
    q = int(t_base) + P3_SUB_US * np.arange(n_sub, dtype=np.int64)
    table = np.empty(n_sub + 1, dtype=np.int64)
    table[:n_sub] = np.searchsorted(t.astype(np.int64, copy=False), q, side="left")
    table[n_sub] = t.size
    return table


def pack_p3(t, x, y, p, t_base: int | None = None, check: bool = True):
    """Host-side packer of the 3-byte wire form: ``(rec3 uint8 [3 n], sub_to_idx int64 [n_sub + 1], t_base)``; ``t_base``
    as in :func:`pack_p4` (the two forms of one stream share it).  Raises ``ValueError`` when the stream does not fit
    (x >= 1024, y >= 512, polarity beyond {0, 1}, descending timestamps)."""
    t = np.ascontiguousarray(t, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.uint16)
    y = np.ascontiguousarray(y, dtype=np.uint16)
    p = np.ascontiguousarray(p, dtype=np.uint8)
    if t_base is None:
        t_base = int(t[0]) // 1000 * 1000 if t.size else 0
    if t_base % 1000:
        raise ValueError("t_base is a whole millisecond (the P4 records the wire form unpacks to count from it)")
    if check:
        if t.size and (int(t[0]) < t_base or np.any(t[1:] < t[:-1])):
            raise ValueError("P3 needs ascending timestamps at or after t_base")
        if x.size and (int(x.max()) >= 1 << P3_X_BITS or int(y.max()) >= 1 << P3_Y_BITS or int(p.max()) > 1):
            raise ValueError("P3 holds x < 1024, y < 512 and polarity in {0, 1}")
    d = (t - np.uint32(t_base)) % np.uint32(P3_SUB_US)
    v = x.astype(np.uint32) | (y.astype(np.uint32) << np.uint32(P3_X_BITS)) | \
        (p.astype(np.uint32) << np.uint32(P3_X_BITS + P3_Y_BITS)) | (d << np.uint32(P3_X_BITS + P3_Y_BITS + 1))
    rec3 = np.ascontiguousarray(v.view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)      # little endian: the low 3 bytes
    return rec3, sub_table(t, t_base), int(t_base)


def _p3_words(rec3) -> np.ndarray:
    b = np.asarray(rec3, dtype=np.uint8).reshape(-1, 3).astype(np.uint32)
    return b[:, 0] | (b[:, 1] << np.uint32(8)) | (b[:, 2] << np.uint32(16))


def unpack_p3(rec3, sub_to_idx, t_base: int):
    """Inverse of :func:`pack_p3` (numpy): ``(t uint32, x uint16, y uint16, p uint8)``."""
    v = _p3_words(rec3)
    sub_to_idx = np.asarray(sub_to_idx, dtype=np.int64)
    j = np.searchsorted(sub_to_idx[:-1], np.arange(v.size, dtype=np.int64), side="right") - 1
    t = (np.int64(t_base) + P3_SUB_US * j + (v >> np.uint32(P3_X_BITS + P3_Y_BITS + 1)).astype(np.int64)).astype(np.uint32)
    return (t, (v & np.uint32((1 << P3_X_BITS) - 1)).astype(np.uint16),
            ((v >> np.uint32(P3_X_BITS)) & np.uint32((1 << P3_Y_BITS) - 1)).astype(np.uint16),
            ((v >> np.uint32(P3_X_BITS + P3_Y_BITS)) & np.uint32(1)).astype(np.uint8))


def p3_to_p4(rec3, sub_to_idx) -> np.ndarray:
    """The P4 records of a P3 stream (numpy restatement of ``cmda_unpack_p3_to_p4``)."""
    v = _p3_words(rec3)
    sub_to_idx = np.asarray(sub_to_idx, dtype=np.int64)
    j = np.searchsorted(sub_to_idx[:-1], np.arange(v.size, dtype=np.int64), side="right") - 1
    sub = ((P3_SUB_US * j + (v >> np.uint32(P3_X_BITS + P3_Y_BITS + 1)).astype(np.int64)) % 1000).astype(np.uint32)
    return (v & np.uint32(1023)) | (((v >> np.uint32(10)) & np.uint32(511)) << np.uint32(X_BITS)) | \
        (((v >> np.uint32(19)) & np.uint32(1)) << np.uint32(X_BITS + Y_BITS)) | (sub << np.uint32(X_BITS + Y_BITS + 1))


class PackedEventStore:
    """A device-resident packed event stream + rectify map(s): the P4 counterpart of ``voxel.EventStore`` (same
    role: the reference's ``self.events_h5`` + ``self.rectify_map``, dsec.py:287-291).  ``events_vg_batch`` accepts
    either."""

    def __init__(self, rec, ms_to_idx, rectify_map=None, height=480, width=640, device=None, plan=True, t_base=0):
        from .voxel import _as_tensor, _cuda_device
        self.device = _cuda_device(device)
        self.height, self.width = int(height), int(width)
        if self.width > 1 << X_BITS or self.height > 1 << Y_BITS:
            raise ValueError("P4 holds x < 2048 and y < 1024")
        if isinstance(rec, torch.Tensor):
            if rec.dtype not in (torch.uint32, torch.int32):
                raise TypeError(f"packed records are uint32, got {rec.dtype}")
            self.rec = rec.view(torch.uint32).to(self.device, non_blocking=True).contiguous()
        else:
            self.rec = torch.from_numpy(np.ascontiguousarray(rec, dtype=np.uint32)).to(self.device, non_blocking=True)
        self.h_ms_to_idx = np.ascontiguousarray(ms_to_idx.cpu().numpy() if isinstance(ms_to_idx, torch.Tensor) else ms_to_idx,
                                                dtype=np.int64)
        if self.h_ms_to_idx.ndim != 1 or self.h_ms_to_idx.size < 2 or self.h_ms_to_idx[0] != 0 or \
                self.h_ms_to_idx[-1] != self.rec.shape[0] or np.any(np.diff(self.h_ms_to_idx) < 0):
            raise ValueError("ms_to_idx must ascend from 0 to the number of events")
        self.ms_to_idx = torch.from_numpy(self.h_ms_to_idx).to(self.device, non_blocking=True)
        self.n_ms = int(self.h_ms_to_idx.size - 1)
        self.t_base = int(t_base)
        self.rectify_map = None
        if rectify_map is not None:
            m = _as_tensor(rectify_map, torch.float32, self.device)
            if m.ndim == 3:
                m = m[None]
            assert m.shape[1:] == (self.height, self.width, 2), "rectify_map is [H, W, 2] (dsec.py:351-353)"
            self.rectify_map = m.contiguous()
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
        return int(self.rec.shape[0])

    @classmethod
    def from_event_store(cls, store, t_base: int | None = None, plan=True):
        """Pack a device-resident SoA ``EventStore`` on the device (``cmda_pack_events_p4``).  Raises ``ValueError``
        when the stream does not fit the format (polarity bytes beyond {0, 1}, descending timestamps, ...)."""
        L = _lib.lib()
        n = len(store)
        dev = store.device
        if n == 0:
            raise ValueError("empty store")
        first, last = int(store.t[0].item()), int(store.t[-1].item())
        if t_base is None:
            t_base = first // 1000 * 1000
        n_ms = (last - t_base) // 1000 + 1
        if first < t_base or n_ms < 1:
            raise ValueError("P4 needs ascending timestamps at or after t_base")
        rec = torch.empty((n,), dtype=torch.uint32, device=dev)
        table = torch.empty((n_ms + 1,), dtype=torch.int64, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with _lib.on_device(dev):
            _lib.check(L.cmda_pack_events_p4(_lib.ptr(store.t), _lib.ptr(store.x), _lib.ptr(store.y), _lib.ptr(store.p), n,
                                             t_base, n_ms, _lib.ptr(rec), _lib.ptr(table), _lib.ptr(status),
                                             _lib.stream_ptr(dev)), "cmda_pack_events_p4")
        bad = int(status.item())
        if bad:
            raise ValueError(f"{bad} events do not fit the P4 format (x < 2048, y < 1024, polarity in {{0, 1}}, t ascending)")
        obj = cls(rec, table, None, height=store.height, width=store.width, device=dev, plan=False, t_base=t_base)
        obj.rectify_map, obj.plans = store.rectify_map, store.plans
        if plan and obj.plans is None and obj.rectify_map is not None:
            obj2 = cls(rec, obj.h_ms_to_idx, obj.rectify_map, height=store.height, width=store.width, device=dev, plan=True,
                       t_base=t_base)
            return obj2
        return obj
