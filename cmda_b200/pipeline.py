"""Host-buffer front door of the voxel path: pinned host events in, host grids out.

This is what a DataLoader-side caller uses when the event arrays live in host memory (the
reference reads them from events.h5 into numpy, dsec.py:342-345): windows are cut into
groups, each group's slices are copied host->device on a copy stream while the previous
group is voxelised on the compute stream and the group before that is copied back.  The
kernels only ever see device memory; the overlap is plain CUDA streams + events.

The host->device copy of the events is what bounds this path end to end (the kernels run an order
of magnitude faster than PCIe delivers their input), so the wire format matters more than any
kernel: ``wire="soa"`` ships the four DSEC arrays as they are (9 bytes per event), ``wire="p4"``
ships the packed stream of ``cmda_b200.packed`` (4 bytes per event, packed ONCE when the pipeline --
or the decoded-sequence cache of ``store_io`` -- is built; bit-identical results), ``wire="p3"`` its 3-byte
wire form (sensors up to 1024 x 512; unpacked to P4 records on the device by ``cmda_unpack_p3_to_p4``,
0.17 ms per 80 M events, before the same kernels run).  Windows of a group that touch or overlap in the
store travel as one copy per array.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import packed as _packed
from .voxel import _cuda_device, default_clip_range

__all__ = ["HostEventsPipeline"]

_ALIGN = 64          # events: every copied range starts on a 64-event boundary of the staging buffer (128-bit loads)


class HostEventsPipeline:
    """Voxelise windows of a HOST-resident event store.

    ``t, x, y, p`` are numpy arrays or CPU tensors in DSEC dtypes; they are wrapped (and
    pinned once, if they are not already) so that every later call is pure DMA.  With
    ``wire="p4"`` they are packed once here (or pass ``packed=(rec, ms_to_idx)`` from a cache) and
    only the packed stream is kept; ``wire="p3"`` likewise (``packed=(rec3, sub_to_idx, ms_to_idx)``).  The
    rectify map is uploaded once.
    """

    def __init__(self, t, x, y, p, rectify_map, num_bins, height=480, width=640, device=None,
                 windows_per_group=4, mode="auto", max_window_events=None, wire="soa", packed=None):
        self.device = _cuda_device(device)
        self.H, self.W, self.B = int(height), int(width), int(num_bins)
        self.mode = mode
        self.group = int(windows_per_group)
        assert wire in ("soa", "p4", "p3")
        self.wire = wire
        if wire == "p3":
            if packed is None:
                tt = self._np(t, np.uint32)
                rec3, sub, t_base = _packed.pack_p3(tt, self._np(x, np.uint16), self._np(y, np.uint16), self._np(p, np.uint8))
                table = _packed.ms_table(tt, t_base)
            else:
                rec3, sub, table = packed
            if self.W > 1 << _packed.P3_X_BITS or self.H > 1 << _packed.P3_Y_BITS:
                raise ValueError("the P3 wire holds x < 1024 and y < 512: use wire='p4'")
            self.host = [self._pin(rec3, np.uint8)]
            self.h_sub_to_idx = np.ascontiguousarray(sub, dtype=np.int64)
            self.d_sub_to_idx = torch.from_numpy(self.h_sub_to_idx).to(self.device)
            self.h_ms_to_idx = np.ascontiguousarray(table, dtype=np.int64)
            self.d_ms_to_idx = torch.from_numpy(self.h_ms_to_idx).to(self.device)
            self.bytes_per_event = 3
        elif wire == "p4":
            if packed is None:
                rec, table, _ = _packed.pack_p4(self._np(t, np.uint32), self._np(x, np.uint16), self._np(y, np.uint16),
                                                self._np(p, np.uint8))
            else:
                rec, table = packed
            self.host = [self._pin(rec, np.uint32)]
            self.h_ms_to_idx = np.ascontiguousarray(table, dtype=np.int64)
            self.d_ms_to_idx = torch.from_numpy(self.h_ms_to_idx).to(self.device)
            self.bytes_per_event = 4
        else:
            self.host = [self._pin(a, dt) for a, dt in ((t, np.uint32), (x, np.uint16), (y, np.uint16), (p, np.uint8))]
            self.bytes_per_event = 9
        self.n_total = int(self.host[0].shape[0]) // (3 if wire == "p3" else 1)
        self.rmap = None
        if rectify_map is not None:
            m = torch.as_tensor(np.ascontiguousarray(rectify_map, dtype=np.float32))
            self.rmap = m.to(self.device).reshape(-1, self.H, self.W, 2).contiguous()
        self.plans = None
        if self.rmap is not None:        # map-derived gather plans: built once, the map is static
            L = _lib.lib()
            nbytes = L.cmda_rectify_plan_bytes(self.H, self.W)
            if nbytes:
                self.plans = torch.empty((int(self.rmap.shape[0]) * nbytes,), dtype=torch.uint8, device=self.device)
                with torch.cuda.device(self.device):
                    _lib.check(L.cmda_rectify_plan_build(_lib.ptr(self.rmap), int(self.rmap.shape[0]), self.H, self.W,
                                                         _lib.ptr(self.plans), _lib.stream_ptr(self.device)),
                               "cmda_rectify_plan_build")
                torch.cuda.current_stream(self.device).synchronize()
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self.compute = torch.cuda.Stream(self.device)
        self._cap = 0
        self._slots = []
        self.last_h2d_bytes = 0
        if max_window_events:
            self._ensure(int(max_window_events) * self.group + 2 * _ALIGN * self.group)

    @staticmethod
    def _np(a, dt):
        return a.numpy() if isinstance(a, torch.Tensor) else np.ascontiguousarray(a, dtype=dt)

    @staticmethod
    def _pin(a, np_dtype):
        if isinstance(a, torch.Tensor):
            tns = a.contiguous()
        else:
            tns = torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype))
        assert tns.element_size() == np.dtype(np_dtype).itemsize and tns.ndim == 1
        return tns if tns.is_pinned() else tns.pin_memory()

    def _ensure(self, n_events):
        if n_events <= self._cap and self._slots:
            return
        self._cap = int(n_events)
        dts = [torch.uint32] if self.wire == "p3" else [h.dtype for h in self.host]
        self._slots = []
        for _ in range(2):   # double buffering
            self._slots.append(dict(
                ev=[torch.empty((self._cap,), dtype=dt, device=self.device) for dt in dts],
                ev3=torch.empty((3 * self._cap if self.wire == "p3" else 0,), dtype=torch.uint8, device=self.device),
                out=torch.empty((self.group, self.B, self.H, self.W), dtype=torch.float32, device=self.device),
                ready=torch.cuda.Event(), done=torch.cuda.Event(), drained=torch.cuda.Event()))
        L = _lib.lib()
        nbytes = L.cmda_events_vg_workspace_bytes(self._cap, self.group, self.H, self.W, self.B,
                                                  _lib.VOXEL_MODES[self.mode])
        self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)

    @staticmethod
    def _plan_copies(starts, ends, members):
        """Merge the windows of a group into copy ranges (windows that touch or overlap in the store share one) and
        place every range on an _ALIGN boundary of the staging buffer.  Returns (ranges [(src_a, src_b, pos)],
        device start of every member, events needed)."""
        order = sorted(members, key=lambda s: (int(starts[s]), int(ends[s])))
        ranges, where = [], {}
        pos = 0
        for s in order:
            a, b = int(starts[s]), int(ends[s])
            if b <= a:
                where[s] = 0
                continue
            if ranges and a <= ranges[-1][1]:
                ra, rb, rp = ranges[-1]
                ranges[-1] = (ra, max(rb, b), rp)
            else:
                if ranges:
                    ra, rb, rp = ranges[-1]
                    pos = rp + (rb - ra + _ALIGN - 1) // _ALIGN * _ALIGN + _ALIGN
                ranges.append((a, b, pos))
            where[s] = ranges[-1][2] + (a - ranges[-1][0])
        need = 0
        if ranges:
            ra, rb, rp = ranges[-1]
            need = rp + (rb - ra + _ALIGN - 1) // _ALIGN * _ALIGN + _ALIGN
        return ranges, where, need

    def __call__(self, starts, finishes, clip_ranges=None, out=None, map_ids=None):
        """Inclusive windows ``[start, finish]`` -> pinned host tensor ``[S, B, H, W]``.
        ``map_ids[s]`` selects the rectify map of window ``s`` when several were given.
        Returns after the last device->host copy has completed."""
        L = _lib.lib()
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        ends = np.ascontiguousarray(finishes, dtype=np.int64) + 1
        S = int(starts.shape[0])
        if S and (starts.min() < 0 or ends.max() > self.n_total):
            raise IndexError("event window outside the store")
        mids = None if map_ids is None else np.ascontiguousarray(map_ids, dtype=np.int32)
        if mids is not None and S and (self.rmap is None or mids.min() < 0 or mids.max() >= self.rmap.shape[0]):
            raise IndexError("map id outside the rectify maps of the pipeline")
        if out is None:
            out = torch.empty((S, self.B, self.H, self.W), dtype=torch.float32).pin_memory()
        assert out.is_pinned() and out.shape == (S, self.B, self.H, self.W)
        groups = [list(range(g, min(g + self.group, S))) for g in range(0, S, self.group)]
        plans = [self._plan_copies(starts, ends, g) for g in groups]
        self._ensure(max((pl[2] for pl in plans), default=0))
        mode_id = _lib.VOXEL_MODES[self.mode]
        h2d = 0
        with torch.cuda.device(self.device):
            for gi, g in enumerate(groups):
                ranges, where, _ = plans[gi]
                slot = self._slots[gi % 2]
                # the slot's previous result must have left the device before it is overwritten
                self.copy_in.wait_event(slot["drained"])
                self.copy_in.wait_event(slot["done"])
                with torch.cuda.stream(self.copy_in):
                    for a, b, pos in ranges:
                        if self.wire == "p3":
                            slot["ev3"][3 * pos:3 * (pos + b - a)].copy_(self.host[0][3 * a:3 * b], non_blocking=True)
                        else:
                            for dev_arr, host_arr in zip(slot["ev"], self.host):
                                dev_arr[pos:pos + (b - a)].copy_(host_arr[a:b], non_blocking=True)
                        h2d += (b - a) * self.bytes_per_event
                    slot["ready"].record(self.copy_in)
                clips = np.array([default_clip_range(int(ends[s]) - 1, int(starts[s]))
                                  if (clip_ranges is None or clip_ranges[s] is None) else clip_ranges[s] for s in g],
                                 dtype=np.float32)
                hs = np.array([where[s] for s in g], dtype=np.int64)
                he = np.array([where[s] + max(int(ends[s] - starts[s]), 0) for s in g], dtype=np.int64)
                hm = None if mids is None else np.ascontiguousarray(mids[g[0]:g[0] + len(g)])
                self.compute.wait_event(slot["ready"])
                with torch.cuda.stream(self.compute):
                    ev = slot["ev"]
                    if self.wire == "p3":        # wire records -> P4 records, range by range, then the P4 path
                        for a, b, pos in ranges:
                            j_lo = int(np.searchsorted(self.h_sub_to_idx, a, side="right")) - 1
                            j_hi = int(np.searchsorted(self.h_sub_to_idx, b - 1, side="right")) - 1
                            _lib.check(L.cmda_unpack_p3_to_p4(_lib.ptr(slot["ev3"][3 * pos:]), _lib.ptr(self.d_sub_to_idx), j_lo, j_hi,
                                                              a, b, _lib.ptr(ev[0][pos:]), self.compute.cuda_stream),
                                       "cmda_unpack_p3_to_p4")
                    if self.wire in ("p4", "p3"):
                        src = np.ascontiguousarray(starts[g[0]:g[0] + len(g)])
                        _lib.check(L.cmda_events_vg_batch_p4(
                            _lib.ptr(ev[0]), _lib.ptr(self.d_ms_to_idx), _lib.host_ptr(self.h_ms_to_idx), len(self.h_ms_to_idx) - 1,
                            _lib.host_ptr(hs), _lib.host_ptr(he), _lib.host_ptr(src), len(g), _lib.ptr(self.rmap), _lib.host_ptr(hm),
                            self.H, self.W, self.B, _lib.host_ptr(clips), 1.0, 1, 1, _lib.ptr(slot["out"]), None, None,
                            _lib.ptr(self._ws), self._ws.numel(), mode_id, _lib.ptr(self.plans), self.compute.cuda_stream),
                            "cmda_events_vg_batch_p4")
                    else:
                        _lib.check(L.cmda_events_vg_batch_planned(
                            _lib.ptr(ev[0]), _lib.ptr(ev[1]), _lib.ptr(ev[2]), _lib.ptr(ev[3]), _lib.host_ptr(hs),
                            _lib.host_ptr(he), len(g), _lib.ptr(self.rmap), _lib.host_ptr(hm), self.H, self.W, self.B,
                            _lib.host_ptr(clips), 1.0, 1, 1, _lib.ptr(slot["out"]), None, None, _lib.ptr(self._ws),
                            self._ws.numel(), mode_id, _lib.ptr(self.plans), self.compute.cuda_stream),
                            "cmda_events_vg_batch_planned")
                    slot["done"].record(self.compute)
                self.copy_out.wait_event(slot["done"])
                with torch.cuda.stream(self.copy_out):
                    out[g[0]:g[0] + len(g)].copy_(slot["out"][:len(g)], non_blocking=True)
                    slot["drained"].record(self.copy_out)
            self.copy_out.synchronize()
        self.last_h2d_bytes = h2d
        return out

    def bytes_per_call(self, starts, finishes):
        """(host->device, device->host) bytes of one call: counted from the copies the call issues."""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        ends = np.ascontiguousarray(finishes, dtype=np.int64) + 1
        S = len(starts)
        groups = [list(range(g, min(g + self.group, S))) for g in range(0, S, self.group)]
        n = sum(b - a for g in groups for a, b, _ in self._plan_copies(starts, ends, g)[0])
        return self.bytes_per_event * n, 4 * S * self.B * self.H * self.W

    def close(self):
        """Drop the staging buffers and the cached scratch buffers of this pipeline's streams."""
        self._slots, self._ws, self._cap = [], None, 0
        _lib.release_workspaces()
