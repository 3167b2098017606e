"""ctypes binding of ``libcmda_b200.so`` (the C ABI declared in include/cmda_b200.h).

There is NO fallback: if the shared library is missing or a call fails this module
raises.  PyTorch is used for device memory and streams only; every pointer that crosses
the boundary is a plain address.
"""
from __future__ import annotations

import collections
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CMDA_B200_LIB points at another build of the same ABI (kernel-shape sweeps: tools/build_variant.sh)
LIB_PATH = os.environ.get("CMDA_B200_LIB") or os.path.join(_HERE, "libcmda_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

OK = 0
VOXEL_GLOBAL, VOXEL_TILED, VOXEL_AUTO, VOXEL_EXACT, VOXEL_FACTORED, VOXEL_BANDED, VOXEL_BANDED2 = 0, 1, 2, 3, 4, 5, 6
VOXEL_MODES = {"global": VOXEL_GLOBAL, "tiled": VOXEL_TILED, "auto": VOXEL_AUTO, "exact": VOXEL_EXACT,
               "factored": VOXEL_FACTORED, "banded": VOXEL_BANDED, "banded2": VOXEL_BANDED2}
VOXEL_MODE_NAMES = {v: k for k, v in VOXEL_MODES.items()}
DIRECTIONS = {"rightdown": 0, "rightup": 1, "leftdown": 2, "leftup": 3, "all": 4}

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/cmda_b200.h one to one
SIGNATURES = {
    "cmda_strerror": (ctypes.c_char_p, [_int]),
    "cmda_version": (_int, []),
    "cmda_last_cuda_error": (_int, []),
    "cmda_profiler_attach": (_int, [_vp, _int]),
    "cmda_profiler_detach": (_int, []),
    "cmda_event_create": (_vp, []),
    "cmda_event_destroy": (_int, [_vp]),
    "cmda_event_elapsed_ms": (_int, [_vp, _vp, _vp]),
    "cmda_searchsorted_right_u32": (_int, [_vp, _i64, _vp, _int, _vp, _vp]),
    "cmda_images_to_events_index": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _int, _vp, _vp, _vp]),
    "cmda_events_vg_workspace_bytes": (_sz, [_i64, _int, _int, _int, _int, _int]),
    "cmda_events_vg_resolved_mode": (_int, [_i64, _int, _int, _int, _int, _int]),
    "cmda_events_vg_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp, _f32, _int,
                                    _int, _vp, _vp, _vp, _vp, _sz, _int, _vp]),
    "cmda_events_vg_augmented_workspace_bytes": (_sz, [_i64, _int, _int, _int, _int, _int]),
    "cmda_events_vg_augmented_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp, _f32, _int,
                                              _vp, _int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _sz, _int, _vp,
                                              _vp]),
    "cmda_rectify_plan_bytes": (_sz, [_int, _int]),
    "cmda_rectify_plan_build": (_int, [_vp, _int, _int, _int, _vp, _vp]),
    "cmda_events_vg_batch_planned": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp, _f32, _int,
                                            _int, _vp, _vp, _vp, _vp, _sz, _int, _vp, _vp]),
    "cmda_pack_events_p4": (_int, [_vp, _vp, _vp, _vp, _i64, ctypes.c_uint32, _i64, _vp, _vp, _vp, _vp]),
    "cmda_unpack_p3_to_p4": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "cmda_events_vg_batch_p4": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _int, _vp, _vp, _int, _int, _int, _vp, _f32, _int,
                                       _int, _vp, _vp, _vp, _vp, _sz, _int, _vp, _vp]),
    "cmda_voxel_grid_f32": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _sz, _int, _vp]),
    "cmda_events_norm_workspace_bytes": (_sz, [_int]),
    "cmda_events_norm_batch": (_int, [_vp, _int, _i64, _vp, _f32, _int, _vp, _sz, _vp]),
    "cmda_remap_events": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp]),
    "cmda_image_workspace_bytes": (_sz, [_int, _int, _int, _int]),
    "cmda_logdiff_pair_u8": (_int, [_vp, _vp, _int, _int, _int, _vp, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "cmda_isr_shift_u8": (_int, [_vp, _int, _int, _int, _int, _int, _int, _vp, _f32, _f32, _vp, _vp, _sz, _vp]),
    "cmda_rgb_to_gray_u8": (_int, [_vp, _i64, _vp, _vp]),
    "cmda_resize_bilinear_workspace_bytes": (_sz, [_int, _int, _int, _int, _int, _int]),
    "cmda_resize_bilinear_u8": (_int, [_vp, _int, _int, _int, _int, _int, _int, _vp, _vp, _sz, _vp]),
    "cmda_u8_crop_to_centered_f32": (_int, [_vp, _int, _int, _int, _vp, _int, _int, _int, _vp, _vp]),
    "cmda_denorm_rgb_to_gray_u8": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class CmdaError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA translation unit for sm_100a into ``libcmda_b200.so`` (in tree)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise CmdaError("building libcmda_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded shared library; raises (no fallback) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise CmdaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(or `make -C cmda_b200/csrc`); there is no CPU fallback")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = lib().cmda_strerror(rc).decode()
        extra = f" (cudaError {lib().cmda_last_cuda_error()})" if rc == -2 else ""
        raise CmdaError(f"{what}: {msg}{extra}")


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise CmdaError(f"{name} must be a CUDA tensor: the cmda_b200 path has no CPU implementation")
    return t


def ptr(t) -> int | None:
    """Device address of a tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def host_ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(_vp)


def stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream on ``device`` (the raw-handle query: this sits on every call's path)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


class on_device:
    """``torch.cuda.device(dev)`` that costs nothing when ``dev`` is already the current device."""
    __slots__ = ("_ctx",)

    def __init__(self, device):
        idx = device.index
        self._ctx = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


_WORKSPACE_SLOTS = 8          # scratch buffers kept per process (least recently used beyond that are dropped)
_workspaces: "collections.OrderedDict" = collections.OrderedDict()


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """A scratch buffer per (device, stream); the library itself never allocates.  A buffer is allocated while its
    stream is torch's current stream, so the caching allocator orders its reuse after the work queued on that
    stream: dropping a buffer (the cache keeps the `_WORKSPACE_SLOTS` most recently used ones, so short-lived
    streams do not pin memory for ever) is safe without a synchronisation."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream_ptr(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)           # release the smaller buffer before the larger one is allocated
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    _workspaces.move_to_end(key)
    while len(_workspaces) > _WORKSPACE_SLOTS:
        _workspaces.popitem(last=False)
    return buf


def release_workspaces() -> None:
    """Drop every cached scratch buffer (pipelines call this on teardown)."""
    _workspaces.clear()
