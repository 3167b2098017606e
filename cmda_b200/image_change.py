"""Log-intensity-change pseudo-events: the reference's call signatures over K4 / K5.

Mirrors ``get_ic`` / ``get_image_change_from_pil`` (reference mmseg/datasets/utils.py:87-152)
and ``get_image_change`` (reference create_cityscapes_image_change.py:16-35).  The 256-entry
log table is evaluated here by numpy with the reference's own expression, so the values the
kernel subtracts are the values the reference subtracts.  Outputs live where the inputs
live (PIL / numpy / CPU tensor -> CPU tensor, CUDA tensor -> CUDA tensor) unless
``out_device`` says otherwise.
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import _lib
from .voxel import _cuda_device

__all__ = ["get_ic", "get_image_change_from_pil", "get_image_change", "isr_batch", "image_change_batch",
           "rgb_to_gray", "log_lut_val_range", "log_lut_log_add"]

# module-level parameters of create_cityscapes_image_change.py:169-172
image_change_range = 1
log_add = 50
threshold = 0.1
clip_range = 0.8


@functools.lru_cache(maxsize=64)
def _lut_val_range(v0, v1) -> np.ndarray:
    g = np.arange(256, dtype=np.float32)
    with np.errstate(all="ignore"):
        return np.ascontiguousarray(np.log(g / 255 * (v1 - v0) + v0), dtype=np.float32)    # utils.py:88-91


def log_lut_val_range(val_range) -> np.ndarray:
    return _lut_val_range(val_range[0], val_range[1])


@functools.lru_cache(maxsize=64)
def log_lut_log_add(add) -> np.ndarray:
    g = np.arange(256, dtype=np.float32)
    with np.errstate(all="ignore"):
        return np.ascontiguousarray(np.log(g + add), dtype=np.float32)   # create_cityscapes_image_change.py:17-20


def _to_u8_cuda(img, dev) -> torch.Tensor:
    if isinstance(img, torch.Tensor):
        assert img.dtype == torch.uint8, "uint8 image expected"
        return img.to(dev, non_blocking=True).contiguous()
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if not a.flags.writeable:          # e.g. np.asarray(PIL image): torch wants a writable buffer to wrap
        a = a.copy()
    return torch.from_numpy(a).to(dev, non_blocking=True)


def _home(img) -> torch.device:
    return img.device if isinstance(img, torch.Tensor) else torch.device("cpu")


def rgb_to_gray(rgb, *, out_device=None) -> torch.Tensor:
    """``PIL.Image.convert('L')`` of ``[..., 3]`` uint8 RGB (utils.py:126) on the device."""
    home = _home(rgb)
    dev = _cuda_device(home if home.type == "cuda" else None)
    src = _to_u8_cuda(rgb, dev)
    assert src.shape[-1] == 3
    out = torch.empty(src.shape[:-1], dtype=torch.uint8, device=dev)
    with _lib.on_device(dev):
        _lib.check(_lib.lib().cmda_rgb_to_gray_u8(_lib.ptr(src), out.numel(), _lib.ptr(out), _lib.stream_ptr(dev)),
                   "cmda_rgb_to_gray_u8")
    return out.to(out_device if out_device is not None else home)


def isr_batch(images, shift_pixel, val_range, _threshold, _clip_range, shift_direction="rightdown", *,
              out=None, out_device=None) -> torch.Tensor:
    """Shift-pair pseudo-events of ``[S, H, W]`` gray or ``[S, H, W, 3]`` RGB uint8 images
    -> ``[S, 1, H, W]`` float32 (``get_image_change_from_pil`` batched)."""
    home = _home(images)
    dev = _cuda_device(home if home.type == "cuda" else None)
    src = _to_u8_cuda(images, dev)
    channels = 3 if (src.ndim == 4 and src.shape[-1] == 3) else 1
    assert src.ndim == 3 + (channels == 3)
    S, H, W = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
    if shift_direction not in _lib.DIRECTIONS:
        raise AssertionError(shift_direction)                      # utils.py:142 / 147
    lut = log_lut_val_range(tuple(val_range))
    span = np.log(val_range[1]) - np.log(val_range[0])             # utils.py:93-94, float64
    thr = np.float32(span * _threshold)                            # compared in float32
    clip = np.float32(span * _clip_range)
    L = _lib.lib()
    if out is None:
        out = torch.empty((S, 1, H, W), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.is_contiguous() and out.numel() == S * H * W
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_image_workspace_bytes(S, H, W, channels))
        _lib.check(L.cmda_isr_shift_u8(_lib.ptr(src), channels, S, H, W, int(shift_pixel),
                                       _lib.DIRECTIONS[shift_direction], _lib.host_ptr(lut), float(thr), float(clip),
                                       _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                   "cmda_isr_shift_u8")
    return out.to(out_device if out_device is not None else home)


def get_image_change_from_pil(pil_image, width, height, data_type=None, shift_pixel=4, val_range=None,
                              _threshold=None, _clip_range=None, auto_threshold=None, shift_direction='rightdown',
                              *, out_device=None):
    """Drop-in for ``get_image_change_from_pil`` (reference utils.py:108-152) -> ``[1, H, W]``.

    ``pil_image`` may be a PIL image ('RGB' or 'L'), a uint8 array or a uint8 tensor
    (``[H, W, 3]`` or ``[H, W]``).  ``width`` / ``height`` / ``data_type`` are accepted for
    signature compatibility; like the reference's slicing, the image's own size rules.
    """
    if auto_threshold is not None:
        raise ValueError('auto_threshold function not implement！')   # utils.py:124-125
    if hasattr(pil_image, "convert"):
        if pil_image.mode not in ("RGB", "L"):
            pil_image = pil_image.convert("RGB")
        pil_image = np.asarray(pil_image)
    img = pil_image[None]
    return isr_batch(img, shift_pixel, val_range, _threshold, _clip_range, shift_direction, out_device=out_device)[0]


def _three_floats(v, dev):
    """The 3 channel constants of ``denorm`` as ``(keep_alive, pointer)``.  The reference keeps ``means`` / ``stds``
    as ``[B, 3, 1, 1]`` CUDA tensors (dacs_transforms.py:38-49) and uses row 0 for every sample
    (``.numpy()[0]`` of the broadcast, dacs.py:730-731): a float32 CUDA tensor is handed to the kernel as it is
    (no read-back, which would synchronise the device once per train step); anything else goes through the host."""
    if isinstance(v, torch.Tensor) and v.is_cuda and v.device == dev:
        t = v.detach().reshape(-1)[:3].to(torch.float32).contiguous()       # views only for the reference's tensors
        assert t.numel() == 3
        return t, t.data_ptr()
    a = np.ascontiguousarray(torch.as_tensor(v).detach().float().cpu().reshape(-1)[:3].numpy(), dtype=np.float32)
    assert a.size == 3
    return a, _lib.host_ptr(a)


def denorm_to_gray(img, means, stds, *, return_rgb=False):
    """``clamp(denorm(img, means, stds), 0, 1) * 255 -> uint8 -> PIL 'L'`` on the device
    (reference dacs.py:730-733 with dacs_transforms.py:52-53).  ``img`` is a CUDA float32
    ``[S, 3, H, W]`` batch, ``means`` / ``stds`` anything holding 3 numbers (the reference's
    ``[1, 3, 1, 1]`` tensors are fine).  Returns uint8 ``[S, H, W]`` (and ``[S, H, W, 3]``)."""
    img = _lib.require_cuda(img, "img")
    assert img.ndim == 4 and img.shape[1] == 3 and img.dtype == torch.float32
    img = img.contiguous()
    dev = img.device
    (m_keep, m_ptr), (sd_keep, sd_ptr) = _three_floats(means, dev), _three_floats(stds, dev)
    if isinstance(m_keep, torch.Tensor) != isinstance(sd_keep, torch.Tensor):     # both on the same side
        m_keep, m_ptr = _three_floats(torch.as_tensor(means).cpu(), dev)
        sd_keep, sd_ptr = _three_floats(torch.as_tensor(stds).cpu(), dev)
    S, _, H, W = (int(v) for v in img.shape)
    gray = torch.empty((S, H, W), dtype=torch.uint8, device=dev)
    rgb = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev) if return_rgb else None
    with _lib.on_device(dev):
        _lib.check(_lib.lib().cmda_denorm_rgb_to_gray_u8(_lib.ptr(img), S, H, W, m_ptr, sd_ptr,
                                                         _lib.ptr(gray), _lib.ptr(rgb), _lib.stream_ptr(dev)),
                   "cmda_denorm_rgb_to_gray_u8")
    return (gray, rgb) if return_rgb else gray


def mixed_image_isr(mixed_img, means, stds, shift_direction='rightdown', shift_pixel=4, val_range=None, _threshold=None,
                    _clip_range=None):
    """The mixed-image ISR of the DACS train step (reference dacs.py:729-744) without leaving the
    GPU: ``[S, 3, H, W]`` normalised float image batch -> ``[S, 3, H, W]`` ISR
    (``get_image_change_from_pil(...).repeat(3, 1, 1)[None]`` per sample, stacked)."""
    gray = denorm_to_gray(mixed_img, means, stds)
    isr = isr_batch(gray, shift_pixel, val_range, _threshold, _clip_range, shift_direction)
    return isr.expand(-1, 3, -1, -1).contiguous()


def pil_resize_bilinear(images, size, *, out_device=None) -> torch.Tensor:
    """``PIL.Image.resize(size, Image.BILINEAR)`` of uint8 ``[S, H, W]`` ('L') or ``[S, H, W, 3]`` ('RGB') images,
    bit for bit, on the device (reference cityscapes_ic.py:152-153, 175-176).  ``size`` is ``(width, height)`` as in
    PIL.  Returns uint8 ``[S, h, w(, 3)]``."""
    home = _home(images)
    dev = _cuda_device(home if home.type == "cuda" else None)
    src = _to_u8_cuda(images, dev)
    channels = 3 if (src.ndim == 4 and src.shape[-1] == 3) else 1
    assert src.ndim == 3 + (channels == 3)
    S, H, W = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
    ow, oh = int(size[0]), int(size[1])
    out = torch.empty((S, oh, ow, 3) if channels == 3 else (S, oh, ow), dtype=torch.uint8, device=dev)
    L = _lib.lib()
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_resize_bilinear_workspace_bytes(S, H, W, channels, oh, ow))
        _lib.check(L.cmda_resize_bilinear_u8(_lib.ptr(src), channels, S, H, W, oh, ow, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr(dev)), "cmda_resize_bilinear_u8")
    return out.to(out_device if out_device is not None else home)


def u8_crop_to_centered(gray, crop_xy, crop_size, flips=None, repeat=3, *, out_device=None) -> torch.Tensor:
    """``crop -> HorizontalFlip -> float32 -> (v / 255.0 - 0.5) / 0.5 -> repeat(repeat, 1, 1)`` of uint8 gray images
    ``[S, H, W]`` (reference cityscapes_ic.py:177-183, 207-209).  ``crop_size`` is ``(w, h)``; returns float32
    ``[S, repeat, h, w]``."""
    home = _home(gray)
    dev = _cuda_device(home if home.type == "cuda" else None)
    src = _to_u8_cuda(gray, dev)
    assert src.ndim == 3
    S, H, W = (int(v) for v in src.shape)
    cw, ch = int(crop_size[0]), int(crop_size[1])
    aug = np.zeros((S, 3), dtype=np.int32)
    xy = np.asarray(crop_xy, dtype=np.int64).reshape(-1, 2)
    aug[:, 0:2] = xy if xy.shape[0] == S else np.broadcast_to(xy, (S, 2))
    if flips is not None:
        aug[:, 2] = np.asarray(flips, dtype=np.int32).reshape(-1)
    if S and (aug[:, 0].min() < 0 or aug[:, 1].min() < 0 or (aug[:, 0] + cw).max() > W or (aug[:, 1] + ch).max() > H):
        raise IndexError("crop outside the image")
    out = torch.empty((S, int(repeat), ch, cw), dtype=torch.float32, device=dev)
    with _lib.on_device(dev):
        _lib.check(_lib.lib().cmda_u8_crop_to_centered_f32(_lib.ptr(src), S, H, W, _lib.host_ptr(aug), cw, ch, int(repeat),
                                                           _lib.ptr(out), _lib.stream_ptr(dev)), "cmda_u8_crop_to_centered_f32")
    return out.to(out_device if out_device is not None else home)


def source_img_time_res(now, front, resize_size=(1024, 512), crop_xy=(0, 0), crop_size=(512, 512), flips=None, repeat=3):
    """The Cityscapes source branch's ``img_time_res`` straight from a frame pair, on the device: the offline
    ``get_image_change`` PNG (create_cityscapes_image_change.py:16-35) -> ``resize(BILINEAR)`` -> crop -> flip ->
    ``(x / 255 - 0.5) / 0.5`` -> ``repeat(3, 1, 1)`` (cityscapes_ic.py:175-183, 207-209).  ``now`` / ``front``:
    uint8 gray ``[S, H, W]``."""
    u8 = image_change_batch(now, front, want_f32=False, want_u8=True, out_device=_cuda_device(None))
    small = pil_resize_bilinear(u8, resize_size)
    return u8_crop_to_centered(small, crop_xy, crop_size, flips, repeat)


def get_ic(image_front, image_now, val_range, threshold, clip_range, *, out_device=None):
    """Drop-in for ``get_ic`` (reference utils.py:87-105) on two uint8 gray images ->
    ``[1, H, W]`` float32.  Same arithmetic as K4 with the val_range table."""
    home = _home(image_now)
    dev = _cuda_device(home if home.type == "cuda" else None)
    now, front = _to_u8_cuda(image_now, dev), _to_u8_cuda(image_front, dev)
    span = np.log(val_range[1]) - np.log(val_range[0])
    out, _ = _pair(now[None], front[None], log_lut_val_range(tuple(val_range)), np.float32(span * threshold),
                   np.float32(span * clip_range), dev, want_f32=True, want_u8=False)
    return out[0][None].to(out_device if out_device is not None else home)


def _pair(now, front, lut, thr, clip, dev, want_f32, want_u8):
    assert now.shape == front.shape and now.ndim == 3
    S, H, W = (int(v) for v in now.shape)
    L = _lib.lib()
    out_f = torch.empty((S, H, W), dtype=torch.float32, device=dev) if want_f32 else None
    out_u = torch.empty((S, H, W), dtype=torch.uint8, device=dev) if want_u8 else None
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, L.cmda_image_workspace_bytes(S, H, W, 1))
        _lib.check(L.cmda_logdiff_pair_u8(_lib.ptr(now), _lib.ptr(front), S, H, W, _lib.host_ptr(lut), float(thr),
                                          float(clip), _lib.ptr(out_f), _lib.ptr(out_u), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr(dev)), "cmda_logdiff_pair_u8")
    return out_f, out_u


def image_change_batch(now, front, *, log_add=None, threshold=None, clip_range=None, want_f32=False, want_u8=True,
                       out_device=None):
    """Frame-pair pseudo-events of ``[S, H, W]`` uint8 gray stacks (``get_image_change``
    batched).  Returns the uint8 'L' payload and/or the float32 image in [-1, 1]."""
    g = globals()
    la = g["log_add"] if log_add is None else log_add
    th = g["threshold"] if threshold is None else threshold
    cr = g["clip_range"] if clip_range is None else clip_range
    home = _home(now)
    dev = _cuda_device(home if home.type == "cuda" else None)
    a, b = _to_u8_cuda(now, dev), _to_u8_cuda(front, dev)
    out_f, out_u = _pair(a, b, log_lut_log_add(la), np.float32(th), np.float32(cr), dev, want_f32, want_u8)
    tgt = out_device if out_device is not None else home
    res = tuple(o.to(tgt) for o in (out_f, out_u) if o is not None)
    return res[0] if len(res) == 1 else res


def get_image_change(image_now, image_front):
    """Drop-in for ``get_image_change`` (reference create_cityscapes_image_change.py:16-35):
    two 'L' images -> PIL 'L' image, using this module's ``log_add`` / ``threshold`` /
    ``clip_range`` globals exactly like the reference script uses its own."""
    from PIL import Image
    now = np.asarray(image_now, dtype=np.uint8)
    front = np.asarray(image_front, dtype=np.uint8)
    u8 = image_change_batch(now[None], front[None], want_u8=True)[0]
    return Image.fromarray(u8.cpu().numpy(), mode='L')
