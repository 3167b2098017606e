"""The events branch of ``DSECDataset.__getitem__`` over the CUDA path.

Mirrors reference mmseg/datasets/dsec.py:286-320 and 341-366 with the same attribute and
argument names.  File handling (events.h5 through h5py/hdf5plugin, rectify_map.h5,
images_to_events_index.txt) is I/O and stays with the caller: a ``DSECEvents`` is built
from arrays that are already in memory and keeps them resident on the GPU, instead of
re-opening three files per sample as the reference does (dsec.py:287-293).

CUDA cannot run in forked DataLoader workers, so this object is meant to be used from the
process that owns the GPU; the returned dict entry has the
reference's key, shape, dtype and value range.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn.functional as F

from .slicer import images_to_events_index, window_bounds
from .voxel import EventStore, events_vg_augmented_batch, events_vg_batch

__all__ = ["DSECEvents", "DSECDataset"]


class DSECEvents:
    def __init__(self, t, x, y, p, rectify_map, images_to_events_index, events_num=-1, events_bins=5,
                 events_clip_range=None, crop_size=(400, 400), after_crop_resize_size=(512, 512),
                 image_change_range=1, outputs={'events_vg', 'image'}, output_num=1, events_bins_5_avg_1=False,
                 enforce_3_channels=True, device=None, mode="auto", fused_augment=True, isr_parms='',
                 shift_type='rightdown', isr_type='real_time'):
        self.events_num = events_num
        self.events_bins = events_bins
        self.events_bins_5_avg_1 = events_bins_5_avg_1
        if self.events_bins_5_avg_1:                           # dsec.py:145-148
            assert events_bins == 1
            self.events_bins = 5
        self.events_clip_range = events_clip_range
        self.outputs = outputs
        # (H, W) --> (W, H) unless labels are requested -- dsec.py:150-152
        self.crop_size = (crop_size[1], crop_size[0]) if 'label' not in outputs else crop_size
        self.after_crop_resize_size = (after_crop_resize_size[1], after_crop_resize_size[0]) \
            if 'label' not in outputs else after_crop_resize_size
        self.image_change_range = image_change_range
        assert self.image_change_range in {1, 2}               # dsec.py:154
        self.output_num = output_num
        self.events_height = 480                               # dsec.py:160
        self.events_width = 640                                # dsec.py:161
        self.rectify_events = True                             # dsec.py:167
        self.enforce_3_channels = enforce_3_channels
        self.isr_type = isr_type
        assert self.isr_type in {'raw', 'denoised', 'real_time'}           # dsec.py:175
        self.image_change_parms = {'val_range': (1, 10 ** 2), '_threshold': 0.04, '_clip_range': 0.2, 'shift_pixel': 3}
        if isr_parms != '':                                                # dsec.py:178-181
            assert isinstance(isr_parms, dict)
            self.image_change_parms = isr_parms
        self.shift_type = shift_type
        assert self.shift_type in {'all', 'random', 'rightdown'}           # dsec.py:183
        self.images_to_events_index = [int(v) for v in images_to_events_index]
        self.mode = mode
        self.fused_augment = fused_augment     # crop / flip / resize / repeat inside the normaliser kernel (one launch less,
                                               # no normalised full grid); False keeps them as torch ops on the grid
        self.store = EventStore(t, x, y, p, rectify_map if self.rectify_events else None,
                                height=self.events_height, width=self.events_width, device=device)

    @classmethod
    def from_timestamps(cls, t, x, y, p, rectify_map, ms_to_idx, t_offset, images_timestamps, **kwargs):
        """Build the object from what a DSEC sequence holds on disk (events.h5: ``events/{t,x,y,p}``,
        ``ms_to_idx``, ``t_offset``; images/timestamps.txt) instead of from a precomputed
        ``images_to_events_index.txt``: the per-image event index of create_dsec_dataset_txt.py:10-47 is computed
        on the device from the resident ``t`` array (K1), once, instead of being read back with ``np.loadtxt`` for
        every sample (dsec.py:292-293).  Same exceptions as the reference script (``ValueError('range error!')``)."""
        obj = cls(t, x, y, p, rectify_map, [], **kwargs)
        obj.images_to_events_index = images_to_events_index(obj.store.t, t_offset, ms_to_idx, images_timestamps,
                                                            device=obj.store.device)
        return obj

    @classmethod
    def from_cache(cls, path, images_to_events_index=None, device=None, **kwargs):
        """Build the object from a decoded-sequence cache directory (``cmda_b200.store_io``: events.h5 decoded once,
        instead of being re-opened and blosc-decoded per sample as in dsec.py:287, 342-345).  The arrays are streamed
        from the memory-mapped files to the device; the per-image event index comes from the cached image
        timestamps (K1 on the device) unless ``images_to_events_index`` is given."""
        from . import store_io
        from .voxel import _cuda_device
        seq = store_io.load_sequence(path)
        dev = _cuda_device(device)
        t, x, y, p = (store_io.upload(seq[k], dev) for k in ("t", "x", "y", "p"))
        rmap = store_io.upload(seq["rectify_map"], dev)
        if images_to_events_index is not None:
            return cls(t, x, y, p, rmap, images_to_events_index, device=dev, **kwargs)
        if "images_timestamps" not in seq:
            raise ValueError(f"{path}: no image timestamps in the cache and no images_to_events_index given")
        return cls.from_timestamps(t, x, y, p, rmap, np.asarray(seq["ms_to_idx"]), seq["t_offset"],
                                   np.asarray(seq["images_timestamps"]), device=dev, **kwargs)

    # ---- dsec.py:341-366 -------------------------------------------------------------
    def _clip_for(self, finish, start):
        if self.events_clip_range is not None:                 # dsec.py:359-360
            return random.uniform(self.events_clip_range[0], self.events_clip_range[1])
        return (finish - start) / 500000 * 1.5                 # dsec.py:362

    def get_events_vg(self, events_finish_index, events_start_index):
        """``[events_bins, 480, 640]`` float32 in [-1, 1] on the GPU.  An empty slice raises ``IndexError`` as the
        reference's ``events_t[0]`` does (dsec.py:347); the batched entry point documents its own convention."""
        if events_start_index > events_finish_index:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        clip = self._clip_for(events_finish_index, events_start_index)
        return events_vg_batch(self.store, [events_start_index], [events_finish_index], self.events_bins, [clip],
                               mode=self.mode)[0]

    # ---- dsec.py:228-262 -------------------------------------------------------------
    def warp_img_self_res(self, warp_image, crop_xy=None, flip_flag=False):
        """The ``'warp_img_self_res'`` entry of ``__getitem__`` for ``isr_type='real_time'`` (dsec.py:252-262): the
        warp image (uint8 RGB ``[H, W, 3]``; array, tensor or PIL) is cropped at ``crop_xy``, flipped and resized with
        PIL's BILINEAR in train mode (dsec.py:229-233), then ``get_image_change_from_pil`` with the shift direction of
        ``shift_type`` (``'random'``: ``direct[x % 2][y % 2]``, dsec.py:253-255) and ``repeat(3, 1, 1)``.  Everything
        after the upload runs on the device; returns a CUDA float32 ``[3 or 1, h, w]`` tensor.  The ``'raw'`` /
        ``'denoised'`` types read precomputed PNGs (file I/O) and stay with the caller."""
        from .image_change import _to_u8_cuda, get_image_change_from_pil, pil_resize_bilinear
        if self.isr_type != 'real_time':
            raise NotImplementedError("isr_type 'raw' / 'denoised' read precomputed images (dsec.py:237-250)")
        if hasattr(warp_image, "convert"):
            warp_image = np.asarray(warp_image.convert('RGB'))
        img = _to_u8_cuda(warp_image, self.store.device)
        assert img.ndim == 3 and img.shape[2] == 3
        x, y = crop_xy if crop_xy is not None else (None, None)
        if 'label' not in self.outputs:                                    # train-time augmentation, dsec.py:229-233
            assert crop_xy is not None
            img = img[y: y + self.crop_size[1], x: x + self.crop_size[0]]
            if flip_flag:
                img = img.flip(1)
            img = pil_resize_bilinear(img.contiguous()[None], self.after_crop_resize_size)[0]
        if self.shift_type == 'random':                                    # dsec.py:253-255
            direct = [['leftdown', 'leftup'], ['rightdown', 'rightup']]
            this_shift_direction = direct[x % 2][y % 2]
        else:
            this_shift_direction = self.shift_type
        res = get_image_change_from_pil(img, width=int(img.shape[1]), height=int(img.shape[0]),
                                        shift_direction=this_shift_direction, **self.image_change_parms)
        if self.enforce_3_channels and res.shape[0] == 1:                  # dsec.py:259-260
            res = res.repeat(3, 1, 1)
        return res

    # ---- dsec.py:286-320 -------------------------------------------------------------
    def events_vg_for_image(self, now_image_index, crop_xy=None, flip_flag=False):
        """The ``'events_vg'`` entry of ``__getitem__`` for image ``now_image_index``.
        ``crop_xy``/``flip_flag`` are the augmentation draws of dsec.py:206-210 (made by the
        caller so that image, ISR and events share them).  Returns ``None`` where the
        reference does (start > finish, dsec.py:301-302).

        ``output_num > 1``: the reference's own post-voxel statements only work for one window -- on the 4-D
        ``[output_num, bins, H, W]`` tensor ``events_vg[:, y:y+ch, x:x+cw]`` (dsec.py:310) slices the BINS and ROW
        axes and ``F.interpolate(events_vg[None], size=(h, w))`` (dsec.py:314) raises for a 5-D input, and test mode's
        ``[:, :440, :]`` (dsec.py:317) cuts the bins axis.  This method INTENTIONALLY differs there: crop / flip /
        resize / ``[:440]`` always act on the last two (spatial) axes of every window, the repeat on the channel
        axis; for ``output_num == 1`` (the only value the reference's configs use) it is the reference's result."""
        bounds = []
        for i in range(self.output_num):                       # dsec.py:295-302
            b = window_bounds(self.images_to_events_index, now_image_index, self.image_change_range,
                              self.events_num, i)
            if b is None:
                return None
            bounds.append(b)
        clips = [self._clip_for(f, s) for s, f in bounds]
        if self.output_num == 1 and self.fused_augment:
            # dsec.py:304-319 fused into the normaliser (one window: the squeeze of :306-307 applies)
            train = 'label' not in self.outputs
            crop_size = self.crop_size if train else (self.events_width, 440)
            out_size = self.after_crop_resize_size if train else crop_size
            x, y = (crop_xy if crop_xy is not None else (0, 0)) if train else (0, 0)
            return events_vg_augmented_batch(self.store, [bounds[0][0]], [bounds[0][1]], self.events_bins, clips,
                                             crop_xy=[(x, y)], crop_size=crop_size, out_size=out_size,
                                             flips=[int(bool(flip_flag) and train)], avg_bins=self.events_bins_5_avg_1,
                                             repeat=3 if self.enforce_3_channels else 1, mode=self.mode)[0]
        vg = events_vg_batch(self.store, [s for s, _ in bounds], [f for _, f in bounds], self.events_bins, clips,
                             mode=self.mode)
        events_vg = vg.flip(0)                                 # events_vg[output_num - 1 - i] = window i, :303
        if self.events_bins_5_avg_1:
            events_vg = torch.mean(events_vg, dim=1, keepdim=True)   # dsec.py:304-305
        if self.output_num == 1:
            events_vg = events_vg[0]                           # dsec.py:306-307
        if 'label' not in self.outputs:                        # train-time augmentation, dsec.py:309-315
            x, y = crop_xy if crop_xy is not None else (0, 0)
            events_vg = events_vg[..., y: y + self.crop_size[1], x: x + self.crop_size[0]]
            if flip_flag:
                events_vg = events_vg.flip(-1)
            hw = (self.after_crop_resize_size[1], self.after_crop_resize_size[0])
            lead = events_vg.shape[:-2]
            events_vg = F.interpolate(events_vg.reshape(1, -1, *events_vg.shape[-2:]), size=hw, mode='bilinear',
                                      align_corners=False)[0].reshape(*lead, *hw)
        else:                                                  # test mode, dsec.py:316-317
            events_vg = events_vg[..., :440, :]
        if self.enforce_3_channels:                            # dsec.py:318-319
            reps = [1] * events_vg.ndim
            reps[-3] = 3
            events_vg = events_vg.repeat(*reps)
        return events_vg


class DSECDataset:
    """``DSECDataset.__getitem__`` (reference mmseg/datasets/dsec.py:124, 189-339) over the CUDA path: the dict a sample
    is, with the reference's keys, shapes, dtypes and value ranges for the entries this path produces --
    ``'events_vg'`` (dsec.py:286-320), ``'warp_img_self_res'`` (dsec.py:236-262, ``isr_type='real_time'``),
    ``'warp_image'`` (dsec.py:222-234, the crop / flip / BILINEAR resize / ToTensor / Normalize chain), ``'path'`` and
    ``'img_metas'`` (dsec.py:322-337, a plain dict: mmcv's DataContainer belongs to the caller).  The augmentation
    draws (flip, crop origin) are made here in the reference's order (dsec.py:204-209: ``random.random()``, then two
    ``random.randint``), so a seeded run picks the same crops.

    Files stay with the caller (image / label decoding is out of scope): one
    ``DSECEvents`` per sequence holds the events, ``samples`` lists ``(sequence_key, now_image_index)`` -- what a line
    of ``night_dataset_warp.txt`` encodes (dsec.py:199-201, 211) -- and ``warp_image_loader(sequence_key, index)``
    returns the warp image as a uint8 RGB ``[440..480, 640, 3]`` array (dsec.py:223-224) when an output needs it.
    Entries that need files this class is not given (``'image'``, ``'label'``, ``'19classes'``) raise ``KeyError``.
    Registers itself under mmseg's ``DATASETS`` registry as ``DSECDatasetB200`` where mmseg is importable.
    Everything returned lives on the sequence's CUDA device: use it from the process that owns the GPU."""

    CLASSES = ('road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'traffic light', 'traffic sign', 'vegetation',
               'terrain', 'sky', 'person', 'rider', 'car', 'truck', 'bus', 'train', 'motorcycle', 'bicycle')
    _HANDLED = {'events_vg', 'warp_image', 'warp_img_self_res', 'path', 'img_metas'}

    def __init__(self, sequences, samples, outputs={'events_vg', 'warp_image', 'warp_img_self_res'}, warp_image_loader=None):
        self.sequences = dict(sequences)
        self.samples = [(k, int(i)) for k, i in samples]
        self.outputs = set(outputs)
        missing = self.outputs - self._HANDLED
        if missing:
            raise KeyError(f"outputs {sorted(missing)} read files this adapter is not given (image / label I/O is the caller's)")
        self.warp_image_loader = warp_image_loader
        if ({'warp_image', 'warp_img_self_res'} & self.outputs) and warp_image_loader is None:
            raise ValueError("'warp_image' / 'warp_img_self_res' need a warp_image_loader")
        self.mean_std = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])              # dsec.py:162

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        from .image_change import _to_u8_cuda, pil_resize_bilinear
        key, now_image_index = self.samples[idx]
        ev = self.sequences[key]
        train = 'label' not in ev.outputs
        output = dict()
        flip_flag, x, y = False, None, None
        if train:                                                                    # dsec.py:204-209
            flip_flag = True if random.random() < 0.5 else False
            x = random.randint(0, 640 - ev.crop_size[0])
            y = random.randint(0, 480 - ev.crop_size[1])
        if 'path' in self.outputs:
            output['path'] = f"{key}/{now_image_index:06d}.png"
        warp = None
        if {'warp_image', 'warp_img_self_res'} & self.outputs:
            warp = _to_u8_cuda(np.asarray(self.warp_image_loader(key, now_image_index)), ev.store.device)
        if 'warp_image' in self.outputs:                                             # dsec.py:222-234
            img = warp
            if train:
                img = img[y: y + ev.crop_size[1], x: x + ev.crop_size[0]]
                if flip_flag:
                    img = img.flip(1)
                img = pil_resize_bilinear(img.contiguous()[None], ev.after_crop_resize_size)[0]
            chw = img.permute(2, 0, 1).to(torch.float32).div(255)                    # ToTensor
            mean = torch.tensor(self.mean_std[0], device=chw.device).view(3, 1, 1)
            std = torch.tensor(self.mean_std[1], device=chw.device).view(3, 1, 1)
            chw = (chw - mean) / std                                                 # Normalize
            output['warp_image'] = chw if train else chw[:, :440]
        if 'warp_img_self_res' in self.outputs:                                      # dsec.py:236-262
            output['warp_img_self_res'] = ev.warp_img_self_res(warp, crop_xy=(x, y) if train else None, flip_flag=flip_flag)
        if 'events_vg' in self.outputs:                                              # dsec.py:286-320
            events_vg = ev.events_vg_for_image(now_image_index, crop_xy=(x, y) if train else None, flip_flag=flip_flag)
            if events_vg is None:
                return None                                                          # dsec.py:301-302
            output['events_vg'] = events_vg
        if 'img_metas' in self.outputs:                                              # dsec.py:322-337
            output['img_metas'] = {
                'img_norm_cfg': {'mean': [123.675, 116.28, 103.53], 'std': [58.395, 57.12, 57.375], 'to_rgb': True},
                'img_shape': (440, 640), 'pad_shape': (440, 640), 'ori_shape': (440, 640),
                'ori_filename': f"{key}_{now_image_index:06d}.png", 'flip': False}
        return output


try:        # the reference's plugin mechanism (dsec.py:124): only where mmseg / mmcv exist
    from mmseg.datasets.builder import DATASETS as _DATASETS
    _DATASETS.register_module(name="DSECDatasetB200", module=DSECDataset)
except Exception:       # noqa: BLE001 -- mmseg is not part of this image
    pass
