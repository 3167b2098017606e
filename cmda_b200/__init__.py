"""cmda_b200 -- B200-native event-representation path of XiaRho/CMDA.

Only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI, built into
``libcmda_b200.so``) and thin host-side mirrors of the reference's call signatures.
"""
from ._lib import CmdaError, build, lib  # noqa: F401
from .voxel import (EventStore, default_clip_range, events_norm, events_to_voxel_grid, events_vg_augmented_batch,  # noqa: F401
                    events_vg_batch, remap_events)
from .image_change import (denorm_to_gray, get_ic, get_image_change, get_image_change_from_pil,  # noqa: F401
                           image_change_batch, isr_batch, mixed_image_isr, pil_resize_bilinear, rgb_to_gray,
                           source_img_time_res, u8_crop_to_centered)
from .slicer import images_to_events_index, searchsorted_right, window_bounds, write_index_txt  # noqa: F401
from .dsec import DSECDataset, DSECEvents  # noqa: F401
from .packed import PackedEventStore, pack_p3, pack_p4, unpack_p3, unpack_p4  # noqa: F401
from . import packed, sharding, store_io, synth  # noqa: F401

__version__ = "0.1.0"
