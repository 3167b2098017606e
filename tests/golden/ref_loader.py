"""Load the reference's hot-path functions from ``/root/reference`` WITHOUT copying them.

Used only by ``make_golden.py`` (and by optional cross-check tests that skip
when ``/root/reference`` is absent, as it is on the GPU box).  The package
``mmseg`` cannot be imported here (mmcv / h5py / hdf5plugin are not installed),
so the individual ``FunctionDef`` nodes are parsed out of the source files with
``ast`` and executed in a namespace that only holds numpy / torch.  Nothing is
written into the repository by this module; the fixtures it helps produce are
outputs of the reference, not its source.
"""
from __future__ import annotations

import ast
import importlib.util
import math
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF_ROOT = os.environ.get("CMDA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mmseg", "datasets", "dsec.py"))


def _extract(path: str, names, namespace: dict, class_name: str | None = None) -> dict:
    src = open(path, "r", encoding="utf-8").read()
    tree = ast.parse(src)
    body = tree.body
    if class_name is not None:
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name == class_name:
                body = node.body
                break
        else:
            raise KeyError(class_name)
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise KeyError(f"{missing} not found in {path}")
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace


def pin_deterministic():
    """1 intra-op thread + deterministic algorithms: the only mode in which the
    reference's ``put_(accumulate=True)`` is bit-repeatable."""
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)


def dsec_functions() -> dict:
    """events_to_voxel_grid / tensor_normalize_to_range / events_norm
    (reference mmseg/datasets/dsec.py:26-121)."""
    ns = {"torch": torch, "np": np, "F": F, "random": random}
    return _extract(os.path.join(REF_ROOT, "mmseg", "datasets", "dsec.py"),
                    ["events_to_voxel_grid", "tensor_normalize_to_range", "events_norm"], ns)


class _DSECStub:
    """Minimal object the extracted ``get_events_vg`` method can be bound to."""

    def __init__(self, t, x, y, p, rectify_map, width, height, bins, clip_range=None):
        self.events_h5 = {"events/t": t, "events/x": x, "events/y": y, "events/p": p}
        self.rectify_map = rectify_map
        self.rectify_events = True
        self.events_width = width
        self.events_height = height
        self.events_bins = bins
        self.events_clip_range = clip_range


def get_events_vg(t, x, y, p, rectify_map, width, height, bins, finish, start, clip_range=None):
    """Run the reference's ``DSECDataset.get_events_vg`` (dsec.py:341-366) on numpy
    arrays standing in for the h5py datasets (h5py slicing == numpy slicing)."""
    ns = dsec_functions()
    _extract(os.path.join(REF_ROOT, "mmseg", "datasets", "dsec.py"), ["get_events_vg"], ns,
             class_name="DSECDataset")
    stub = _DSECStub(t, x, y, p, rectify_map, width, height, bins, clip_range)
    return ns["get_events_vg"](stub, finish, start)


def utils_module():
    """mmseg/datasets/utils.py only depends on numpy/torch/math -> import by path."""
    path = os.path.join(REF_ROOT, "mmseg", "datasets", "utils.py")
    spec = importlib.util.spec_from_file_location("_cmda_ref_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def image_change_functions(log_add=50, threshold=0.1, clip_range=0.8) -> dict:
    """get_image_change (create_cityscapes_image_change.py:9-35) with the module
    globals of lines 169-172 injected."""
    from PIL import Image
    ns = {"torch": torch, "np": np, "Image": Image, "log_add": log_add, "threshold": threshold,
          "clip_range": clip_range}
    return _extract(os.path.join(REF_ROOT, "create_cityscapes_image_change.py"),
                    ["tensor_normalize_to_range", "get_image_change"], ns)


class _FakeH5Dataset:
    def __init__(self, arr):
        self._a = np.asarray(arr)
        self.shape = self._a.shape

    def __getitem__(self, k):
        if isinstance(k, tuple) and len(k) == 0:
            return self._a[()]
        return self._a[k]

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __len__(self):
        return len(self._a)


def images_to_events_index(t, t_offset, ms_to_idx, timestamps, tmpdir: str):
    """Run create_images_to_events_index (create_dsec_dataset_txt.py:10-47) against
    a dict-backed fake ``h5py`` module; returns the int list it writes."""
    store = {"events/t": _FakeH5Dataset(t), "t_offset": _FakeH5Dataset(np.int64(t_offset)),
             "ms_to_idx": _FakeH5Dataset(ms_to_idx)}
    fake_h5py = types.ModuleType("h5py")
    fake_h5py.File = lambda path, mode="r": store
    ns = {"os": os, "math": math, "np": np, "h5py": fake_h5py, "tqdm": (lambda it, **kw: it)}
    _extract(os.path.join(REF_ROOT, "create_dsec_dataset_txt.py"), ["create_images_to_events_index"], ns)
    ts_path = os.path.join(tmpdir, "timestamps.txt")
    out_path = os.path.join(tmpdir, "images_to_events_index.txt")
    np.savetxt(ts_path, np.asarray(timestamps, dtype=np.int64), fmt="%d")
    ns["create_images_to_events_index"](ts_path, "unused.h5", out_path)
    return [int(v) for v in open(out_path).read().split()]
