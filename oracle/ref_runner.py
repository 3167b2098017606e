"""Execute the reference's own functions from ``oracle/_ref`` (staged by ``oracle/ref_stage.py``).

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by ``cmda_b200``.  Used by ``bench.py --impl reference`` and
by the ``cpu_baseline`` legs: the reference's CPU implementation of the path, unmodified, on the host cores.

The package ``mmseg`` cannot be imported (mmcv / h5py / hdf5plugin are not in the image), so the individual
``FunctionDef`` nodes are parsed out of the staged files with ``ast`` and executed in a namespace that holds
numpy / torch only; ``DSECDataset.get_events_vg`` is bound to a stub whose ``events_h5`` is a dict of numpy
arrays (h5py slicing == numpy slicing).  Same technique as ``tests/golden/ref_loader.py`` (which reads
``/root/reference`` directly to generate the golden fixtures).

Throughput protocol (``WorkerPool``): the reference runs this path inside DataLoader worker PROCESSES, one
sample per worker at a time (builder.py:151-163), so the baseline uses one spawned process per host core
(at most one per work item), each owning its share of the step's items; a step = every worker runs its items
once; the time of a step is the wall time between "go" and the last worker's answer.
"""
from __future__ import annotations

import ast
import importlib.util
import multiprocessing as mp
import os
import random
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "mmseg", "datasets", "dsec.py"))


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _extract(path, names, namespace, class_name=None):
    tree = ast.parse(open(path, "r", encoding="utf-8").read())
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise KeyError(f"{missing} not found in {path}")
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), namespace)
    return namespace


class _DSECStub:
    def __init__(self, t, x, y, p, rectify_map, width, height, bins):
        self.events_h5 = {"events/t": t, "events/x": x, "events/y": y, "events/p": p}
        self.rectify_map = rectify_map
        self.rectify_events = True
        self.events_width, self.events_height, self.events_bins = width, height, bins
        self.events_clip_range = None


class Reference:
    """The reference's hot-path functions, loaded once per process."""

    def __init__(self):
        import numpy as np
        import torch
        import torch.nn.functional as F
        from PIL import Image
        if not available():
            raise FileNotFoundError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
        dsec = os.path.join(REF_DST, "mmseg", "datasets", "dsec.py")
        self.ns = {"torch": torch, "np": np, "F": F, "random": random}
        _extract(dsec, ["events_to_voxel_grid", "tensor_normalize_to_range", "events_norm"], self.ns)
        _extract(dsec, ["get_events_vg"], self.ns, class_name="DSECDataset")
        spec = importlib.util.spec_from_file_location("_cmda_ref_utils", os.path.join(REF_DST, "mmseg", "datasets", "utils.py"))
        self.utils = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(self.utils)
        self.ic = {"torch": torch, "np": np, "Image": Image, "log_add": 50, "threshold": 0.1, "clip_range": 0.8}
        _extract(os.path.join(REF_DST, "create_cityscapes_image_change.py"), ["tensor_normalize_to_range", "get_image_change"],
                 self.ic)
        self.Image, self.F, self.torch = Image, F, torch

    def get_events_vg(self, t, x, y, p, rectify_map, width, height, bins, finish, start):
        """dsec.py:341-366"""
        return self.ns["get_events_vg"](_DSECStub(t, x, y, p, rectify_map, width, height, bins), finish, start)

    def isr(self, gray_u8, **kw):
        """utils.py:108-152 on a PIL image"""
        h, w = gray_u8.shape[:2]
        return self.utils.get_image_change_from_pil(self.Image.fromarray(gray_u8), w, h, **kw)

    def image_change(self, now_u8, front_u8):
        """create_cityscapes_image_change.py:16-35 (PIL 'L' in, PIL 'L' out)"""
        return self.ic["get_image_change"](self.Image.fromarray(now_u8), self.Image.fromarray(front_u8))


# ------------------------------------------------------------------------------------------ worker processes
def _make_item(spec):
    """Build one work item inside the worker (the parent never ships bulk data)."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from cmda_b200 import synth
    kind = spec["kind"]
    if kind == "events_vg":
        H, W = spec["H"], spec["W"]
        t, x, y, p = synth.make_events(spec["n"], H, W, window_us=spec.get("window_us", 50_000),
                                       t_base=spec.get("t_base", 10_000_000), seed=spec["seed"])
        rmap = synth.make_rectify_map(H, W, seed=spec["map_seed"])
        return dict(kind=kind, t=t, x=x, y=y, p=p, rmap=rmap, H=H, W=W, B=spec["bins"], n=spec["n"], post=spec.get("post"))
    if kind in ("isr", "pair"):
        now, front = synth.make_frame_pair(spec["H"], spec["W"], seed=spec["seed"])
        return dict(kind=kind, now=np.ascontiguousarray(now), front=np.ascontiguousarray(front), parms=spec.get("parms", {}))
    if kind == "c5":       # one sample of the train-step input path: a 50 ms window + the target ISR + the mixed-image ISR
        H, W = spec["H"], spec["W"]
        t, x, y, p = synth.make_events(spec["n"], H, W, seed=spec["seed"])
        rmap = synth.make_rectify_map(H, W, seed=spec["map_seed"])
        ev = dict(kind="events_vg", t=t, x=x, y=y, p=p, rmap=rmap, H=H, W=W, B=spec["bins"], n=spec["n"], post=spec["post"])
        oh, ow = spec["post"][5], spec["post"][4]
        imgs = [np.ascontiguousarray(synth.make_smooth_image(oh, ow, seed=spec["seed"] + k)) for k in (1, 2)]
        return dict(kind="c5", ev=ev, imgs=imgs, parms=spec["parms"])
    raise KeyError(kind)


def _run_item(ref, it):
    if it["kind"] == "events_vg":
        vg = ref.get_events_vg(it["t"], it["x"], it["y"], it["p"], it["rmap"], it["W"], it["H"], it["B"], it["n"] - 1, 0)
        post = it.get("post")
        if post:       # the loader's crop -> flip -> resize -> x3 (dsec.py:309-319), as the reference's statements
            x0, y0, cw, ch, ow, oh = post
            vg = vg[:, y0:y0 + ch, x0:x0 + cw]
            vg = ref.torch.flip(vg, dims=[2])
            vg = ref.F.interpolate(vg[None], size=(oh, ow), mode="bilinear", align_corners=False)[0]
            vg = vg.repeat(3, 1, 1)
        return float(vg.sum())
    if it["kind"] == "isr":
        return float(ref.isr(it["now"], **it["parms"]).sum())
    if it["kind"] == "c5":
        acc = _run_item(ref, it["ev"])
        for k, img in enumerate(it["imgs"]):
            acc += float(ref.isr(img, shift_direction=("rightdown", "leftup")[k], **it["parms"]).repeat(3, 1, 1).sum())
        return acc
    out = ref.image_change(it["now"], it["front"])
    return float(out.size[0])


def _worker(conn, specs, threads):
    import torch
    torch.set_num_threads(max(1, int(threads)))
    ref = Reference()
    items = [_make_item(s) for s in specs]
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            return
        kind = msg[1] if isinstance(msg, tuple) else None          # ("go", kind): only the items of that kind
        t0 = time.perf_counter()
        acc = 0.0
        for it in items:
            if kind is None or it["kind"] == kind:
                acc += _run_item(ref, it)
        conn.send((time.perf_counter() - t0, acc))


class WorkerPool:
    """One spawned process per host core (at most one per item); item k belongs to worker k mod n."""

    def __init__(self, specs, cores=None):
        cores = cores or host_cores()
        self.n_workers = max(1, min(cores, len(specs)))
        self.cores = cores
        ctx = mp.get_context("spawn")
        env_keep = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "CUDA_VISIBLE_DEVICES")}
        # torchrun exports OMP_NUM_THREADS=1; the workers size their own thread pools, and never touch a GPU
        os.environ.pop("OMP_NUM_THREADS", None)
        os.environ.pop("MKL_NUM_THREADS", None)
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        try:
            self.procs, self.conns = [], []
            for w in range(self.n_workers):
                a, b = ctx.Pipe()
                pr = ctx.Process(target=_worker, args=(b, specs[w::self.n_workers], max(1, cores // self.n_workers)), daemon=True)
                pr.start()
                self.procs.append(pr)
                self.conns.append(a)
        finally:
            for k, v in env_keep.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        for c in self.conns:
            assert c.recv() == "ready"

    def step(self, kind=None) -> float:
        """Every worker runs its items (of `kind`, if given) once; wall seconds from "go" to the last answer."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("go", kind))
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send("stop")
            except (BrokenPipeError, OSError):
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
