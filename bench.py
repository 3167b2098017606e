#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json: voxelized Mevents/s).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one pass of the voxel path over one batch of synthetic windows per GPU:
BASELINE config C2 -- 16 DSEC-shaped (640x480) 50 ms windows of 5 M Poisson events each,
rectify_map remap + trilinear voxel grid (B bins) + events_norm.  Windows are independent,
so N GPUs each take their own 16 windows (weak scaling, no collective on the data path).

value   : whole-job Mevents/s with the events already resident in HBM (device timed)
e2e     : the same metric through the host-buffer front door (pinned host SoA in, host grids
          out, H2D and D2H inside the timed region)
roofline: the dominant kernel's algorithmic bytes / its own device time vs the measured HBM peak
cpu_baseline : the CPU oracle port (C restatement of the reference, exact reference order)
          timed on a bounded sample on the host cores
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 480, 640
WINDOWS_PER_GPU = 16
EVENTS_PER_WINDOW = 5_000_000
WINDOW_US = 50_000


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def algorithmic_bytes_per_window(n, bins):
    """events read once in native dtypes + map read once + final grid written once."""
    return 9 * n + 8 * H * W + 4 * bins * H * W


def make_workload(n_windows, n_events, seed_base):
    """Concatenated SoA store of n_windows independent Poisson windows + their inclusive bounds."""
    from cmda_b200 import synth
    ts, xs, ys, ps, starts, fins = [], [], [], [], [], []
    pos = 0
    for w in range(n_windows):
        t, x, y, p = synth.make_events(n_events, H, W, window_us=WINDOW_US, t_base=10_000_000 + w * WINDOW_US,
                                       seed=synth.seed_for(2, seed_base + w))
        ts.append(t); xs.append(x); ys.append(y); ps.append(p)
        starts.append(pos)
        fins.append(pos + n_events - 1)
        pos += n_events
    rmap = synth.make_rectify_map(H, W, seed=synth.seed_for(2, 999))
    return (np.concatenate(ts), np.concatenate(xs), np.concatenate(ys), np.concatenate(ps), rmap,
            np.array(starts, np.int64), np.array(fins, np.int64))


_SAMPLER_CHILD = r"""
import sys, time
idx, uuid = int(sys.argv[1]), sys.argv[2]
import pynvml
pynvml.nvmlInit()
try:
    h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode()) if uuid else pynvml.nvmlDeviceGetHandleByIndex(idx)
except Exception:
    h = pynvml.nvmlDeviceGetHandleByIndex(idx)
mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
out = sys.stdout
while True:
    try:
        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        out.write("%.6f %d %d %.1f %d\n" % (time.time(), sm, mx, pw, rs))
        out.flush()
    except Exception:
        pass
    time.sleep(0.0005)
"""


class ClockSampler:
    """SM clocks and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The timed
    region of this path lasts milliseconds, so a child process polls NVML as fast as it answers (about every
    millisecond) from before the warm-up on; the samples whose wall-clock stamp falls inside the timed
    region are the ones reported (all samples under load if the region was shorter than one poll)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.t_begin = self.t_end = None

    def start(self):
        uuid = ""
        try:
            import torch
            u = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            uuid = u if u.startswith("GPU-") else "GPU-" + u
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_CHILD, str(self.gpu_index), uuid],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def wait_ready(self, timeout=10.0):
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML sampler unavailable"]}
        time.sleep(0.005)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        samples = []
        for ln in list(self.lines):
            f = ln.split()
            if len(f) == 5:
                samples.append((float(f[0]), float(f[1]), float(f[2]), float(f[3]), int(f[4])))
        inside = [x for x in samples if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or 1e30)]
        # fall back to the samples taken under load since the warm-up began
        used = inside if inside else [x for x in samples if self.t_begin is None or x[0] <= (self.t_end or 1e30)][-20:]
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock samples"]}
        mask = 0
        for x in used:
            mask |= x[4]
        return {"sm_mhz": float(np.median([x[1] for x in used])), "sm_max_mhz": max(x[2] for x in used),
                "power_w_max": max(x[3] for x in used), "samples": len(used),
                "samples_in_timed_region": len(inside),
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_key):
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic_json.py from the capture of tools/r02_evidence.sh) and the
    commit that capture was taken from: a capture of another build says nothing about this one, so the label travels
    with the number.  Returns (bytes or None, label or None)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            key = kernel_key.split("+")[0].split(" ")[0]
            return d.get(key, d.get(kernel_key)), {"git": d.get("_git"), "source": d.get("_source")}
        except Exception:
            return None, None
    return None, None


# ------------------------------------------------------------------------------------ CPU legs
def host_cores():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU legs size
    their own pools and ignore it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_rate(args, n_windows=WINDOWS_PER_GPU, min_seconds=3.0):
    """Mevents/s of the oracle PORT (oracle/cmda_oracle.c: the reference's arithmetic in the reference's order, one
    window per OpenMP thread) on the step's windows: one untimed pass, then whole passes until `min_seconds` of
    wall time have been measured.  Secondary figure next to the reference's own Python."""
    from oracle import c_oracle
    c_oracle.build()
    cores = host_cores()
    t, x, y, p, rmap, starts, fins = make_workload(n_windows, args.events, seed_base=0)

    def one_pass():
        t0 = time.perf_counter()
        c_oracle.get_events_vg_batch(t, x, y, p, starts, fins, rmap, W, H, args.bins, nthreads=cores)
        return time.perf_counter() - t0

    one_pass()
    reps, dt = 0, 0.0
    while reps < 2 or (dt < min_seconds and reps < 12):
        dt += one_pass()
        reps += 1
    ev = float((fins - starts + 1).sum()) * reps
    return {"value": ev / dt / 1e6, "unit": "Mevents/s", "cores": min(cores, n_windows), "kind": "port",
            "sample": f"{reps} passes over {n_windows} windows x {args.events} events, one window per OpenMP thread, {dt:.1f} s wall"}


def reference_event_specs(args, n_windows=WINDOWS_PER_GPU):
    from cmda_b200 import synth
    return [dict(kind="events_vg", n=args.events, H=H, W=W, bins=args.bins, seed=synth.seed_for(2, w),
                 map_seed=synth.seed_for(2, 999), t_base=10_000_000 + w * WINDOW_US, window_us=WINDOW_US)
            for w in range(n_windows)]


def reference_rate(specs, units_per_step, steps, warmup, unit="Mevents/s", what="windows"):
    """The reference's OWN Python (oracle/_ref, staged by build()) on the host cores: one spawned worker process per
    core (at most one per work item), like the DataLoader workers the reference runs this path in
    (builder.py:151-163); a step = every item once.  Returns (cpu_baseline dict, per-step seconds)."""
    from oracle import ref_runner
    cores = host_cores()
    with ref_runner.WorkerPool(specs, cores=cores) as pool:
        for _ in range(warmup):
            pool.step()
        times = [pool.step() for _ in range(steps)]
        nw = pool.n_workers
    value = units_per_step * len(times) / sum(times) / 1e6
    return ({"value": value, "unit": unit, "cores": cores, "kind": "reference",
             "sample": f"{len(times)} steps x {len(specs)} {what}, the reference's own functions (oracle/_ref) in {nw} worker "
                       f"processes x {max(1, cores // nw)} torch threads, {sum(times):.1f} s wall"}, times)


ISR_SHIPPED = dict(shift_pixel=1, val_range=(0.01, 1.01), _threshold=0.005, _clip_range=0.1)      # configs/fusion/cs2dsec...:46-49
C3_IMAGES, C3_H, C3_W = 32, 1024, 2048
C5_SAMPLES, C5_EVENTS = 2, 330_000


def reference_side_legs(args, steps=1, warmup=1):
    """CPU baselines beside C2, C3 and C5 from ONE pool of worker processes running the reference's own functions
    (oracle/_ref): get_events_vg on the step's 16 windows; get_image_change / get_image_change_from_pil on C3's 32
    images of 2048x1024 (create_cityscapes_image_change.py:16-35, utils.py:108-152); per sample of C5 the event
    window + the loader's crop / flip / resize / x3 (dsec.py:286-320) + the two 512x512 ISR calls of a train step
    (dsec.py:258-261, dacs.py:741-744).  Returns a dict of cpu_baseline entries."""
    from cmda_b200 import synth
    from oracle import ref_runner
    cores = host_cores()
    specs = reference_event_specs(args)
    specs += [dict(kind="pair", H=C3_H, W=C3_W, seed=synth.seed_for(3, k)) for k in range(C3_IMAGES)]
    specs += [dict(kind="isr", H=C3_H, W=C3_W, seed=synth.seed_for(3, 100 + k), parms=dict(ISR_SHIPPED, shift_direction="rightdown"))
              for k in range(C3_IMAGES)]
    specs += [dict(kind="c5", n=C5_EVENTS, H=H, W=W, bins=1, seed=synth.seed_for(5, k), map_seed=synth.seed_for(5, 99),
                   post=(37, 61, 400, 400, 512, 512), parms=ISR_SHIPPED) for k in range(C5_SAMPLES)]
    out = {}
    with ref_runner.WorkerPool(specs, cores=cores) as pool:
        nw = pool.n_workers
        for key, kind, units, unit, what in (
                ("events", "events_vg", WINDOWS_PER_GPU * args.events / 1e6, "Mevents/s", f"{WINDOWS_PER_GPU} windows x {args.events} events"),
                ("frame_pair", "pair", C3_IMAGES, "images/s", f"{C3_IMAGES} pairs of {C3_W}x{C3_H}"),
                ("shift_pair_isr", "isr", C3_IMAGES, "images/s", f"{C3_IMAGES} images of {C3_W}x{C3_H}"),
                ("train_step_input_path", "c5", C5_SAMPLES, "samples/s", f"{C5_SAMPLES} samples (330 k-event window + crop/flip/resize + 2 ISR of 512x512)")):
            n_steps = (2 if key == "events" else 3) * steps
            for _ in range(warmup):
                pool.step(kind)
            times = [pool.step(kind) for _ in range(n_steps)]
            out[key] = {"value": units * len(times) / sum(times), "unit": unit, "cores": cores, "kind": "reference",
                        "sample": f"{len(times)} steps x {what}, the reference's own functions (oracle/_ref) in {nw} worker "
                                  f"processes x {max(1, cores // nw)} torch threads, {sum(times):.1f} s wall"}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the
    unmodified Python of oracle/_ref (dsec.py:341-366 get_events_vg: slice, t-normalise, rectify_map gather,
    events_to_voxel_grid, events_norm) when it is staged, else the C port of oracle/ (kind says which).  Same
    config / metric / unit as the CUDA arm; rank 0 alone runs, the other ranks exit."""
    if rank != 0:
        return
    from oracle import ref_runner
    events_per_step = WINDOWS_PER_GPU * args.events
    extra = {}
    if ref_runner.available():
        cpu, times = reference_rate(reference_event_specs(args), events_per_step, args.steps, min(args.warmup, 1))
        ms = 1e3 * sum(times) / len(times)
        extra["cpu_port"] = cpu_port_rate(args, min_seconds=2.0)
    else:
        cpu = cpu_port_rate(args, min_seconds=max(3.0, 0.5 * args.steps))
        ms = 1e3 * events_per_step / (cpu["value"] * 1e6)
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": "voxelized_events_per_s", "value": value, "unit": "Mevents/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "Mevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"C2: {WINDOWS_PER_GPU} DSEC 640x480 50ms windows x {args.events} events per GPU, "
                        f"rectify_map remap + voxel grid (B={args.bins}) + events_norm",
            "windows_per_gpu": WINDOWS_PER_GPU, "events_per_window": args.events, "bins": args.bins,
            "height": H, "width": W, "voxel_mode": args.mode, "parallelism": f"shard-by-window x{world}",
            "event_store": "GPU arm: packed P4 stream, 4 B/event (SoA DSEC arrays as a variant); CPU arm: the DSEC arrays",
            "l2": f"inputs ({4 * WINDOWS_PER_GPU * args.events / 1e6:.0f} MB of packed events per step; "
                  f"{9 * WINDOWS_PER_GPU * args.events / 1e6:.0f} MB as SoA arrays) and the {8 * WINDOWS_PER_GPU * args.bins * H * W / 1e6:.0f} MB "
                  "sensor-space grid the step fills and reads back exceed L2 (126 MB); no flush needed"}


# ------------------------------------------------------------------------------------ pseudo-event leg
def pseudo_events_leg(dev, peak, steps=10, warmup=3):
    """BASELINE config C3 on one GPU: 32 pairs of 2048x1024 uint8 gray frames -> frame-pair pseudo-events
    (f32 and u8 output) and shift-pair ISR (shipped cs2dsec parameters), images/s + achieved algorithmic
    GB/s ((1+1+4)HW, (1+1+1)HW, (1+4)HW bytes per image).  Device resident, CUDA-event timed;
    the 64 MB of inputs are re-read by the second pass out of L2 by design."""
    import torch
    import cmda_b200
    from cmda_b200 import synth
    Hc, Wc, S = 1024, 2048, 32
    base = [synth.make_frame_pair(Hc, Wc, seed=synth.seed_for(3, k)) for k in range(2)]
    now = torch.from_numpy(np.stack([base[k % 2][0] for k in range(S)])).to(dev)
    front = torch.from_numpy(np.stack([base[k % 2][1] for k in range(S)])).to(dev)
    for k in range(S):                      # make the 32 pairs distinct
        now[k] = torch.roll(now[k], shifts=(3 * k, 5 * k), dims=(0, 1))
        front[k] = torch.roll(front[k], shifts=(3 * k, 5 * k), dims=(0, 1))
    out_isr = torch.empty((S, 1, Hc, Wc), dtype=torch.float32, device=dev)

    def timed(fn):
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    res = {"config": f"C3: {S} pairs of {Wc}x{Hc} uint8 gray frames, device resident"}
    for name, fn, bpp in (
            ("frame_pair_f32", lambda: cmda_b200.image_change_batch(now, front, want_f32=True, want_u8=False), 6),
            ("frame_pair_u8", lambda: cmda_b200.image_change_batch(now, front, want_f32=False, want_u8=True), 3),
            ("shift_pair_isr", lambda: cmda_b200.isr_batch(now, 1, (0.01, 1.01), 0.005, 0.1, "rightdown", out=out_isr), 5)):
        ms = timed(fn)
        gbs = S * Hc * Wc * bpp / (ms * 1e-3) / 1e9
        res[name] = {"images_per_s": S / (ms * 1e-3), "ms_per_batch": ms, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    return res


# ------------------------------------------------------------------------------------ train-step input path leg
def train_step_input_leg(dev, steps=20, warmup=3):
    """BASELINE config C5: the input path of one CMDA fusion train step on one GPU, as the reference's loader
    and DACS step produce it for samples_per_gpu = 2 (configs/_base_/datasets/...512x512.py:11): per sample one
    50 ms DSEC window at real density (330 k events, create_dsec_dataset_txt.py:15-18) -> voxel grid -> events_norm
    -> crop 400x400 -> flip -> resize -> x3 (dsec.py:286-320), the target ISR of the warp image and the
    mixed-image ISR of the train step (dacs.py:729-744), at the reference's 512x512 and at BASELINE's 1024x512.
    Latency per step, device resident, CUDA-event timed."""
    import torch
    import cmda_b200
    from cmda_b200 import synth
    n, S = 330_000, 2
    ts, xs, ys, ps, starts, fins = [], [], [], [], [], []
    for k in range(S):
        t, x, y, p = synth.make_events(n, H, W, seed=synth.seed_for(5, k))
        ts.append(t); xs.append(x); ys.append(y); ps.append(p)
        starts.append(k * n); fins.append((k + 1) * n - 1)
    rmap = synth.make_rectify_map(H, W, seed=synth.seed_for(5, 99))
    store = cmda_b200.EventStore(np.concatenate(ts), np.concatenate(xs), np.concatenate(ys), np.concatenate(ps), rmap,
                                 height=H, width=W, device=dev)
    parms = dict(shift_pixel=1, val_range=(0.01, 1.01), _threshold=0.005, _clip_range=0.1)      # shipped cs2dsec isr_parms
    means = torch.tensor([123.675, 116.28, 103.53], device=dev).view(1, 3, 1, 1)
    stds = torch.tensor([58.395, 57.12, 57.375], device=dev).view(1, 3, 1, 1)
    res = {"config": f"C5: {S} samples/GPU, {n} events per 50 ms window, B=1 (shipped events_bins), device resident"}
    for name, (ow, oh) in (("crop_512x512", (512, 512)), ("crop_1024x512", (1024, 512))):
        g = torch.Generator(device="cpu").manual_seed(7)
        mixed = torch.randn((S, 3, oh, ow), generator=g).to(dev)
        warp = (torch.rand((S, oh, ow), generator=g) * 255).to(torch.uint8).to(dev)

        def step():
            ev = cmda_b200.events_vg_augmented_batch(store, starts, fins, 1, crop_xy=[(37, 61), (140, 20)], crop_size=(400, 400),
                                                     out_size=(ow, oh), flips=[1, 0], repeat=3)
            tgt = cmda_b200.isr_batch(warp, parms["shift_pixel"], parms["val_range"], parms["_threshold"],
                                      parms["_clip_range"], "rightdown")
            mix = cmda_b200.mixed_image_isr(mixed, means, stds, shift_direction="leftup", **parms)
            return ev, tgt, mix

        for _ in range(warmup):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        res[name] = {"ms_per_step": ms, "samples_per_s": S / (ms * 1e-3), "events_per_s": S * n / (ms * 1e-3)}
    return res


# ------------------------------------------------------------------------------------ C4: strong scaling by sample
C4_WINDOWS, C4_EVENTS = 64, 20_000_000


def c4_strong_scaling_leg(args, rank, world, dev, barrier, steps=3, warmup=1):
    """BASELINE config C4 as stated: 64 dense night windows of 20 M events each, sharded BY SAMPLE over the ranks
    (longest-processing-time greedy, cmda_b200.sharding.shard_lpt: what the reference's DistributedSampler does with
    whole samples, builder.py:135-141), no collective on the data path.  STRONG scaling: the 64 windows are the job at
    every N.  Window k is generated on the device from a seed of k alone, so every N voxelises the same 64 windows.
    Returns this rank's (seconds per pass, windows, events); the caller takes the max over ranks."""
    import torch
    import cmda_b200
    from cmda_b200 import sharding, synth
    mine = sharding.shard_lpt([C4_EVENTS] * C4_WINDOWS, world, rank)
    ts, xs, ys, ps = [], [], [], []
    for k in mine:
        g = torch.Generator(device=dev).manual_seed(40_000 + k)
        ts.append((torch.sort(torch.randint(0, WINDOW_US, (C4_EVENTS,), generator=g, device=dev, dtype=torch.int32))[0]
                   + (10_000_000 + k * WINDOW_US)).view(torch.uint32))
        xs.append(torch.randint(0, W, (C4_EVENTS,), generator=g, device=dev, dtype=torch.int16).view(torch.uint16))
        ys.append(torch.randint(0, H, (C4_EVENTS,), generator=g, device=dev, dtype=torch.int16).view(torch.uint16))
        ps.append(torch.randint(0, 2, (C4_EVENTS,), generator=g, device=dev, dtype=torch.uint8))
    rmap = synth.make_rectify_map(H, W, seed=synth.seed_for(4, 999))
    store = cmda_b200.EventStore(torch.cat(ts), torch.cat(xs), torch.cat(ys), torch.cat(ps), rmap, height=H, width=W, device=dev)
    del ts, xs, ys, ps
    n = len(mine)
    starts = np.arange(n, dtype=np.int64) * C4_EVENTS
    fins = starts + C4_EVENTS - 1
    group = 4                                   # windows per call: 80 M events, like the C2 step
    out = torch.empty((group, args.bins, H, W), dtype=torch.float32, device=dev)

    def one_pass():
        for g0 in range(0, n, group):
            g1 = min(g0 + group, n)
            cmda_b200.events_vg_batch(store, starts[g0:g1], fins[g0:g1], args.bins, mode=args.mode, out=out[:g1 - g0])

    for _ in range(warmup):
        one_pass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    del store, out
    return ms, n


def model_step_leg(dev, steps=3, warmup=2):
    """The consumer of the path, for scale: one MiT-b5 encoder + MLP head (SegFormer-B5; the reference's
    FusionEncoderDecoder runs two such backbones, encoder_decoder.py:626-1003) in PLAIN PyTorch -- transformers'
    SegformerForSemanticSegmentation with random weights -- forward + backward, float32, batch 2 x 3 x 512 x 512
    (samples_per_gpu = 2).  Not part of the path and not optimised here: printed next to C5 as SURVEY.md 8(d) asks."""
    try:
        import torch
        from transformers import SegformerConfig, SegformerForSemanticSegmentation
        cfg = SegformerConfig(num_channels=3, depths=[3, 6, 40, 3], hidden_sizes=[64, 128, 320, 512], num_attention_heads=[1, 2, 5, 8],
                              sr_ratios=[8, 4, 2, 1], decoder_hidden_size=768, num_labels=19)
        torch.manual_seed(0)
        model = SegformerForSemanticSegmentation(cfg).to(dev).train()
        x = torch.randn((2, 3, 512, 512), device=dev)
        y = torch.randint(0, 19, (2, 512, 512), device=dev)

        def step():
            model.zero_grad(set_to_none=True)
            out = model(pixel_values=x, labels=y)
            out.loss.backward()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        n_params = sum(p.numel() for p in model.parameters())
        del model, x, y
        torch.cuda.empty_cache()
        return {"ms_per_step": ms, "what": "plain PyTorch SegFormer-B5 (one MiT-b5 backbone + MLP head, random init), forward + backward, "
                                           "float32, batch 2 x 3 x 512 x 512", "parameters": n_params}
    except Exception as e:      # noqa: BLE001 -- an informational leg must not cost the bench line
        return {"error": f"{type(e).__name__}: {e}"[:200]}


# ------------------------------------------------------------------------------------ variants of the headline
def variants_leg(store, starts, fins, rmap, args, dev, steps=10, warmup=3):
    """SURVEY.md 8(d)'s other device-resident cases, reported beside the headline (same kernels, same timing
    rules: CUDA events, 3 warm-ups, inputs larger than L2 except C1): the shipped events_bins = 1; the 5-map variant
    of C2 (5 night sequences, window s -> map s mod 5, plans rebuilt inside every step); the skewed variant (10 % of
    the events on 1 % of the pixels: contention stress for the per-event atomics); C4's 20 M-event windows; C1's
    single 1 M-event window (latency).  The skewed and C4 inputs are generated on the device with torch's
    generator (they feed no parity test)."""
    import torch
    import cmda_b200
    from cmda_b200 import synth

    def timed(st, s0, f0, bins, mode=None, **kw):
        mode = mode or args.mode
        s0, f0 = np.asarray(s0, dtype=np.int64), np.asarray(f0, dtype=np.int64)
        out = torch.empty((len(s0), bins, H, W), dtype=torch.float32, device=dev)
        for _ in range(warmup):
            cmda_b200.events_vg_batch(st, s0, f0, bins, mode=mode, out=out, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(steps):
            cmda_b200.events_vg_batch(st, s0, f0, bins, mode=mode, out=out, **kw)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        n = int((f0 - s0 + 1).sum())
        return {"Mevents_per_s": n / (ms * 1e-3) / 1e6, "ms_per_step": ms, "windows": int(len(s0)), "events_per_step": n,
                "bins": bins}

    res = {}
    S = len(starts)
    if args.bins != 1:
        res["C2_bins_1"] = timed(store, starts, fins, 1)
    if args.mode in ("auto", "factored"):
        # the opt-in BANDED stage A (band partition + shared-memory accumulation instead of one L2 atomic per event;
        # bit-identical output): the same step, for the record
        res["C2_banded_stage_A"] = timed(store, starts, fins, args.bins, mode="banded")
        res["C2_banded2_stage_A"] = timed(store, starts, fins, args.bins, mode="banded2")
        res["C2_bins_1_red_stage_A"] = timed(store, starts, fins, 1, mode="factored")
        res["C2_bins_1_banded_stage_A"] = timed(store, starts, fins, 1, mode="banded")
    # the same step from the device-resident PACKED store (4 bytes per event instead of 9; bit-identical output)
    # the headline step from the SoA store (the four DSEC arrays, 9 bytes per event) and, at B = 1, from the packed one
    res["C2_soa_store"] = timed(store, starts, fins, args.bins)
    pstore = cmda_b200.PackedEventStore.from_event_store(store, plan=False)
    if args.bins != 1:
        res["C2_bins_1_packed_store"] = timed(pstore, starts, fins, 1)
    del pstore
    maps = np.stack([rmap] + [synth.make_rectify_map(H, W, seed=synth.seed_for(2, 900 + k)) for k in range(4)])
    store5 = cmda_b200.EventStore(store.t, store.x, store.y, store.p, maps, height=H, width=W, device=dev, plan=False)
    res["C2_five_maps"] = timed(store5, starts, fins, args.bins, map_ids=[s % 5 for s in range(S)])
    store5 = cmda_b200.EventStore(store.t, store.x, store.y, store.p, maps, height=H, width=W, device=dev)
    res["C2_five_maps_prebuilt_plans"] = timed(store5, starts, fins, args.bins, map_ids=[s % 5 for s in range(S)])
    del store5
    g = torch.Generator(device=dev).manual_seed(20251)
    n = len(store)
    hot = torch.randperm(H * W, generator=g, device=dev)[: H * W // 100]
    pick = hot[torch.randint(hot.numel(), (n,), generator=g, device=dev)]
    move = torch.rand((n,), generator=g, device=dev) < 0.1
    xs = torch.where(move, pick % W, store.x.view(torch.int16).to(torch.int64)).to(torch.int16).view(torch.uint16)
    ys = torch.where(move, pick // W, store.y.view(torch.int16).to(torch.int64)).to(torch.int16).view(torch.uint16)
    del pick, move
    skew = cmda_b200.EventStore(store.t, xs, ys, store.p, rmap, height=H, width=W, device=dev, plan=False)
    res["C2_skewed_10pct_on_1pct"] = timed(skew, starts, fins, args.bins)
    del skew, xs, ys
    n4, s4 = 20_000_000, 4
    t4 = torch.cat([torch.sort(torch.randint(0, 50_000, (n4,), generator=g, device=dev, dtype=torch.int32))[0] + 10_000_000
                    for _ in range(s4)]).view(torch.uint32)
    x4 = torch.randint(0, W, (s4 * n4,), generator=g, device=dev, dtype=torch.int16).view(torch.uint16)
    y4 = torch.randint(0, H, (s4 * n4,), generator=g, device=dev, dtype=torch.int16).view(torch.uint16)
    p4 = torch.randint(0, 2, (s4 * n4,), generator=g, device=dev, dtype=torch.uint8)
    c4 = cmda_b200.EventStore(t4, x4, y4, p4, rmap, height=H, width=W, device=dev, plan=False)
    res["C4_20M_event_windows"] = timed(c4, [k * n4 for k in range(s4)], [(k + 1) * n4 - 1 for k in range(s4)], args.bins)
    del c4, t4, x4, y4, p4
    n1 = min(1_000_000, int(fins[0] - starts[0] + 1))
    res["C1_single_1M_window"] = timed(store, [int(starts[0])], [int(starts[0]) + n1 - 1], args.bins)
    return res


# ------------------------------------------------------------------------------------ experimental: BANDED2
def banded2_leg(args):
    """The three forms of stage A (RED kernel, BANDED, BANDED with the second cut of the partition pass) on the C2 step with
    prebuilt map plans, in a process of its own (tools/banded2_check.py): bit-identity of the banded forms against
    FACTORED and ms per step of each."""
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "tools", "banded2_check.py"), "--events", str(args.events), "--bins", str(args.bins)]
    if args.bins != 1:
        cmd.append("1")
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        if res.returncode != 0:
            return {"error": f"exit {res.returncode}: {res.stderr.strip()[-300:]}"}
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:      # noqa: BLE001 -- an experimental leg must not cost the bench line
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------ GPU leg
class _StdoutGuard:
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version banner to fd 1 when
    NCCL_DEBUG is set): everything written to stdout while the guard is active goes to stderr; `emit` writes the
    line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())

    def close(self):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    guard = _StdoutGuard()

    import cmda_b200
    from cmda_b200 import _lib
    from cmda_b200.pipeline import HostEventsPipeline

    L = cmda_b200.lib()     # raises if the CUDA extension is missing
    torch.cuda.set_device(local_rank)
    from cmda_b200.sharding import bind_to_gpu_numa
    affinity = bind_to_gpu_numa(local_rank)      # before any pinned host buffer is allocated
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    t, x, y, p, rmap, starts, fins = make_workload(WINDOWS_PER_GPU, args.events, seed_base=rank * WINDOWS_PER_GPU)
    # headline: the events are resident in the framework's own device format -- the packed P4 stream (cmda_b200.packed:
    # 4 bytes per event instead of the 9 of the four DSEC arrays; the same stream the e2e leg ships over PCIe; outputs
    # bit-identical to the SoA store's, which is timed as the variant C2_soa_store) -- and the map-derived gather plans
    # are rebuilt inside every step (plan=False), like the reference re-reads the map for every sample; the variant
    # with plans prebuilt once per sequence is reported separately below
    soa_store = cmda_b200.EventStore(t, x, y, p, rmap, height=H, width=W, device=dev, plan=False)
    store = cmda_b200.PackedEventStore.from_event_store(soa_store, plan=False)
    out = torch.empty((WINDOWS_PER_GPU, args.bins, H, W), dtype=torch.float32, device=dev)
    events_per_step = int((fins - starts + 1).sum())

    def step():
        cmda_b200.events_vg_batch(store, starts, fins, args.bins, mode=args.mode, out=out)

    # ---- device-resident timing ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # per-step phase events (cmda_profiler_attach) for the per-kernel roofline
    n_phase = 8
    phase_events = [[L.cmda_event_create() for _ in range(n_phase)] for _ in range(args.steps)]
    phase_arrays = [(ctypes.c_void_p * n_phase)(*pe) for pe in phase_events]
    used = 0
    barrier()
    sampler.mark_begin()
    ev[0].record()
    for k in range(args.steps):
        L.cmda_profiler_attach(phase_arrays[k], n_phase)
        step()
        used = L.cmda_profiler_detach()
    ev[1].record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev[0].elapsed_time(ev[1])
    phase_ms = np.zeros(max(used - 1, 0))
    for pe in phase_events:
        for j in range(used - 1):
            ms = ctypes.c_float()
            L.cmda_event_elapsed_ms(pe[j], pe[j + 1], ctypes.byref(ms))
            phase_ms[j] += ms.value / args.steps
    for pe in phase_events:
        for e in pe:
            L.cmda_event_destroy(e)

    # ---- same step with the rectify-map plans prebuilt once (EventStore default) -----------------
    store_p = cmda_b200.PackedEventStore(store.rec, store.h_ms_to_idx, store.rectify_map, height=H, width=W, device=dev, plan=True)
    for _ in range(3):
        cmda_b200.events_vg_batch(store_p, starts, fins, args.bins, mode=args.mode, out=out)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        cmda_b200.events_vg_batch(store_p, starts, fins, args.bins, mode=args.mode, out=out)
    p1.record()
    barrier()
    planned_ms = p0.elapsed_time(p1) / args.steps

    # ---- end to end through the host-buffer front door ----------------------------------------
    # Three ways a caller can hold the events (the timed region of each includes every copy it needs, every step):
    #   p3        pinned host memory, the 3-byte wire form of the packed stream (packed once when the pipeline / cache is
    #             built; unpacked to P4 records on the device, inside the timed region)                          <- `e2e`
    #   p4        pinned host memory, packed stream (4 B/event)
    #   soa       pinned host memory, the four DSEC arrays as the reference slices them (9 B/event)
    #   resident  events already on the device (DSECEvents.from_cache): only the grids travel, device -> host
    host_out = torch.empty((WINDOWS_PER_GPU, args.bins, H, W), dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def time_calls(fn):
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            fn()
        e1.record()
        barrier()
        return max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / e2e_steps

    e2e_legs = {}
    for wire in ("p3", "p4", "soa"):
        pipe = HostEventsPipeline(t, x, y, p, rmap, args.bins, H, W, device=dev, windows_per_group=args.e2e_group, mode=args.mode,
                                  max_window_events=args.events, wire=wire)
        ms = time_calls(lambda: pipe(starts, fins, out=host_out))
        h2d_w, d2h_w = pipe.bytes_per_call(starts, fins)
        # the paths must agree bit for bit (same kernels, same inputs)
        e2e_legs[wire] = {"ms": ms, "h2d": h2d_w, "d2h": d2h_w, "same": bool(torch.equal(host_out, out.cpu()))}
        pipe.close()
        del pipe
    store_r = store_p
    half = WINDOWS_PER_GPU // 2
    d_out2 = [torch.empty((half, args.bins, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
    side = torch.cuda.Stream(dev)

    def resident_step():
        # two halves: the second half's kernels run while the first half's grids leave the device
        main = torch.cuda.current_stream(dev)
        for h in range(2):
            sl = slice(h * half, (h + 1) * half)
            cmda_b200.events_vg_batch(store_r, starts[sl], fins[sl], args.bins, mode=args.mode, out=d_out2[h])
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                host_out[sl].copy_(d_out2[h], non_blocking=True)
        side.synchronize()

    ms = time_calls(resident_step)
    e2e_legs["resident"] = {"ms": ms, "h2d": 0, "d2h": 4 * WINDOWS_PER_GPU * args.bins * H * W,
                            "same": bool(torch.equal(host_out, out.cpu()))}
    del store_r, store_p, d_out2
    HEAD_WIRE = "p3"
    e2e_ms, h2d, d2h, same = (e2e_legs[HEAD_WIRE][k] for k in ("ms", "h2d", "d2h", "same"))

    # ---- the host link alone: every rank copies 256 MB pinned <-> device at the same time (what bounds e2e at N > 1) ----
    probe_h = torch.empty((256 << 20,), dtype=torch.uint8).pin_memory()
    probe_h2 = torch.empty((256 << 20,), dtype=torch.uint8).pin_memory()
    probe_d = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    probe_d2 = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def link_probe(up, down):
        """GB/s per direction with `up` (H2D) and / or `down` (D2H) copies in flight on every rank."""
        def issue():
            if up:
                with torch.cuda.stream(s_up):
                    probe_d.copy_(probe_h, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    probe_h2.copy_(probe_d2, non_blocking=True)
        issue()
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            issue()
        s_up.synchronize(); s_down.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return 4 * (256 << 20) / dt / 1e9

    link = torch.tensor([-link_probe(True, False), -link_probe(False, True), -link_probe(True, True)], dtype=torch.float64, device=dev)
    del probe_h, probe_h2, probe_d, probe_d2

    # ---- C4: 64 x 20 M-event windows sharded by sample, strong scaling ---------------------------------
    c4_ms, c4_n = (0.0, 0)
    if not args.no_c4:
        del store, soa_store
        torch.cuda.empty_cache()
        c4_ms, c4_n = c4_strong_scaling_leg(args, rank, world, dev, barrier)
        soa_store = cmda_b200.EventStore(t, x, y, p, rmap, height=H, width=W, device=dev, plan=False)

    # ---- max over ranks ----------------------------------------------------------------------
    times = torch.cat([torch.tensor([ms_total, e2e_ms, planned_ms, e2e_legs["soa"]["ms"], e2e_legs["resident"]["ms"], c4_ms],
                                    dtype=torch.float64, device=dev), link,
                       torch.tensor([e2e_legs["p4"]["ms"]], dtype=torch.float64, device=dev)])
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    link_gbps = [-float(v) for v in times[6:9]]       # the slowest rank's link: H2D alone, D2H alone, each direction with both busy
    ms_total, e2e_ms, planned_ms = float(times[0]), float(times[1]), float(times[2])
    e2e_legs["soa"]["ms"], e2e_legs["resident"]["ms"], c4_ms = float(times[3]), float(times[4]), float(times[5])
    e2e_legs["p4"]["ms"] = float(times[9])

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = world * events_per_step / (ms_per_step * 1e-3) / 1e6
        e2e_value = world * events_per_step / (e2e_ms * 1e-3) / 1e6
        peak, peak_src = measured_peaks()
        resolved = L.cmda_events_vg_resolved_mode(events_per_step, WINDOWS_PER_GPU, H, W, args.bins,
                                                  _lib.VOXEL_MODES[args.mode])
        if resolved == _lib.VOXEL_GLOBAL:
            names = ["memset(int64 grid)", "voxel_scatter_global_kernel", "convert_stats_kernel", "norm_apply_kernel"]
            launches_per_step = 3
        elif resolved == _lib.VOXEL_FACTORED:
            names = ["memset(sensor grid)", "launch of rectify_index_{build,sort}+stencil_build+out_tile_box kernels (side stream: they run under stage A)",
                     "sensor_accumulate_kernel", "rectify_gather+regroup_partials kernels", "norm_apply_kernel"]
            # kernels only (memsets not counted): 4 plan kernels, sensor_accumulate, fallback_zero + fallback_scatter (the
            # capacity guard: always launched), rectify_gather, regroup_partials, norm_apply
            launches_per_step = 10
        elif resolved in (_lib.VOXEL_BANDED, _lib.VOXEL_BANDED2):
            names = ["memset(none: every sensor-grid cell is stored)", "launch of rectify_index_{build,sort}+stencil_build+out_tile_box kernels (side stream: they run under stage A)",
                     "band_partition_kernel", "band_accumulate(+fixup) kernel", "rectify_gather+regroup_partials kernels",
                     "norm_apply_kernel"]
            launches_per_step = 11  # as above with band_partition + band_accumulate in place of sensor_accumulate
        else:
            names = ["memset(int64 grid)", "tile_bbox+tile_count+tile_scan kernels", "tile_partition_kernel",
                     "tile_accumulate_kernel", "convert_stats_kernel", "norm_apply_kernel"]
            launches_per_step = 7
        phases = {names[j] if j < len(names) else f"phase{j}": float(phase_ms[j]) for j in range(len(phase_ms))}
        kernel_phases = {k: v for k, v in phases.items() if not k.startswith("memset")}
        dom = max(kernel_phases, key=kernel_phases.get) if kernel_phases else None
        alg = WINDOWS_PER_GPU * algorithmic_bytes_per_window(args.events, args.bins)
        roofline = None
        if dom:
            achieved = alg / (kernel_phases[dom] * 1e-3) / 1e9
            traffic, traffic_label = ncu_traffic(dom)
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_capture": traffic_label, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg, "kernel_ms": kernel_phases[dom],
                        "phase_ms": phases,
                        "whole_step": {"achieved": alg / (ms_per_step * 1e-3) / 1e9,
                                       "frac": alg / (ms_per_step * 1e-3) / 1e9 / peak}}
        # CPU baseline on a bounded sample of the same workload (rank 0, N=1 only): the reference's own Python
        # (oracle/_ref) in worker processes, same protocol as `--impl reference`; the C port as a second figure
        cpu = cpu_port = side = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_runner
            cpu_port = cpu_port_rate(args)
            if ref_runner.available():
                side = reference_side_legs(args)
                cpu = side.pop("events")
            else:
                cpu = cpu_port
        pseudo = c5 = variants = None
        if world == 1 and not args.no_pseudo:
            pseudo = pseudo_events_leg(dev, peak)
            c5 = train_step_input_leg(dev)
            c5["model_step"] = model_step_leg(dev)
            if side:       # the reference's CPU path beside each GPU number (BASELINE.json: north_star)
                for k in ("frame_pair", "shift_pair_isr"):
                    pseudo["cpu_baseline_" + k] = side[k]
                c5["cpu_baseline"] = side["train_step_input_path"]
        if world == 1 and not args.no_variants:
            del host_out
            variants = variants_leg(soa_store, starts, fins, rmap, args, dev)
        experimental = None
        if world == 1 and not args.no_variants and args.mode in ("auto", "factored"):
            experimental = {"stage_A_forms": banded2_leg(args)}      # a process of its own (see there)
        line = {
            "metric": "voxelized_events_per_s", "value": value, "unit": "Mevents/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "resolved_mode": _lib.VOXEL_MODE_NAMES[resolved], "rank0_cpu_affinity": affinity,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mevents/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "matches_device_path": same,
                    "wire": "p3: the 3-byte wire form of the packed event stream from pinned host memory (cmda_b200.packed; packed "
                            "once, outside the timed region, like the reference's events.h5 decode), unpacked to the 4-byte "
                            "records on the device inside the timed region; grids back to pinned host memory",
                    "windows_per_group": args.e2e_group,
                    "h2d_GBps_per_gpu": h2d / (e2e_ms * 1e-3) / 1e9,
                    "host_link_probe": {"h2d_GBps_per_gpu_all_ranks_copying": link_gbps[0],
                                        "d2h_GBps_per_gpu_all_ranks_copying": link_gbps[1],
                                        "each_direction_GBps_per_gpu_both_busy": link_gbps[2],
                                        "what": "256 MB pinned <-> device copies issued by every rank at once, slowest rank: the "
                                                "ceiling of any host-fed path at this N on this box"}},
            "e2e_other_wires": {
                k: {"value": world * events_per_step / (v["ms"] * 1e-3) / 1e6, "unit": "Mevents/s", "ms_per_step": v["ms"],
                    "h2d_bytes_per_step": v["h2d"], "d2h_bytes_per_step": v["d2h"], "matches_device_path": v["same"]}
                for k, v in e2e_legs.items() if k != HEAD_WIRE},
            "gpu_launches": launches_per_step * args.steps,
            "with_prebuilt_map_plans": {"value": world * events_per_step / (planned_ms * 1e-3) / 1e6, "unit": "Mevents/s",
                                        "ms_per_step": planned_ms,
                                        "note": "cmda_rectify_plan_build once per sequence instead of inside every step"},
            "c4_strong_scaling": None if args.no_c4 else {
                "config": f"C4: {C4_WINDOWS} windows x {C4_EVENTS} events (B={args.bins}) sharded by sample (shard_lpt) over {world} GPU(s), "
                          "device resident, no collective; the same 64 windows at every N",
                "Mevents_per_s": C4_WINDOWS * C4_EVENTS / (c4_ms * 1e-3) / 1e6, "ms_per_pass": c4_ms, "n_gpus": world,
                "windows_rank0": c4_n, "scaling": "strong"},
            "roofline": roofline, "cpu_baseline": cpu, "cpu_port": cpu_port, "pseudo_events": pseudo, "train_step_input_path": c5,
            "variants": variants, "experimental": experimental,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    guard.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cmda_b200", choices=["cmda_b200", "reference"])
    ap.add_argument("--bins", type=int, default=5)
    ap.add_argument("--events", type=int, default=EVENTS_PER_WINDOW)
    ap.add_argument("--mode", default="auto", choices=["auto", "global", "tiled", "factored", "banded", "banded2"])
    ap.add_argument("--e2e-group", type=int, default=2, help="windows per host->device copy group of the e2e pipeline")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 strong-scaling leg (64 x 20 M-event windows)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pseudo", action="store_true", help="skip the pseudo-event (config C3) leg")
    ap.add_argument("--no-variants", action="store_true", help="skip the other device-resident cases of SURVEY.md 8(d)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cmda_b200" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        cmd += sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
