"""CPU-side checks: the C-ABI library loads and exports every symbol include/cmda_b200.h
declares (no compute calls without a GPU), argument validation that needs no device, and
the host logic (window bounds, sharding)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import cmda_b200
    from cmda_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return cmda_b200.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "cmda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(L):
    from cmda_b200 import _lib
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cmda_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures mirror the header one to one"


def test_version_and_strerror(L):
    assert L.cmda_version() == 100
    assert L.cmda_strerror(0) == b"ok"
    assert b"workspace" in L.cmda_strerror(-3)


def test_argument_validation_without_device(L):
    # all of these return before touching CUDA
    assert L.cmda_events_vg_workspace_bytes(1000, 0, 480, 640, 5, 0) == 0
    need = L.cmda_events_vg_workspace_bytes(5_000_000, 16, 480, 640, 5, 0)
    assert need >= 16 * 5 * 480 * 640 * 8
    assert L.cmda_searchsorted_right_u32(None, -1, None, 0, None, None) == -1
    assert L.cmda_events_vg_batch(None, None, None, None, None, None, 1, None, None, 480, 640, 5, None, 1.0, 1, 1,
                                  None, None, None, None, 0, 0, None) == -1
    assert L.cmda_isr_shift_u8(None, 2, 1, 4, 4, 1, 0, None, 0.0, 0.0, None, None, 0, None) == -1
    assert L.cmda_isr_shift_u8(None, 1, 1, 4, 4, 1, 9, None, 0.0, 0.0, None, None, 0, None) == -1
    assert L.cmda_events_norm_batch(None, 1, 0, None, 1.0, 1, None, 0, None) == -1


def test_voxel_modes_and_banded_workspace(L):
    """Mode ids mirror the header; AUTO resolves to FACTORED for DSEC-shaped grids; the opt-in BANDED mode needs
    FACTORED's workspace plus its record buffer (5 bytes per event for B > 1, 2 for B == 1, chunk-granular) and
    refuses grids whose rows do not fit a band (host code only: nothing here touches CUDA)."""
    from cmda_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cmda_b200.h")).read()
    for name, mode_id in _lib.VOXEL_MODES.items():
        assert re.search(rf"CMDA_VOXEL_{name.upper()}\s*=\s*{mode_id}\b", hdr), name
    n, S, H, W = 5_000_000, 16, 480, 640
    assert L.cmda_events_vg_resolved_mode(n * S, S, H, W, 5, _lib.VOXEL_AUTO) == _lib.VOXEL_FACTORED
    assert L.cmda_events_vg_resolved_mode(n * S, S, H, W, 5, _lib.VOXEL_BANDED) == _lib.VOXEL_BANDED
    assert L.cmda_events_vg_resolved_mode(n * S, S, H, W, 5, _lib.VOXEL_BANDED2) == _lib.VOXEL_BANDED2
    assert L.cmda_events_vg_resolved_mode(n * S, S, H, W, 5, _lib.VOXEL_BANDED2 + 1) == -1
    chunk = 8192
    for bins, rec_bytes in ((5, 5), (1, 2)):
        fact = L.cmda_events_vg_workspace_bytes(n * S, S, H, W, bins, _lib.VOXEL_FACTORED)
        band = L.cmda_events_vg_workspace_bytes(n * S, S, H, W, bins, _lib.VOXEL_BANDED)
        assert L.cmda_events_vg_workspace_bytes(n * S, S, H, W, bins, _lib.VOXEL_BANDED2) == band      # the two cuts share it
        extra = band - fact
        chunks = n * S // chunk + 2 * S
        assert extra >= rec_bytes * n * S                                   # every event has a record slot
        assert extra <= rec_bytes * chunks * chunk + 4 * chunks * (24 * bins + 2) + 4096     # ... and little else
        # more events -> more workspace, never less
        assert L.cmda_events_vg_workspace_bytes(2 * n * S, S, H, W, bins, _lib.VOXEL_BANDED) > band
    # a grid wider than one band can hold (B > 1: 24 576 cells): BANDED adds nothing, FACTORED's size remains
    wide = L.cmda_events_vg_workspace_bytes(1000, 1, 4, 30_000, 5, _lib.VOXEL_BANDED)
    assert wide == L.cmda_events_vg_workspace_bytes(1000, 1, 4, 30_000, 5, _lib.VOXEL_GLOBAL)


def test_product_has_no_oracle_import():
    """The product package must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "cmda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "cmda_oracle" not in text, f
                # ... nor through the CPU emulation of tests/emu (a kernel-logic checker, test infrastructure only)
                assert "emu_launch" not in text and "libcmda_b200_emu" not in text and "build_emu" not in text, f


def test_missing_library_fails_loudly(monkeypatch):
    from cmda_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcmda_b200.so")
    with pytest.raises(_lib.CmdaError):
        _lib.lib()


def test_cpu_tensors_without_gpu_raise():
    import torch
    import cmda_b200
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    e = torch.zeros(4)
    with pytest.raises(cmda_b200.CmdaError):
        cmda_b200.events_to_voxel_grid(e, e, e, e, 8, 8, 2)
    with pytest.raises(cmda_b200.CmdaError):
        cmda_b200.get_image_change_from_pil(np.zeros((4, 4), np.uint8), 4, 4, val_range=(1, 100), _threshold=0.04,
                                            _clip_range=0.2)


def test_window_bounds_matches_oracle():
    from cmda_b200 import window_bounds
    from oracle import cmda_oracle as O
    table = [-1, 10, 250, 250, 900, 700]
    for now in range(2, 6):
        for icr in (1, 2):
            for en in (-1, 100):
                for i in range(0, 1 if now < 4 else 2):
                    assert window_bounds(table, now, icr, en, i) == O.window_bounds(table, now, icr, en, i)
    assert window_bounds(table, 5, 1, -1, 0) is None          # start 900 > finish 700
    assert window_bounds(table, 3, 1, -1, 0) == (250, 250)     # single-event window is legal


def test_default_clip_range():
    from cmda_b200 import default_clip_range
    from oracle import cmda_oracle as O
    assert default_clip_range(330_000, 0) == O.default_clip_range(330_000, 0) == 330_000 / 500000 * 1.5


def test_sharding():
    from cmda_b200.sharding import shard_lpt, shard_round_robin
    for ws in (1, 2, 4, 8):
        seen = sorted(u for r in range(ws) for u in shard_round_robin(64, ws, r))
        assert seen == list(range(64))
        assert all(len(shard_round_robin(64, ws, r)) == 64 // ws for r in range(ws))
    costs = [20, 1, 1, 1, 20, 5, 5, 7, 3]
    parts = [shard_lpt(costs, 3, r) for r in range(3)]
    assert sorted(u for p in parts for u in p) == list(range(len(costs)))
    loads = [sum(costs[u] for u in p) for p in parts]
    assert max(loads) - min(loads) <= 5
    assert parts == [shard_lpt(costs, 3, r) for r in range(3)], "deterministic"


def test_index_txt_format(tmp_path):
    from cmda_b200 import write_index_txt
    p = tmp_path / "images_to_events_index.txt"
    write_index_txt([-1, 5, 77], str(p))
    assert p.read_text() == "-1\n5\n77\n"
    assert [int(v) for v in np.loadtxt(str(p), dtype=str, encoding="utf-8")] == [-1, 5, 77]


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    """No GPU / NVML here: the affinity helper reports what it did and leaves the process where it was."""
    import os
    from cmda_b200.sharding import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa(0)
    assert info["bound"] is False and os.sched_getaffinity(0) == before
    os.environ["CMDA_NO_NUMA_BIND"] = "1"
    try:
        assert bind_to_gpu_numa(0) == {"bound": False, "skipped": "CMDA_NO_NUMA_BIND"}
    finally:
        del os.environ["CMDA_NO_NUMA_BIND"]


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny sample: one JSON line with
    the contract's keys, the same metric / unit / config as the GPU arm, zero copy bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--events", "20000"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-400:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "voxelized_events_per_s" and line["unit"] == "Mevents/s"
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "Mevents/s", "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": 0}
    # the reference's own Python when oracle/_ref is staged (build() does that where /root/reference exists), else the C port
    from oracle import ref_runner
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_runner.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    # torchrun exports OMP_NUM_THREADS=1 to its ranks: the arm must not shrink its sample or its thread count with it
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    out2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "0", "--events", "20000"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out2.returncode == 0, out2.stderr[-400:]
    line2 = json.loads(out2.stdout.strip().splitlines()[-1])
    assert line2["cpu_baseline"]["cores"] == line["cpu_baseline"]["cores"] and "16 windows" in line2["cpu_baseline"]["sample"]
    assert {k: v for k, v in line2["config"].items() if k != "parallelism"} == \
           {k: v for k, v in line["config"].items() if k != "parallelism"}
    env["RANK"] = "1"
    out3 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "0", "--events", "20000"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert out3.returncode == 0 and out3.stdout.strip() == ""          # the other ranks exit without work
