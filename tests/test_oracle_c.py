"""Pin the C restatement (oracle/cmda_oracle.c) against the golden fixtures and
against the numpy oracle."""
import numpy as np
import pytest

import golden_io
from oracle import c_oracle as C
from oracle import cmda_oracle as O

VOXEL = golden_io.load("voxel")
NORM = golden_io.load("norm")
VG = golden_io.load("events_vg")
ISR = golden_io.load("isr")
IC = golden_io.load("image_change")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", sorted(VOXEL))
def test_voxel_bit_exact(name):
    c = VOXEL[name]
    got = C.voxel_grid(c["time"], c["x"], c["y"], c["pol"], int(c["width"]), int(c["height"]), int(c["bins"]))
    assert np.array_equal(bits(got), bits(c["grid"]))


@pytest.mark.parametrize("name", sorted(k for k in NORM if k.startswith("norm_")))
def test_norm(name):
    c = NORM[name]
    events = NORM["normgrid_" + str(c["grid"])]["events"]
    got = C.events_norm(events, float(c["clip_range"]), float(c["final_range"]), bool(c["enforce"]))
    np.testing.assert_allclose(got, c["result"], rtol=0, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("name", sorted(VG))
def test_events_vg(name):
    c = VG[name]
    rmap = golden_io.rectify_map_of(c)
    clip = [float(c["clip"][0])] if c["clip"].size else None
    got, raw = C.get_events_vg_batch(c["t"], c["x"], c["y"], c["p"], [int(c["start"])], [int(c["finish"])], rmap,
                                     int(c["width"]), int(c["height"]), int(c["bins"]), clips=clip, return_raw=True)
    np.testing.assert_allclose(got[0], c["result"], rtol=0, atol=1e-5)
    _, raw_py = O.get_events_vg(c["t"], c["x"], c["y"], c["p"], rmap, int(c["width"]), int(c["height"]),
                                int(c["bins"]), int(c["finish"]), int(c["start"]), clip_range=clip[0] if clip else None,
                                return_raw=True)
    assert np.array_equal(bits(raw[0]), bits(raw_py))


@pytest.mark.parametrize("name", sorted(k for k in ISR if "lut" in ISR[k]))
def test_isr_bit_exact(name):
    c = ISR[name]
    gray = ISR["isr_input"]["gray"]
    got = C.isr_batch(gray, int(c["shift_pixel"]), tuple(float(v) for v in c["val_range"]), float(c["threshold"]),
                      float(c["clip_range"]), str(c["direction"]))
    assert np.array_equal(bits(got), bits(c["result"]))


@pytest.mark.parametrize("name", sorted(IC))
def test_image_change(name):
    c = IC[name]
    _, got = C.image_change_batch(c["now"], c["front"])
    assert np.array_equal(got[0], c["result"])


def test_batch_threads_deterministic():
    from cmda_b200 import synth
    t, x, y, p = synth.make_events(30000, 48, 64, seed=5)
    rmap = synth.make_rectify_map(48, 64, seed=6)
    starts, fins = [0, 1000, 5000, 7], [29999, 20000, 5000, 15000]
    a = C.get_events_vg_batch(t, x, y, p, starts, fins, rmap, 64, 48, 5, nthreads=1)
    b = C.get_events_vg_batch(t, x, y, p, starts, fins, rmap, 64, 48, 5, nthreads=4)
    assert np.array_equal(bits(a), bits(b))
