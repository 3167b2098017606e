"""Pin the CPU oracle (oracle/cmda_oracle.py) against the golden fixtures, which are
outputs of the reference's own functions (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_io
from oracle import cmda_oracle as O

VOXEL = golden_io.load("voxel")
NORM = golden_io.load("norm")
VG = golden_io.load("events_vg")
ISR = golden_io.load("isr")
IC = golden_io.load("image_change")
INDEX = golden_io.load("index")

# Normalised grids live in [-1, 1]; north_star's tolerance for them is 1e-5.  The
# reference's float32 torch.sum order is a third-party detail: it moves mean/std by
# ~1e-7 relative, which the clamp-then-min-max step amplifies by ~1/clip_range
# (4e-6 at the extreme clip_range=0.018 fixture, <= 4e-7 at clip_range >= 0.75).
NORM_ATOL = 1e-5


@pytest.mark.parametrize("name", sorted(VOXEL))
def test_voxel_grid_bit_exact(name):
    c = VOXEL[name]
    got = O.events_to_voxel_grid(c["time"], c["x"], c["y"], c["pol"], int(c["width"]), int(c["height"]), int(c["bins"]))
    assert got.dtype == np.float32 and got.shape == c["grid"].shape
    assert np.array_equal(got.view(np.uint32), c["grid"].view(np.uint32)), "sequential f32 order must be bit-exact"


@pytest.mark.parametrize("name", sorted(VOXEL))
def test_voxel_f64_truth_close(name):
    c = VOXEL[name]
    truth = O.voxel_grid_f64(c["time"], c["x"], c["y"], c["pol"], int(c["width"]), int(c["height"]), int(c["bins"]))
    _, aux = O.events_to_voxel_grid(c["time"], c["x"], c["y"], c["pol"], int(c["width"]), int(c["height"]),
                                    int(c["bins"]), return_aux=True)
    tol = 1e-5 * np.maximum(np.abs(c["grid"]), aux["abs_weight_sum"]) + 1e-12
    assert np.all(np.abs(truth - c["grid"]) <= tol)


@pytest.mark.parametrize("name", sorted(k for k in NORM if k.startswith("norm_")))
def test_events_norm(name):
    c = NORM[name]
    events = NORM["normgrid_" + str(c["grid"])]["events"]
    got = O.events_norm(events, clip_range=float(c["clip_range"]), final_range=float(c["final_range"]),
                        enforce_no_events_zero=bool(c["enforce"]))
    ref = c["result"]
    assert got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=0, atol=NORM_ATOL, equal_nan=True)
    # zeros of the input stay exactly where the reference leaves them
    assert np.array_equal(got == 0, ref == 0) or np.allclose(got, ref, atol=NORM_ATOL, equal_nan=True)


@pytest.mark.parametrize("name", sorted(VG))
def test_get_events_vg(name):
    c = VG[name]
    rmap = golden_io.rectify_map_of(c)
    clip = c["clip"]
    clip_range = None
    if clip.size:
        assert clip[0] == clip[1]
        clip_range = float(clip[0])  # degenerate uniform(lo, lo): no RNG dependence
    got, raw = O.get_events_vg(c["t"], c["x"], c["y"], c["p"], rmap, int(c["width"]), int(c["height"]),
                               int(c["bins"]), int(c["finish"]), int(c["start"]), clip_range=clip_range,
                               return_raw=True)
    np.testing.assert_allclose(got, c["result"], rtol=0, atol=NORM_ATOL)


def test_one_event_window_is_minus_one():
    """Single-timestamp window: NaN t -> all corners masked -> zero raw grid (Q3); the
    reference's min-max of the all-zero negative part then yields -1 everywhere."""
    c = VG["vg_one_event"]
    assert np.all(c["result"] == -1.0)


def test_pil_gray_formula():
    c = ISR["isr_input"]
    assert np.array_equal(O.pil_gray_L(c["rgb"]), c["gray"])


@pytest.mark.parametrize("name", sorted(k for k in ISR if k.startswith("isr_") and "result" in ISR[k] and "lut" in ISR[k]))
def test_isr_bit_exact(name):
    c = ISR[name]
    rgb = ISR["isr_input"]["rgb"]
    vr = tuple(float(v) for v in c["val_range"])
    # int-valued ranges were ints in the reference call; float() of them is the same number
    got = O.get_image_change_from_pil(rgb, rgb.shape[1], rgb.shape[0], shift_pixel=int(c["shift_pixel"]),
                                      val_range=vr, _threshold=float(c["threshold"]),
                                      _clip_range=float(c["clip_range"]), shift_direction=str(c["direction"]))
    assert np.array_equal(O.log_lut_val_range(vr).view(np.uint32), c["lut"].view(np.uint32))
    assert got.shape == c["result"].shape
    assert np.array_equal(got.view(np.uint32), c["result"].view(np.uint32))


def test_isr_flat_and_direct():
    c = ISR["isr_flat"]
    got = O.get_image_change_from_pil(c["rgb"], 24, 16, val_range=(1, 100), _threshold=0.04, _clip_range=0.2,
                                      shift_pixel=3)
    assert np.array_equal(got.view(np.uint32), c["result"].view(np.uint32))
    c = ISR["get_ic_direct"]
    got = O.get_ic(c["front"], c["now"], val_range=(1, 100), threshold=0.04, clip_range=0.2)
    assert np.array_equal(got.view(np.uint32), c["result"].view(np.uint32))


def test_isr_auto_threshold_raises():
    with pytest.raises(ValueError):
        O.get_image_change_from_pil(np.zeros((4, 4, 3), np.uint8), 4, 4, auto_threshold=("x", "image_gray"))


@pytest.mark.parametrize("name", sorted(IC))
def test_image_change_pair_bit_exact(name):
    c = IC[name]
    got = O.get_image_change(c["now"], c["front"])
    assert got.dtype == np.uint8
    assert np.array_equal(got, c["result"])
    assert np.array_equal(O.log_lut_log_add(50).view(np.uint32), c["lut"].view(np.uint32))


def test_images_to_events_index():
    c = INDEX["index_table"]
    got = O.images_to_events_index(c["t"], int(c["t_offset"]), c["ms_to_idx"], c["timestamps"])
    assert got == [int(v) for v in c["result"]]
    assert -1 in got and max(got) == len(c["t"]) - 1


def test_mixed_image_to_gray_matches_torch_and_pil():
    """a9 (dacs.py:730-733): the oracle's float32 restatement against the reference's own statements
    executed with torch + PIL (both present here): denorm, clamp, * 255, np.uint8, fromarray, convert('L')."""
    import torch
    from PIL import Image
    rng = np.random.default_rng(11)
    means = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1)      # img_norm_cfg, dsec.py:325-326
    stds = torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
    img = torch.from_numpy(rng.normal(0.0, 1.4, size=(1, 3, 37, 53)).astype(np.float32))
    mixed = torch.clamp(img.mul(stds).add(means) / 255.0, 0, 1) * 255     # dacs.py:730 + dacs_transforms.py:52-53
    mixed = np.transpose(mixed.cpu().numpy()[0], (1, 2, 0))               # dacs.py:731
    pil = Image.fromarray(np.uint8(mixed))                                # dacs.py:733
    gray, rgb = O.mixed_image_to_gray(img[0].numpy(), means.numpy().ravel(), stds.numpy().ravel(), return_rgb=True,
                                      cuda_division=False)   # torch on the CPU divides; on CUDA it multiplies by fl(1/255)
    assert np.array_equal(rgb, np.asarray(pil))
    assert np.array_equal(gray, np.asarray(pil.convert('L')))             # utils.py:126


@pytest.mark.parametrize("mode,in_hw,out_wh", [("L", (64, 96), (48, 32)), ("RGB", (50, 70), (35, 25)), ("L", (40, 60), (90, 55)),
                                               ("RGB", (33, 47), (47, 33)), ("L", (128, 256), (128, 64)), ("L", (31, 64), (64, 17))])
def test_pil_resize_bilinear_matches_pillow(mode, in_hw, out_wh):
    """f-4: the restated Resample.c algorithm against the installed Pillow, bit for bit (down-, up- and mixed scaling,
    2:1 like 2048x1024 -> 1024x512, odd sizes, one axis unchanged)."""
    from PIL import Image
    rng = np.random.default_rng(7)
    shape = in_hw + ((3,) if mode == "RGB" else ())
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img, mode=mode).resize(out_wh, resample=Image.BILINEAR))
    assert np.array_equal(O.pil_resize_bilinear(img, out_wh), ref)
