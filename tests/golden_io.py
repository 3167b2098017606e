"""Read the committed golden fixtures (tests/golden/*.npz) as {case: {field: array}}."""
import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(group: str) -> dict:
    data = np.load(os.path.join(GOLDEN_DIR, f"{group}.npz"))
    cases: dict = {}
    for key in data.files:
        case, field = key.split("/", 1)
        cases.setdefault(case, {})[field] = data[key]
    return cases


def rectify_map_of(case: dict) -> np.ndarray:
    """Fixtures of full DSEC size store only the map's seed and sha256 digest."""
    from cmda_b200 import synth
    m = case["rectify_map"]
    if m.dtype == np.uint8 and m.ndim == 1:
        full = synth.make_rectify_map(int(case["height"]), int(case["width"]), seed=int(case["map_seed"]))
        digest = np.frombuffer(hashlib.sha256(full.tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(digest, m), "synth.make_rectify_map drifted from the golden fixture"
        return full
    return m
