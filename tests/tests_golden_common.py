import os

import numpy as np

from conftest import GOLDEN_DIR

RENDERS = {"c1_whitted": "c1_torusknot", "c2_reflective": "c2_monkey", "c2_null": "c2_monkey_null",
           "mix_deterministic": "deterministic_mix", "c3_preview": "c3_unitychan", "c3_path": "c3_unitychan",
           "default_path": "default_scene"}


def golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    assert os.path.exists(path), f"{path} missing (tests/golden/make_golden.py)"
    return np.load(path)
