"""Generates tests/golden/data_sequences.npz by running the REAL reference batch construction
(/root/reference/src/data.py) on seeded synthetic profiles.  Run in the build container only."""
import os
import random
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from src.data import get_test_sequences, get_train_sequences  # noqa: E402

rng = np.random.default_rng(20240611)
n_items, A, C, L, T = 200, 5, 6, 12, 20
attrs = rng.random((n_items, A), dtype=np.float32)
out = {"cfg": np.array([n_items, C, L, T])}
lens = [1, 2, 3, 4, 5, 9, 13, 14, 15, 30, 44]
for u, n in enumerate(lens):
    profile = [int(x) for x in rng.integers(1, n_items, size=n)]
    ctx = {(u, i): rng.random(C, dtype=np.float32) for i in profile}      # dict: repeated items share one row
    out[f"u{u}/profile"] = np.array(profile, np.int32)
    out[f"u{u}/ctx_rows"] = np.stack([ctx[(u, i)] for i in profile])
    for test in (True, False):
        for mode in ("train", "val", "test"):
            random.seed(1000 * u + 10 * test + len(mode))
            try:
                if mode == "train":
                    r = get_train_sequences(u, profile, L, attrs, ctx, test)
                else:
                    r = get_test_sequences(u, profile, L, T, attrs, ctx, mode, test)
            except IndexError:           # profile too short for the mode (CARCADataset filters these users out)
                continue
            key = f"u{u}/{mode}/{int(test)}"
            for name, v in zip(("p_x", "p_a", "p_c", "o_x", "o_a", "o_c", "y_true"), r):
                if name in ("p_a", "o_a"):
                    continue
                out[f"{key}/{name}"] = v
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_sequences.npz"), **out)
print("wrote", len(out), "arrays")
