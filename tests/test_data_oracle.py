"""CPU: the data oracle against the fixtures produced by the real reference loader, and the device-side
batch construction kernels (through tools/emu) against the same fixtures."""
import numpy as np
import pytest

import data_suite as S
from oracle import data_oracle as D


def test_data_oracle_matches_reference_fixtures():
    z, n_items, C, L, T, users, profiles, ctx_rows = S.load()
    checked = 0
    for u in users:
        for test in (True, False):
            for mode in ("train", "val", "test"):
                key = f"u{u}/{mode}/{int(test)}"
                if key + "/p_x" not in z.files:
                    continue
                g = {k: z[f"{key}/{k}"] for k in ("p_x", "p_c", "o_x", "o_c", "y_true")}
                if mode == "train":
                    negs = [int(g["o_x"][L + L - 1 - i]) for i in range(L)]     # neg_sample[i] sits at L + (L-1-i)
                    got = D.train_sequences(profiles[u], ctx_rows[u], L, negs, test)
                else:
                    got = D.test_sequences(profiles[u], ctx_rows[u], L, [int(v) for v in g["o_x"][1:]], mode, test)
                for name, v in zip(("p_x", "p_c", "o_x", "o_c", "y_true"), got):
                    np.testing.assert_array_equal(v, g[name], err_msg=f"{key}/{name}")
                checked += 1
    assert checked >= 50
    with pytest.raises(ValueError, match="Invalid mode"):
        D.pad_profile(5, 10, "nope", True)


def test_device_batches_match_reference_fixtures_emulated(emu_backend):
    S.check_batches(emu_backend)


def test_device_loader_drops_into_evaluate_emulated(emu_backend):
    S.check_loader_in_evaluate(emu_backend)


def test_packed_eval_batch_roundtrip(emu_backend):
    S.check_packed_eval_batch(emu_backend)
