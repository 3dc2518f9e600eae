"""Device-side batch construction on a B200 against the real reference loader's fixtures."""
import pytest

import data_suite as S

pytestmark = pytest.mark.gpu


def test_device_batches_match_reference_fixtures():
    S.check_batches("cuda")


def test_sampled_negatives_are_uniform_over_free_items():
    S.check_negatives_are_uniform("cuda")


def test_packed_eval_batch_roundtrip():
    S.check_packed_eval_batch("cuda")


def test_device_loader_drops_into_evaluate():
    S.check_loader_in_evaluate("cuda")


def test_loader_covers_every_user_of_a_beauty_sized_log():
    """Beauty-sized log through DeviceLoader in 8192-user eval batches: every user appears exactly once, windows are
    left-padded and end in a valid item.  (Throughput is reported by bench.py's `device_pipeline` section — there is
    no wall-clock threshold in the parity gate.)"""
    import numpy as np
    import torch

    from carca_replication_b200.device_data import DeviceInteractions, DeviceLoader

    rng = np.random.default_rng(0)
    U, n_items, C = 52204, 57290, 6
    lens = np.clip(np.rint(rng.lognormal(1.8, 0.7, size=U)), 4, 300).astype(np.int64)
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    items = rng.integers(1, n_items, size=int(rowptr[-1])).astype(np.int32)
    ctx = rng.random((int(rowptr[-1]), C), dtype=np.float32)
    log = DeviceInteractions(torch.from_numpy(rowptr), torch.from_numpy(items), torch.from_numpy(ctx)).to("cuda")
    loader = DeviceLoader(log, n_items, 50, 100, "test", batch_size=8192)
    n = 0
    for b in loader:
        p_x, o_x = b[0], b[3]
        n += p_x.shape[0]
        assert bool((p_x[:, -1] != 0).all())                       # left padding: the window ends in a real item
        assert bool((o_x != 0).all()) and o_x.shape[1] == 101        # 1 positive + 100 sampled negatives
        pad = (p_x == 0)
        assert bool((pad[:, 1:] <= pad[:, :-1]).all())               # padding only on the left
    assert n == U
