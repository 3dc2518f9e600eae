"""Device-side batch construction on a B200 against the real reference loader's fixtures."""
import pytest

import data_suite as S

pytestmark = pytest.mark.gpu


def test_device_batches_match_reference_fixtures():
    S.check_batches("cuda")


def test_sampled_negatives_are_uniform_over_free_items():
    S.check_negatives_are_uniform("cuda")


def test_device_loader_drops_into_evaluate():
    S.check_loader_in_evaluate("cuda")


def test_loader_throughput_smoke():
    """Beauty-sized log: one 8192-user eval batch is built in well under a millisecond-scale budget."""
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
    it = iter(loader)
    next(it)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 0
    for b in it:
        n += b[0].shape[0]
    e1.record()
    torch.cuda.synchronize()
    users_per_s = n / (e0.elapsed_time(e1) * 1e-3)
    print(f"device loader: {users_per_s:.0f} users/s")
    assert users_per_s > 1e6          # the reference loader builds ~500 users/s on the host
