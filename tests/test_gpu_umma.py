"""Known-answer tests of the tcgen05 / TMEM primitives (csrc/umma.cuh) on a B200."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SPLIT, A_TMEM = 1, 2     # mode 4 (MN-major no-swizzle tf32 B operand) reads as zeros on B200: unsupported, unused


def _run(A, B, mode):
    from carca_replication_b200 import _native as N

    dev = torch.device("cuda")
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    n, k = B.shape
    c = torch.full((128, n), float("nan"), dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    N.call("carca_umma_selftest", c.data_ptr(), a.data_ptr(), b.data_ptr(), n, k, int(mode), status.data_ptr(),
           N.stream())
    torch.cuda.synchronize()
    assert int(status.item()) == 0, "mbarrier wait timed out: the MMA never completed"
    return c.cpu().numpy()


@pytest.mark.parametrize("mode", [0, A_TMEM])
@pytest.mark.parametrize("n,k", [(64, 64), (16, 8), (32, 32), (128, 64), (256, 64), (64, 128)])
def test_small_integer_operands_are_exact(n, k, mode):
    rng = np.random.default_rng(n * 1000 + k)
    A = rng.integers(-4, 5, size=(128, k)).astype(np.float32)
    B = rng.integers(-4, 5, size=(n, k)).astype(np.float32)
    C = _run(A, B, mode)
    np.testing.assert_array_equal(C, A @ B.T)


@pytest.mark.parametrize("mode", [0, A_TMEM])
@pytest.mark.parametrize("n,k", [(64, 64), (128, 32), (32, 128)])
def test_3xtf32_split_is_fp32_grade(n, k, mode):
    rng = np.random.default_rng(7 + n + k)
    A = rng.standard_normal((128, k)).astype(np.float32)
    B = rng.standard_normal((n, k)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T      # error scale of a dot product
    e1 = np.max(np.abs(_run(A, B, mode) - ref) / scale)
    e3 = np.max(np.abs(_run(A, B, mode | SPLIT) - ref) / scale)
    assert e3 < 2e-6, e3
    assert 1e-5 < e1 < 2e-3, e1          # single-pass tf32 really truncates (the split is what buys fp32 grade)
