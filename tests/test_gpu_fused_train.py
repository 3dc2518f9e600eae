"""Fused training core (csrc/fused_train.cuh) on a real B200: against the per-op kernels on the same
inputs / weights / Philox seed, at tiny and full Beauty sizes.  The fixture- and oracle-based train tests of
tests/test_gpu_parity.py run through the same fused core (it is CARCA.forward's default in train mode)."""
import pytest

import parity_suite as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("kw", [dict(decoder="ca", p=0.3), dict(decoder="dot", p=0.3), dict(decoder="ca", p=0.0, heads=4),
                                dict(decoder="ca", p=0.25, heads=1, odd_masks=True),
                                dict(decoder="dot", p=0.25, odd_masks=True), dict(decoder="ca", p=0.5, n_tuples=1),
                                dict(decoder="ca", p=0.3, B=9, all_valid=True),
                                dict(decoder="ca", p=0.3, B=300), dict(decoder="dot", p=0.0, B=300, odd_masks=True)])
def test_fused_train_vs_per_op_kernels_tiny(kw):
    S.check_fused_train_vs_per_op(DEV, **kw)


@pytest.mark.parametrize("kw", [dict(decoder="ca", p=0.5, B=256), dict(decoder="dot", p=0.5, B=256),
                                dict(decoder="ca", p=0.2, B=40, all_valid=True),
                                dict(decoder="ca", p=0.5, B=64, odd_masks=True)])
def test_fused_train_vs_per_op_kernels_beauty_shape(kw):
    """Full Beauty sizes (57,290 items, A=6,507 CSR, L=50, d=64, g=256, 3 blocks, reference batch 256)."""
    S.check_fused_train_vs_per_op(DEV, shape_name="beauty", **kw)


def test_train_forward_falls_back_to_per_op_kernels_outside_the_fused_range():
    """d = 256 (Men shape) is outside the fused core's range: CARCA.forward must take the per-op path."""
    import torch

    from carca_replication_b200 import synth

    model = synth.build_model(synth.MEN, "ca", p=0.5).train()
    x = torch.zeros(2, 50, dtype=torch.int32)
    assert not model._fused_train_applies((x, None, None), [(x, None, None), (x, None, None)])
    tiny = synth.build_model(synth.TINY, "ca", p=0.5).train()
    x = torch.zeros(2, 12, dtype=torch.int32)
    assert tiny._fused_train_applies((x, None, None), [(x, None, None), (x, None, None)])
    assert not tiny._fused_train_applies((x, None, None), [(x[:, :5], None, None)])


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_vs_torch(wd):
    S.check_fused_adam(DEV, weight_decay=wd)


def test_fused_train_edge_shapes():
    S.check_fused_train_edges(DEV)
