"""tcgen05 fused inference kernel (csrc/fused_eval_tc.cuh) on a B200: every intermediate
activation of the first user tile against a float64 restatement, and the scores against the
per-op kernels and the FFMA fused kernel."""
import dataclasses

import numpy as np
import pytest
import torch

import tc_stages
from helpers import FP32_RTOL, rel_err, topk_equal_up_to_ties

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,decoder,all_valid", [("tiny", "ca", False), ("tiny", "dot", True),
                                                     ("beauty", "ca", True)])
def test_every_stage_matches_float64_restatement(shape, decoder, all_valid):
    errs, score_err, timed_out = tc_stages.stage_errors(shape, decoder, B=5, all_valid=all_valid)
    assert not timed_out
    for stage, e in errs.items():
        assert e < 2e-5, (stage, e)            # 3xTF32 products + fp32 accumulation: fp32 grade
    assert score_err < FP32_RTOL


@pytest.mark.parametrize("H", [2, 4])
@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("B,T,L,all_valid", [(64, 101, 50, False), (33, 101, 50, True), (7, 333, 50, False),
                                             (9, 101, 64, True), (3, 40, 5, True), (1, 1, 1, True)])
def test_tc_kernel_vs_per_op_and_ffma_kernels(H, decoder, B, T, L, all_valid):
    from carca_replication_b200 import fused, synth

    dev = "cuda"
    shape = dataclasses.replace(synth.BEAUTY, n_heads=H, n_targets=T, seq_len=L, n_items=5000, n_attrs=300)
    model = synth.build_model(shape, decoder, p=0.5, seed=3).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=3).to(dev))
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, B, seed=3, all_valid=all_valid).items()}
    prof, tgt = (b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])]
    with torch.no_grad():
        y_tc = fused.forward(model, prof, tgt, variant=2)
        model.use_fused_eval = False
        y_mod = model.forward(prof, tgt)
        model.use_fused_eval = True
        y_ff = fused.forward(model, prof, tgt, variant=1) if L <= fused.MAX_L else None
    assert not fused.mma_timed_out(model)
    assert tuple(y_tc.shape) == (B, T)
    assert rel_err(y_tc.cpu().numpy(), y_mod.cpu().numpy()) < FP32_RTOL
    assert topk_equal_up_to_ties(y_tc.cpu().numpy(), y_mod.cpu().numpy(), 10, tol=1e-6)
    if y_ff is not None:
        assert rel_err(y_tc.cpu().numpy(), y_ff.cpu().numpy()) < FP32_RTOL
    # every decoder of the kernel (variant 2 picks one by mode): 3 = tcgen05 score MMAs, 4 = fp32 loop with one row
    # per thread, 5 = fp32 loop over candidate pairs, 6 = split decoder kernel; and the second context path (one context row per user as an
    # expanded view)
    o_cu = b["o_c"][:, :1, :].contiguous().expand(-1, T, -1)
    with torch.no_grad():
        y_u2 = fused.forward(model, prof, [(b["o_x"], None, o_cu)], variant=2)
        for v in (3, 4, 5, 6):
            y_v = fused.forward(model, prof, tgt, variant=v)
            y_uv = fused.forward(model, prof, [(b["o_x"], None, o_cu)], variant=v)
            assert rel_err(y_v.cpu().numpy(), y_mod.cpu().numpy()) < FP32_RTOL, v
            assert rel_err(y_u2.cpu().numpy(), y_uv.cpu().numpy()) < FP32_RTOL, v
            assert topk_equal_up_to_ties(y_u2.cpu().numpy(), y_uv.cpu().numpy(), 10, tol=1e-6), v
    assert not fused.mma_timed_out(model)


def test_tc_kernel_is_the_default_and_three_launches():
    from carca_replication_b200 import _native as N
    from carca_replication_b200 import fused, synth

    dev = "cuda"
    shape = synth.TINY
    model = synth.build_model(shape, "ca", seed=2).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=2).to(dev))
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, 6, seed=2).items()}
    prof, tgt = (b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])]
    with torch.no_grad():
        y0 = model.forward(prof, tgt)
        n0 = N.lib().carca_launch_count()
        y1 = model.forward(prof, tgt)
        assert N.lib().carca_launch_count() - n0 == 3      # row packing + the fused forward (fp32-decoder kernel and
                                                           # tcgen05-decoder kernel: the one not for this batch returns at once)
        y2 = fused.forward(model, prof, tgt, variant=2)
    assert torch.equal(y0, y1) and torch.equal(y1, y2)       # variant 0 picks the tensor-core kernel


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("L", [100, 200])
def test_long_windows_one_kernel_forward_equals_rows_pipeline_and_per_op(decoder, L):
    """maxlen 100 / 200 (BASELINE configs[4]).  The default fp32 path for windows longer than one 64-row bin is the
    packed-rows pipeline (no per-user limit, nothing decided on the host); when every user of the batch does fit a
    bin, the one-kernel tensor-core forward (forced here) gives the same scores, and so do the per-op kernels.  A
    user with every position valid runs through the SAME default path."""
    from carca_replication_b200 import fused, synth

    dev = "cuda"
    shape = dataclasses.replace(synth.BEAUTY, seq_len=L, n_items=4000, n_attrs=300)
    model = synth.build_model(shape, decoder, p=0.5, seed=6).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=6).to(dev))
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, 40, seed=6).items()}
    n_valid = (b["p_x"] != 0).sum(1)
    keep = n_valid <= 64                                   # users that fit a bin
    assert int(keep.sum()) >= 30
    short = {k: v[keep].contiguous() for k, v in b.items()}
    prof, tgt = (short["p_x"], None, short["p_c"]), [(short["o_x"], None, short["o_c"])]
    with torch.no_grad():
        assert model._fused_eval_mode(prof, tgt) == "rows_fp32"
        y = model.forward(prof, tgt)
        model.force_eval_path = "tc"
        y_tc = model.forward(prof, tgt)
        model.force_eval_path = None
        model.use_fused_eval = False
        y_mod = model.forward(prof, tgt)
        model.use_fused_eval = True
    assert not fused.mma_timed_out(model)
    for other in (y_tc, y_mod):
        assert rel_err(y.cpu().numpy(), other.cpu().numpy()) < FP32_RTOL
        assert topk_equal_up_to_ties(y.cpu().numpy(), other.cpu().numpy(), 10, tol=1e-6)
    full = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, 3, seed=7, all_valid=True).items()}
    profl, tgtl = (full["p_x"], None, full["p_c"]), [(full["o_x"], None, full["o_c"])]
    with torch.no_grad():
        assert model._fused_eval_mode(profl, tgtl) == "rows_fp32"
        yl = model.forward(profl, tgtl)
        model.use_fused_eval = False
        yl_mod = model.forward(profl, tgtl)
        model.use_fused_eval = True
    assert rel_err(yl.cpu().numpy(), yl_mod.cpu().numpy()) < FP32_RTOL
