"""Men-shaped configuration (BASELINE configs[2]: d = 256, 4 heads, dense 512-d attributes) on a B200: eval through
the packed-rows pipeline in fp32 and bf16 against the CPU oracle, and one train step (loss + every gradient)."""
import dataclasses

import numpy as np
import pytest
import torch

import parity_suite as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _men(**kw):
    from carca_replication_b200 import synth

    return dataclasses.replace(synth.MEN, n_items=2500, n_attrs=96, **kw)


@pytest.mark.parametrize("all_valid", [False, True])
def test_men_shape_fp32_rows_vs_oracle(all_valid):
    from carca_replication_b200 import synth

    shape, B = _men(), 10 if all_valid else 40
    batch = synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid)
    model, y_ref = S.oracle_scores(shape, "ca", batch, seed=5)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=5, no_sync=True)
    with torch.no_grad():
        assert model._fused_eval_mode((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])]) == "rows_fp32"
    S.assert_fp32_parity(y, y_ref, d, B)


def test_men_shape_dot_decoder_fp32_rows_vs_oracle():
    """The dot decoder at d = 256 produces logits of magnitude ~70: the fp32 contract is checked on the logits
    (1e-4 of their scale; a probability of 1e-30 cannot carry a relative 1e-4)."""
    from carca_replication_b200 import synth

    shape, B = _men(), 40
    batch = synth.make_eval_batch(shape, B, seed=5)
    model, y_ref = S.oracle_scores(shape, "dot", batch, seed=5)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=5)
    _, dl, top = S.bf16_errors(y.cpu().numpy(), y_ref)
    assert dl < 1e-4 and top == 1.0


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("all_valid", [False, True])
def test_men_shape_bf16_vs_oracle(decoder, all_valid):
    from carca_replication_b200 import synth

    shape, B = _men(), 10 if all_valid else 64
    batch = synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid)
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=5)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=5, dtype="bf16", expand_ctx=True)
    print("bf16 errors (max |dp|, scaled logit error, top-10 overlap):", S.assert_bf16_parity(y, y_ref, d, B, decoder))


def test_men_shape_train_step_vs_oracle():
    """fwd + BCE + bwd at d = 256 (per-op kernels with the 3xTF32 tcgen05 GEMM): loss and every parameter gradient
    against the oracle's autograd, dropout on with the shared Philox stream."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import ops, synth
    from oracle import carca_oracle as O

    shape = _men(seq_len=20)
    model = synth.build_model(shape, "ca", p=0.25, seed=7)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=7)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder="ca", p_drop=0.25, seed=99)
    bt = synth.make_train_batch(shape, 12, seed=4)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss_ref = O.train_batch_loss(sdo, cfg, (bt["p_x"], table.gather_dense(bt["p_x"]), bt["p_c"], bt["o_x"],
                                             table.gather_dense(bt["o_x"]), bt["o_c"], bt["y_true"]))
    loss_ref.backward()
    model = model.to(DEV).train()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=7).to(DEV))
    ops.set_dropout_seed(99)
    try:
        L = shape.seq_len
        o_x, o_c = bt["o_x"].to(DEV), bt["o_c"].to(DEV)
        y = model.forward((bt["p_x"].to(DEV), None, bt["p_c"].to(DEV)),
                          [(o_x[:, :L], None, o_c[:, :L]), (o_x[:, L:], None, o_c[:, L:])])
        loss = cb.BinaryCrossEntropy().forward(y, bt["y_true"].to(DEV), cb.get_mask(o_x))
        loss.backward()
    finally:
        ops.set_dropout_seed(None)
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    for k, prm in model.named_parameters():
        g_ref = sdo[k].grad.numpy()
        err = np.abs(prm.grad.cpu().numpy() - g_ref).max() / max(np.abs(g_ref).max(), 1e-3)
        assert err < 1e-3, (k, err)
