"""bf16 packed-rows pipeline on a B200 (csrc/rows_bf16.cuh, precision 0; BASELINE configs[1] "bf16/fp32") against the
fp32 CPU oracle under the bf16 contract of tests/parity_suite.py (BF16_TOL = 1e-2)."""
import dataclasses

import numpy as np
import pytest
import torch

import parity_suite as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _beauty(L=50, **kw):
    from carca_replication_b200 import synth

    return dataclasses.replace(synth.BEAUTY, n_items=4000, n_attrs=300, seq_len=L, **kw)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("all_valid", [False, True])
def test_beauty_shape_bf16_vs_oracle(decoder, all_valid):
    from carca_replication_b200 import synth

    shape, B = _beauty(), 24 if all_valid else 160
    batch = synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid)
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=5)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=5, dtype="bf16", expand_ctx=True, no_sync=True)
    print("bf16 errors (max |dp|, scaled logit error, top-10 overlap):", S.assert_bf16_parity(y, y_ref, d, B, decoder))


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("L", [100, 200])
def test_long_windows_bf16_vs_oracle(decoder, L):
    shape, B = _beauty(L), 20
    batch = S.long_window_batch(shape, B, seed=31)
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=31)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=31, dtype="bf16")
    S.assert_bf16_parity(y, y_ref, d, B, decoder)


def test_tiny_shape_single_user_chunked_candidates_bf16():
    from carca_replication_b200 import synth

    shape = dataclasses.replace(synth.TINY, n_targets=300)
    batch = synth.make_eval_batch(shape, 1, seed=2)
    batch["o_x"][0, 7] = 0
    model, y_ref = S.oracle_scores(shape, "ca", batch, seed=2)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=2, dtype="bf16")
    assert tuple(y.shape) == (1, 300)
    S.assert_bf16_parity(y, y_ref, d, 1, "ca")


def test_bf16_follows_weight_updates_and_reports_status():
    """The bf16 plan is rebuilt with the fp32 plan when parameters change (FusedAdam / in-place edits), and
    evaluate() runs through it."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import fused, synth

    shape = synth.TINY
    model = synth.build_model(shape, "ca", p=0.0, seed=3).to(DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=3).to(DEV))
    b = {k: v.to(DEV) for k, v in synth.make_eval_batch(shape, 8, seed=3).items()}
    with torch.no_grad():
        y0 = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
        model.decoder.ffn.bias.add_(1.0)
        y1 = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
    assert float((S._logit(y1.cpu().numpy()) - S._logit(y0.cpu().numpy())).mean()) == pytest.approx(1.0, abs=1e-3)
    loader = [(b["p_x"], None, b["p_c"], b["o_x"], None, b["o_c"], b["y_true"])]
    hr, ndcg, loss = cb.evaluate(model, loader, DEV, 10)
    assert 0.0 <= hr <= 1.0 and np.isfinite(loss)
    assert not fused.mma_timed_out(model)


@pytest.mark.parametrize("name", ["beauty_ca", "beauty_dot", "learnable_ca", "sinus_dot", "single_user_ca"])
def test_reference_fixtures_in_bf16(name):
    """Golden fixtures of the real reference (learnable / sinusoidal positions, B = 1, both decoders) through the bf16
    pipeline: scores within the bf16 contract of the reference's fp32 output."""
    import carca_replication_b200 as cb
    from helpers import load_case

    cfg, sd, z = load_case(name)
    model = S.build_model(cfg, sd, DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(cb.ItemAttrTable.from_dense(z["attr_table"], sparse=True).to(DEV))
    p_x, _, p_c, o_x, _, o_c, y_true = [t.to(DEV) for t in S.batch_of(z, "eval")]
    with torch.no_grad():
        assert model._fused_eval_mode((p_x, None, p_c), [(o_x, None, o_c)]) == "rows_bf16"
        y = model.forward((p_x, None, p_c), [(o_x, None, o_c)])
    dp, dl, top = S.bf16_errors(y.cpu().numpy(), z["eval/y_pred"])
    assert dl < S.BF16_TOL and top >= 0.9
    if cfg["decoder"] == "ca":
        assert dp < S.BF16_TOL


def test_bf16_needs_a_supported_shape():
    """eval_dtype == "bf16" on a shape the pipeline does not cover fails loudly instead of silently running fp32."""
    from helpers import load_case

    cfg, sd, z = load_case("men_dot_d128")
    import carca_replication_b200 as cb

    model = S.build_model(cfg, sd, DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(cb.ItemAttrTable.from_dense(z["attr_table"], sparse=False).to(DEV))
    p_x, _, p_c, o_x, _, o_c, _ = [t.to(DEV) for t in S.batch_of(z, "eval")]
    with torch.no_grad(), pytest.raises(RuntimeError, match="bf16"):
        model.forward((p_x, None, p_c), [(o_x, None, o_c)])
