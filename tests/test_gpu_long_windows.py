"""fp32 packed-rows pipeline on a B200 (csrc/rows_bf16.cuh, precision 1): windows longer than one 64-row bin
(BASELINE configs[4]: maxlen 100 / 200), users with more than 64 valid positions, widths outside the fused kernel —
against the CPU oracle under the fp32 contract (scores 1e-4 rel, top-10 identical up to ties, HR@10 equal, NDCG@10 to
3 dp), with no host sync on the path."""
import dataclasses

import numpy as np
import pytest
import torch

import parity_suite as S
from helpers import FP32_RTOL, MODEL_CASES, load_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _beauty(L, **kw):
    from carca_replication_b200 import synth

    return dataclasses.replace(synth.BEAUTY, n_items=4000, n_attrs=300, seq_len=L, **kw)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
@pytest.mark.parametrize("L", [100, 200])
def test_long_windows_with_users_beyond_one_bin_vs_oracle(decoder, L):
    shape, B = _beauty(L), 20
    batch = S.long_window_batch(shape, B, seed=31)
    assert int((batch["p_x"] != 0).sum(1).max()) == L            # a user with every position valid
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=31)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=31, no_sync=True)
    with torch.no_grad():
        assert model._fused_eval_mode((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])]) == "rows_fp32"
    S.assert_fp32_parity(y, y_ref, d, B)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_all_valid_maxlen_100_vs_oracle(decoder):
    from carca_replication_b200 import synth

    shape, B = _beauty(100), 6
    batch = synth.make_eval_batch(shape, B, seed=7, all_valid=True)
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=7)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=7, expand_ctx=True)
    S.assert_fp32_parity(y, y_ref, d, B)


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_rows_path_equals_one_kernel_path_on_short_windows(decoder):
    """L = 50: the packed-rows pipeline and the one-kernel tensor-core forward agree (both within the fp32 contract)."""
    from carca_replication_b200 import synth

    shape, B = _beauty(50), 40
    batch = synth.make_eval_batch(shape, B, seed=3)
    model, y_ref = S.oracle_scores(shape, decoder, batch, seed=3)
    model, y_rows, d = S.run_eval_path(model, shape, batch, DEV, seed=3, path="rows_fp32")
    S.assert_fp32_parity(y_rows, y_ref, d, B)
    model.force_eval_path = None
    with torch.no_grad():
        y_tc = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])])
    assert rel_err(y_rows.cpu().numpy(), y_tc.cpu().numpy()) < FP32_RTOL


ROWS_FIXTURES = ["beauty_ca", "beauty_dot", "men_ca", "men_dot_d128", "noresid_dot", "noresid_ca", "learnable_ca",
                 "sinus_dot", "single_user_ca"]


@pytest.mark.parametrize("name", ROWS_FIXTURES)
def test_reference_fixtures_through_the_rows_pipeline(name):
    """Golden fixtures of the real reference (d = 32 / 64 / 128, 1 / 2 / 4 heads, learnable and sinusoidal positions,
    no-residual blocks, B = 1) with the fp32 rows pipeline forced."""
    assert name in MODEL_CASES
    cfg, sd, z = load_case(name)
    model = S.build_model(cfg, sd, DEV).eval()
    import carca_replication_b200 as cb

    model.embeds.set_attr_table(cb.ItemAttrTable.from_dense(z["attr_table"], sparse=True).to(DEV))
    model.force_eval_path = "rows_fp32"
    p_x, _, p_c, o_x, _, o_c, y_true = [t.to(DEV) for t in S.batch_of(z, "eval")]
    with torch.no_grad():
        y = model.forward((p_x, None, p_c), [(o_x, None, o_c)])
    ref = z["eval/y_pred"]
    assert rel_err(y.cpu().numpy(), ref) < FP32_RTOL
    assert cb.compute_HR(y, y_true, cfg["k"]) == float(z["eval/HR"])
    assert round(cb.compute_NDCG(y, y_true, cfg["k"]), 3) == round(float(z["eval/NDCG"]), 3)


def test_candidate_chunks_per_candidate_context_and_single_user():
    """T = 333 candidates (three 128-candidate slices), a distinct context row per candidate, B = 1."""
    from carca_replication_b200 import synth

    shape = _beauty(100, n_targets=333)
    batch = S.long_window_batch(shape, 1, seed=11, n_long=1)
    g = torch.Generator().manual_seed(5)
    batch["o_c"] = torch.rand(batch["o_c"].shape, generator=g)
    batch["o_x"][0, 5] = 0                                         # a padding candidate scores sigmoid(ffn bias)
    model, y_ref = S.oracle_scores(shape, "ca", batch, seed=11)
    model, y, d = S.run_eval_path(model, shape, batch, DEV, seed=11)
    assert tuple(y.shape) == (1, 333)
    assert rel_err(y.cpu().numpy(), y_ref) < FP32_RTOL


def test_graphed_eval_step_on_long_windows():
    """evaluate()'s per-batch body as a CUDA graph at maxlen 100 with users beyond one bin: replay == eager, and the
    replay follows weight updates (plan refresh into the captured buffers)."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import ops, synth
    from carca_replication_b200.graph import GraphedEvalStep

    shape, B = _beauty(100), 24
    batches = [S.long_window_batch(shape, B, seed=40 + i) for i in range(2)]
    model = synth.build_model(shape, "ca", p=0.5, seed=9).to(DEV).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=9).to(DEV))
    dev_b = [{k: v.to(DEV) for k, v in b.items()} for b in batches]
    stats = torch.zeros(4, dtype=torch.float64, device=DEV)
    step = GraphedEvalStep(model, dev_b[0], k=10, stats=stats)
    for b in dev_b:
        step(b)
    ref = torch.zeros(4, dtype=torch.float64, device=DEV)
    with torch.no_grad():
        for b in dev_b:
            y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
            ops.eval_metrics_(ref, y, b["y_true"], b["o_x"], 10)
    np.testing.assert_allclose(stats.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6)
    with torch.no_grad():
        model.norm.bias.add_(0.25)                                 # weights change: the replay must see them
    stats.zero_()
    ref.zero_()
    step(dev_b[0])
    with torch.no_grad():
        y = model.forward((dev_b[0]["p_x"], None, dev_b[0]["p_c"]), [(dev_b[0]["o_x"], None, dev_b[0]["o_c"])])
        ops.eval_metrics_(ref, y, dev_b[0]["y_true"], dev_b[0]["o_x"], 10)
    np.testing.assert_allclose(stats.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6)
