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


def _fwd(model, d):
    with torch.no_grad():
        y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, d["o_c"])])
    torch.cuda.synchronize()
    return y.clone()


@pytest.mark.parametrize("case", ["beauty_sparse", "beauty_all_valid", "beauty_L129", "beauty_L130_fallback", "men_all_valid",
                                  "men_h8", "beauty_h1", "beauty_L10_empty_profiles", "single_user"])
def test_tensor_core_attention_matches_cuda_core_attention(case, monkeypatch):
    """rows_attn_tc_kernel (tcgen05 QK^T / PV with the key window of a 128-row tile, L <= 129) against the CUDA-core
    attention kernel it replaces (CARCA_ROWS_ATTN_FFMA=1) on the same batch: both are bf16-operand / fp32-softmax
    computations of src/carca.py:246-256, so they agree to bf16 rounding.  Run twice with the pipeline's scratch
    filled with NaN patterns in between: no result may depend on memory the step did not write."""
    from carca_replication_b200 import fused, synth

    men = dataclasses.replace(synth.MEN, n_items=3000, n_attrs=96)
    shape, B, all_valid = {
        "beauty_sparse": (_beauty(), 700, False), "beauty_all_valid": (_beauty(), 333, True),
        "beauty_L129": (_beauty(129), 150, True), "beauty_L130_fallback": (_beauty(130), 40, True),
        "men_all_valid": (men, 100, True), "men_h8": (dataclasses.replace(men, n_heads=8), 300, False),
        "beauty_h1": (dataclasses.replace(_beauty(), n_heads=1), 400, False),
        "beauty_L10_empty_profiles": (_beauty(10), 300, False), "single_user": (_beauty(), 1, False)}[case]
    model = synth.build_model(shape, "ca", p=0.5, seed=5).to(DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=5).to(DEV))
    d = {k: v.to(DEV) for k, v in synth.make_eval_batch(shape, B, seed=5, all_valid=all_valid).items()}
    if case == "beauty_L10_empty_profiles":
        d["p_x"][::7] = 0                                # users without a single valid position (only the L-1 pad row)
        d["p_x"][3, :-1] = 0                             # and one with exactly one
    monkeypatch.setenv("CARCA_ROWS_ATTN_FFMA", "1")
    y_cc = _fwd(model, d)
    monkeypatch.delenv("CARCA_ROWS_ATTN_FFMA")
    y_tc = _fwd(model, d)
    for buf in fused._rows_scratch_cache.values():
        buf.fill_(255)
    y_tc2 = _fwd(model, d)
    assert not fused.mma_timed_out(model)
    assert not torch.isnan(y_tc).any() and not torch.isnan(y_tc2).any()
    assert float((y_tc - y_cc).abs().max()) < 5e-3
    assert float((y_tc2 - y_tc).abs().max()) < 2e-3      # (row order of the packing pass differs from run to run)
    if case == "men_h8":       # 8 heads: the CUDA-core decoder runs one warp per head (256 threads), fp32 flavour alike
        model.set_eval_dtype("fp32")
        model.force_eval_path = "rows_fp32"
        assert float((_fwd(model, d) - y_tc).abs().max()) < S.BF16_TOL
    if case == "beauty_L130_fallback":
        assert torch.equal(y_tc, y_cc)                    # window > 256 keys: the CUDA-core kernel serves both calls


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_ffn_chain_kernel_matches_the_separate_gemm_launches(decoder, monkeypatch):
    """rows_ffn_chain_kernel (FFN-1 -> FFN-2 + LayerNorm -> next block's Q / K / V or the decoder's K / V in one kernel,
    intermediates in shared memory) against the same pipeline with one GEMM launch per link (CARCA_ROWS_NO_CHAIN=1):
    the same bf16 roundings at the same places, so the scores agree far inside the bf16 contract."""
    from carca_replication_b200 import fused, synth

    shape = _beauty()
    model = synth.build_model(shape, decoder, p=0.5, seed=9).to(DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=9).to(DEV))
    d = {k: v.to(DEV) for k, v in synth.make_eval_batch(shape, 500, seed=9).items()}
    monkeypatch.setenv("CARCA_ROWS_NO_CHAIN", "1")
    y_sep = _fwd(model, d)
    monkeypatch.delenv("CARCA_ROWS_NO_CHAIN")
    y_chain = _fwd(model, d)
    assert not fused.mma_timed_out(model)
    diff = float((y_chain - y_sep).abs().max())
    print("chain vs separate launches: max |dp|", diff)
    assert diff < 5e-3          # (fp32 summation order differs in the LayerNorm input: an occasional bf16 ulp downstream)


def test_stage_events_time_every_kernel_of_the_bf16_step():
    """carca_rows_set_stage_events (bench.py's per-kernel roofline): the library records the caller's CUDA events between
    its launches; the ids name the stages of the chained d = 64 pipeline and the intervals are positive."""
    import ctypes as C

    from carca_replication_b200 import _native as N
    from carca_replication_b200 import synth

    shape = _beauty()
    model = synth.build_model(shape, "ca", p=0.5, seed=9).to(DEV).eval().set_eval_dtype("bf16")
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=9).to(DEV))
    d = {k: v.to(DEV) for k, v in synth.make_eval_batch(shape, 300, seed=9).items()}
    _fwd(model, d)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(24)]
    for e in evs:
        e.record()
    torch.cuda.synchronize()
    arr = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
    lib = N.lib()
    try:
        lib.carca_rows_set_stage_events(arr, len(evs))
        _fwd(model, d)
        ids = (C.c_int32 * 64)()
        n = lib.carca_rows_stage_ids(ids, 64)
    finally:
        lib.carca_rows_set_stage_events(None, 0)
    nb = shape.n_blocks
    assert list(ids[:n]) == [0, 1, 2, 3] + [4, 7] * nb + [9]
    assert all(evs[i].elapsed_time(evs[i + 1]) > 0 for i in range(n - 1))
    _fwd(model, d)                      # recording is off again: nothing is recorded
    assert lib.carca_rows_stage_ids(ids, 64) == 0
