"""CUDA-graph train step vs the eager step on a B200 (same weights, same batches)."""
import dataclasses

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_graphed_step_matches_eager_without_dropout(decoder):
    import carca_replication_b200 as cb
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep

    dev = "cuda"
    shape = synth.TINY
    table = synth.make_attr_table(shape, seed=3).to(dev)
    batches = [{k: v.to(dev) for k, v in synth.make_train_batch(shape, 32, seed=50 + i).items()} for i in range(4)]

    def fresh():
        m = synth.build_model(shape, decoder, p=0.0, seed=3).to(dev).train()
        m.embeds.set_attr_table(table)
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, betas=(0.9, 0.98), capturable=True)

    m1, o1 = fresh()
    sd0 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    step = GraphedTrainStep(m1, o1, batches[0], warmup=2)
    # the warm-up / capture steps moved the weights: restart both sides from the same state
    m1.load_state_dict(sd0)
    for st in o1.state.values():          # the captured kernels reference THESE tensors: reset them in place
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    m2, o2 = fresh()
    m2.load_state_dict(sd0)
    L = shape.seq_len
    loss_fn = cb.BinaryCrossEntropy()
    for b in batches:
        l1 = step(b).clone()
        o2.zero_grad()
        y = m2.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]),
                                                    (b["o_x"][:, L:], None, b["o_c"][:, L:])])
        l2 = loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
        l2.backward()
        o2.step()
        assert abs(l1.item() - l2.item()) < 1e-5 * max(1.0, abs(l2.item()))
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if k.endswith("WK.bias"):
            continue      # its gradient is pure summation noise (softmax shift invariance), which Adam turns into +-lr steps
        assert torch.allclose(p1, p2, rtol=1e-3, atol=2e-5), k


def test_graphed_step_draws_new_dropout_masks_every_replay():
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep

    dev = "cuda"
    shape = synth.TINY
    m = synth.build_model(shape, "ca", p=0.5, seed=3).to(dev).train()
    m.embeds.set_attr_table(synth.make_attr_table(shape, seed=3).to(dev))
    opt = torch.optim.Adam(m.parameters(), lr=0.0, capturable=True)      # lr 0: weights stay put
    b = {k: v.to(dev) for k, v in synth.make_train_batch(shape, 32, seed=9).items()}
    step = GraphedTrainStep(m, opt, b)
    losses = {round(step(b).item(), 7) for _ in range(6)}
    assert len(losses) >= 5          # same data, same weights: only the masks differ


def test_graphed_step_with_long_windows_and_fused_adam():
    """maxlen 100: the fused kernels run inside the graph while every user's active positions fit a 64-row bin; a
    batch with a longer user takes the eager per-op step.  Both must track the plain eager loop (FusedAdam on both)."""
    import dataclasses

    import carca_replication_b200 as cb
    from carca_replication_b200 import _native as N
    from carca_replication_b200 import synth
    from carca_replication_b200.graph import GraphedTrainStep

    dev = "cuda"
    shape = dataclasses.replace(synth.TINY, seq_len=100)
    L = shape.seq_len
    table = synth.make_attr_table(shape, seed=3).to(dev)
    batches = []
    for i in range(4):
        b = synth.make_train_batch(shape, 16, seed=80 + i)
        cut = L - 40
        if i != 2:                                  # batch 2 keeps its long users (> 64 active positions)
            b["p_x"][:, :cut] = 0
            b["o_x"][:, :cut] = 0
            b["o_x"][:, L:L + cut] = 0
        else:
            b["p_x"][0, :] = 5
            b["o_x"][0, :] = 6
        batches.append({k: v.to(dev) for k, v in b.items()})

    def fresh():
        m = synth.build_model(shape, "ca", p=0.0, seed=3).to(dev).train()
        m.embeds.set_attr_table(table)
        return m, cb.FusedAdam(m.parameters(), lr=1e-3, betas=(0.9, 0.98))

    m1, o1 = fresh()
    sd0 = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    step = GraphedTrainStep(m1, o1, batches[0], warmup=2)
    assert step.long_windows and step.graph_is_fused
    m1.load_state_dict(sd0)
    for st in o1.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    m2, o2 = fresh()
    m2.load_state_dict(sd0)
    loss_fn = cb.BinaryCrossEntropy()
    for i, b in enumerate(batches):
        n0 = N.lib().carca_launch_count()
        l1 = step(b).clone()
        eager_fallback = N.lib().carca_launch_count() - n0 > 50      # a replay launches nothing through the C ABI
        assert eager_fallback == (i == 2)
        o2.zero_grad()
        y = m2.forward((b["p_x"], None, b["p_c"]), [(b["o_x"][:, :L], None, b["o_c"][:, :L]),
                                                    (b["o_x"][:, L:], None, b["o_c"][:, L:])])
        l2 = loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
        l2.backward()
        o2.step()
        assert abs(l1.item() - l2.item()) < 1e-5 * max(1.0, abs(l2.item()))
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if k.endswith("WK.bias"):
            continue
        assert torch.allclose(p1, p2, rtol=1e-3, atol=2e-5), k


@pytest.mark.parametrize("decoder", ["ca", "dot"])
def test_graphed_eval_step_matches_eager_evaluate_body(decoder):
    """GraphedEvalStep (forward + BCE + HR@10 / NDCG@10 accumulation as one CUDA graph) against the same calls
    issued eagerly, over several batches; also the static-input form with an expanded per-user context view and a
    pinned host result."""
    import carca_replication_b200 as cb
    from carca_replication_b200 import ops, synth
    from carca_replication_b200.graph import GraphedEvalStep

    dev = "cuda"
    shape = synth.TINY
    model = synth.build_model(shape, decoder, seed=4).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=4).to(dev))
    batches = [{k: v.to(dev) for k, v in synth.make_eval_batch(shape, 48, seed=70 + i).items()} for i in range(4)]
    loss_fn = cb.BinaryCrossEntropy()
    ref = torch.zeros(4, dtype=torch.float64, device=dev)
    with torch.no_grad():
        for b in batches:
            y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
            ref[3] += loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
            ops.rank_metrics_(ref[:3], y, b["y_true"], 10)
    step = GraphedEvalStep(model, batches[0], k=10)
    assert float(step.stats.abs().sum()) == 0.0          # warm-up and capture runs are not counted
    for b in batches:
        step(b)
    assert torch.allclose(step.stats, ref, rtol=1e-5, atol=1e-9)     # (loss: one-launch reduction vs three kernels)
    assert float(ref[2]) == 4 * 48

    # static inputs (the caller refills the buffers), one context row per user as an expanded view, host result
    T = batches[0]["o_x"].shape[1]
    bufs = {k: v.clone() for k, v in batches[0].items()}
    bufs["o_c"] = bufs["o_c"][:, :1, :].contiguous()
    host = torch.zeros(4, dtype=torch.float64).pin_memory()
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    step2 = GraphedEvalStep(model, dict(bufs, o_c=bufs["o_c"].expand(-1, T, -1)), k=10, stats=stats, result=host,
                            static_inputs=True)
    ref2 = torch.zeros(4, dtype=torch.float64, device=dev)
    with torch.no_grad():
        for b in batches:
            for k in ("p_x", "p_c", "o_x", "y_true"):
                bufs[k].copy_(b[k])
            bufs["o_c"].copy_(b["o_c"][:, :1, :])
            step2.replay()
            oc = b["o_c"][:, :1, :].expand(-1, T, -1)
            y = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, oc)])
            ref2[3] += loss_fn.forward(y, b["y_true"], cb.get_mask(b["o_x"]))
            ops.rank_metrics_(ref2[:3], y, b["y_true"], 10)
    torch.cuda.synchronize()
    assert torch.allclose(stats, ref2, rtol=1e-5, atol=1e-9)
    assert torch.allclose(host, ref2.cpu(), rtol=1e-5, atol=1e-9)
    with pytest.raises(ValueError):
        GraphedEvalStep(model.train(), batches[0])


def test_graphed_eval_step_long_windows_replays_every_batch():
    """maxlen 100: the captured step runs the packed-rows pipeline, which has no per-user row limit — a batch whose
    users have every position valid replays through the SAME graph (nothing is checked or re-routed on the host)."""
    import dataclasses

    import carca_replication_b200 as cb
    from carca_replication_b200 import ops, synth
    from carca_replication_b200.graph import GraphedEvalStep

    dev = "cuda"
    shape = dataclasses.replace(synth.BEAUTY, seq_len=100, n_items=4000, n_attrs=300)
    model = synth.build_model(shape, "ca", seed=6).to(dev).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=6).to(dev))
    b = {k: v.to(dev) for k, v in synth.make_eval_batch(shape, 40, seed=6).items()}
    keep = (b["p_x"] != 0).sum(1) <= 60
    short = {k: v[keep][:24].contiguous() for k, v in b.items()}
    full = {k: v[:24].to(dev) for k, v in synth.make_eval_batch(shape, 24, seed=7, all_valid=True).items()}
    step = GraphedEvalStep(model, short, k=10)
    assert step.graph_is_fused and not step.long_windows
    ref = torch.zeros(4, dtype=torch.float64, device=dev)
    with torch.no_grad():
        for batch in (short, full, short):
            step(batch)
            y = model.forward((batch["p_x"], None, batch["p_c"]), [(batch["o_x"], None, batch["o_c"])])
            ops.eval_metrics_(ref, y, batch["y_true"], batch["o_x"], 10)
    assert torch.allclose(step.stats, ref, rtol=1e-5, atol=1e-9)
    assert float(ref[2]) == 72


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_evaluate_replays_repeated_batch_shapes_as_graphs(dtype):
    """evaluate() (src/train.py:35-53) replays a batch shape's body as a CUDA graph from its third occurrence on
    (train._graphed_eval_batch): same (HR, NDCG, loss) as the eager loop, also after a weight update between two calls
    and for a trailing batch of another shape."""
    import copy

    import carca_replication_b200 as cb
    from carca_replication_b200 import synth, train

    shape = dataclasses.replace(synth.BEAUTY, n_items=3000, n_attrs=200)
    model = synth.build_model(shape, "ca", p=0.5, seed=21).to(DEV).eval().set_eval_dtype(dtype)
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=21).to(DEV))

    def loader():
        out = []
        for i, B in enumerate((96, 96, 96, 96, 40)):
            b = {k: v.to(DEV) for k, v in synth.make_eval_batch(shape, B, seed=50 + i).items()}
            out.append((b["p_x"], None, b["p_c"], b["o_x"], None, b["o_c"], b["y_true"]))
        return out

    batches = loader()
    train.USE_EVAL_GRAPHS = False
    try:
        want = cb.evaluate(model, batches, DEV, 10)
    finally:
        train.USE_EVAL_GRAPHS = True
    got = cb.evaluate(model, batches, DEV, 10)         # the 96-user shape is captured at its third batch
    assert sum(1 for v in model._eval_graph_steps.values() if not isinstance(v, (int, bool))) == 1
    assert got[0] == want[0] and got[1] == pytest.approx(want[1], rel=1e-6) and got[2] == pytest.approx(want[2], rel=1e-5)
    cb.evaluate(model, batches, DEV, 10)
    again = cb.evaluate(model, batches, DEV, 10)       # third call: the trailing 40-user batch is captured too
    assert sum(1 for v in model._eval_graph_steps.values() if not isinstance(v, (int, bool))) == 2
    assert again[0] == want[0] and again[2] == pytest.approx(want[2], rel=1e-5)
    with torch.no_grad():
        model.decoder.ffn.bias.add_(0.5)               # weights change between validation passes
    train.USE_EVAL_GRAPHS = False
    try:
        want2 = cb.evaluate(model, batches, DEV, 10)
    finally:
        train.USE_EVAL_GRAPHS = True
    got2 = cb.evaluate(model, batches, DEV, 10)
    assert want2[2] != pytest.approx(want[2], rel=1e-3)
    assert got2[0] == want2[0] and got2[2] == pytest.approx(want2[2], rel=1e-5)
    clone = copy.deepcopy(model)                       # captured graphs are not part of the module's state
    assert "_eval_graph_steps" not in clone.__dict__
    # Module.to() re-wraps the parameters even when nothing moves: the plans are rebuilt into new buffers, so the
    # captured graphs must be dropped (never replayed against freed memory) and the result stays the eager one
    model.to(DEV)
    got3 = cb.evaluate(model, batches, DEV, 10)
    assert got3[0] == want2[0] and got3[2] == pytest.approx(want2[2], rel=1e-5)
    from carca_replication_b200 import fused

    assert not fused.mma_timed_out(model)
