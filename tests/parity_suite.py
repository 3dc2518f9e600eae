"""Parity checks shared by the GPU tests (tests/test_gpu_*.py, real kernels on a B200) and the
CPU emulator tests (tests/test_emu_*.py, same kernels' logic under tools/emu).  `device` is
"cuda" or "cpu"."""
import io
import json

import numpy as np
import torch
import torch.nn as nn

from helpers import (FP32_RTOL, GOLDEN, batch_of, grad_err, grad_floor, load_case, oracle_cfg, rel_err,
                     topk_equal_up_to_ties)

import carca_replication_b200 as cb
from carca_replication_b200 import ops


def build_model(cfg, sd, device, p=None):
    p = cfg["p"] if p is None else p
    if cfg["encoding"] == "identity":
        enc = cb.IdentityEncoding()
    elif cfg["encoding"] == "learnable":
        enc = cb.LearnableEncoding(cfg["d"], cfg["L"])
    else:
        enc = cb.PositionalEncoding(cfg["d"], cfg["L"])
    kind = cfg.get("embedding", "all")               # the constructor calls of scripts/training.py:78-89
    if kind == "all":
        emb = cb.AllEmbedding(cfg["n_items"], cfg["d"], cfg["g"], cfg["C"], cfg["A"], enc)
    elif kind == "id":
        emb = cb.IdEmbedding(cfg["n_items"], cfg["d"], enc)
    elif kind == "mlpid":
        emb = cb.MLPIdEmbedding(cfg["n_items"], cfg["d"], cfg["g"], enc)
    elif kind == "attr":
        emb = cb.AttrEmbedding(cfg["d"], cfg["g"], cfg["A"], enc)
    else:
        emb = cb.AttrCtxEmbedding(cfg["d"], cfg["g"], cfg["C"], cfg["A"], enc)
    blocks = nn.ModuleList([cb.SelfAttentionBlock(cfg["d"], cfg["H"], p, cfg["residual_sa"])
                            for _ in range(cfg["n_blocks"])])
    if cfg["decoder"] == "ca":
        dec = cb.CrossAttentionBlock(cfg["d"], cfg["H"], p, cfg["residual_ca"])
    elif cfg["decoder"] == "dot":
        dec = cb.DotProduct()
    else:
        dec = cb.WeightedDotProduct(cfg["gamma"], cfg["L"], cfg["decoder"] == "wdot_norm", device)
    model = cb.CARCA(d=cfg["d"], p=p, emb=emb, enc=blocks, dec=dec)
    model.load_state_dict(sd, strict=True)          # state_dict keys/shapes == the reference's
    return model.to(device)


def _attr_arg(mode, z, model, a_dense, device):
    """dense: the reference API tensor; csr/table: device-resident ItemAttrTable + a=None."""
    if mode == "dense":
        return a_dense.to(device)
    table = cb.ItemAttrTable.from_dense(z["attr_table"], sparse=(mode == "csr")).to(device)
    if hasattr(model.embeds, "set_attr_table"):           # the id-only embeddings take no attributes
        model.embeds.set_attr_table(table)
    return None


def check_eval(name, device, mode="dense"):
    cfg, sd, z = load_case(name)
    model = build_model(cfg, sd, device).eval()
    p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch_of(z, "eval")
    pa = _attr_arg(mode, z, model, p_a, device)
    oa = None if pa is None else o_a.to(device)
    with torch.no_grad():
        y = model.forward(profile=(p_x.to(device), pa, p_c.to(device)), targets=[(o_x.to(device), oa, o_c.to(device))])
        o_mask = cb.get_mask(o_x.to(device))
        loss = cb.BinaryCrossEntropy().forward(y, y_true.to(device), o_mask)
        hr = cb.compute_HR(y, y_true.to(device), cfg["k"])
        ndcg = cb.compute_NDCG(y, y_true.to(device), cfg["k"])
    if mode != "dense":
        from carca_replication_b200 import _native as N
        from carca_replication_b200 import fused

        if fused.supported(model, cfg["L"], cfg["C"]):
            # second call: plan cached -> the whole forward is one kernel launch (+ the row-packing pre-pass of
            # the tensor-core kernel, + the launch of its other decoder variant, which returns at once)
            n0 = N.lib().carca_launch_count()
            with torch.no_grad():
                y_again = model.forward(profile=(p_x.to(device), pa, p_c.to(device)),
                                        targets=[(o_x.to(device), oa, o_c.to(device))])
            assert N.lib().carca_launch_count() - n0 <= 3
            assert torch.equal(y, y_again)
            model.use_fused_eval = False                  # per-op kernels on the same inputs
            with torch.no_grad():
                y_mod = model.forward(profile=(p_x.to(device), pa, p_c.to(device)),
                                      targets=[(o_x.to(device), oa, o_c.to(device))])
            model.use_fused_eval = True
            assert N.lib().carca_launch_count() - n0 > 10
            assert rel_err(y.cpu().numpy(), y_mod.cpu().numpy()) < FP32_RTOL
    ref = z["eval/y_pred"]
    assert tuple(y.shape) == ref.shape
    assert rel_err(y.cpu().numpy(), ref) < FP32_RTOL                       # fp32 scores within 1e-4 rel
    assert topk_equal_up_to_ties(y.cpu().numpy(), ref, cfg["k"], tol=1e-6)  # top-10 identical up to ties
    assert hr == float(z["eval/HR"])
    assert round(ndcg, 3) == round(float(z["eval/NDCG"]), 3)
    assert abs(loss.item() - float(z["eval/loss"])) < 1e-4 * max(1.0, abs(float(z["eval/loss"])))


def _train_pass(model, batch, device, mode, z):
    p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch
    L = p_x.shape[1]
    pa = _attr_arg(mode, z, model, p_a, device)
    o_ad = None if pa is None else o_a.to(device)
    o_xd, o_cd = o_x.to(device), o_c.to(device)
    tg = [(o_xd[:, :L], None if o_ad is None else o_ad[:, :L], o_cd[:, :L]),
          (o_xd[:, L:], None if o_ad is None else o_ad[:, L:], o_cd[:, L:])]
    model.train()
    model.zero_grad()
    y = model.forward(profile=(p_x.to(device), pa, p_c.to(device)), targets=tg)
    loss = cb.BinaryCrossEntropy().forward(y, y_true.to(device), cb.get_mask(o_xd))
    loss.backward()
    return y, loss


def check_train(name, device, mode="dense", gtol=2e-4):
    cfg, sd, z = load_case(name)
    model = build_model(cfg, sd, device)
    y, loss = _train_pass(model, batch_of(z, "train"), device, mode, z)
    assert rel_err(y.detach().cpu().numpy(), z["train/y_pred"]) < FP32_RTOL
    assert abs(loss.item() - float(z["train/loss"])) < 1e-4 * max(1.0, abs(float(z["train/loss"])))
    for k, prm in model.named_parameters():
        ref = z["train/grad/" + k]
        got = np.zeros_like(ref) if prm.grad is None else prm.grad.cpu().numpy()
        assert got.shape == ref.shape, k
        assert grad_err(got, ref, grad_floor(k)) < gtol, (k, grad_err(got, ref, grad_floor(k)))
    if hasattr(model.embeds, "items_embed"):
        assert float(model.embeds.items_embed.weight.grad[0].abs().max()) == 0.0   # padding_idx row


def check_train_dropout(name, device, p=0.3, seed=20240607, mode="dense", gtol=3e-4):
    """Train step with dropout ON: the oracle replays the kernels' Philox keep/drop decisions."""
    from oracle import carca_oracle as O

    cfg, sd, z = load_case(name)
    batch = batch_of(z, "train")
    oc = oracle_cfg(cfg, p_drop=p, seed=seed)
    sdo = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "embeds.enc.pe") for k, v in sd.items()}
    y_ref = O.carca_forward(sdo, oc, batch[:3], O.train_step_targets(*batch[3:6]), training=True)
    loss_ref = O.masked_bce(y_ref, batch[6], O.padding_mask(batch[3]))
    loss_ref.backward()
    model = build_model(cfg, sd, device, p=p)
    ops.set_dropout_seed(seed)
    try:
        y, loss = _train_pass(model, batch, device, mode, z)
    finally:
        ops.set_dropout_seed(None)
    assert rel_err(y.detach().cpu().numpy(), y_ref.detach().numpy()) < FP32_RTOL
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    for k, prm in model.named_parameters():
        ref = sdo[k].grad
        ref = torch.zeros_like(sdo[k]) if ref is None else ref
        got = torch.zeros_like(prm) if prm.grad is None else prm.grad
        e = grad_err(got.cpu().numpy(), ref.numpy(), grad_floor(k))
        assert e < gtol, (k, e)


def check_mha(tag, device):
    z = np.load(f"{GOLDEN}/mha_ops.npz")
    c = json.loads(str(z[f"{tag}/cfg"]))
    mha = cb.MultiHeadAttention(c["d"], c["H"], 0.0)
    mha.load_state_dict({k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}/sd/")})
    mha = mha.to(device).eval()
    q, kv = torch.from_numpy(z[f"{tag}/q"]).to(device), torch.from_numpy(z[f"{tag}/kv"]).to(device)
    qm, km = torch.from_numpy(z[f"{tag}/q_mask"]).to(device), torch.from_numpy(z[f"{tag}/k_mask"]).to(device)
    with torch.no_grad():
        w, o = mha.forward(q, kv, kv, qm, km, causal=c["causal"], return_w=True)
    np.testing.assert_allclose(w.cpu().numpy(), z[f"{tag}/w"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(o.cpu().numpy(), z[f"{tag}/out"], rtol=1e-4, atol=1e-5)
    assert np.all(o.cpu().numpy()[0] == 0.0) and np.all(o.cpu().numpy()[1] == 0.0)   # dead rows exactly 0


def check_metrics(device):
    z = np.load(f"{GOLDEN}/metrics_ops.npz")
    y, yt, m, k = (torch.from_numpy(z["y_pred"]).to(device), torch.from_numpy(z["y_true"]).to(device),
                   torch.from_numpy(z["mask"]).to(device), int(z["k"]))
    assert cb.compute_HR(y, yt, k) == float(z["HR"])
    assert abs(cb.compute_NDCG(y, yt, k) - float(z["NDCG"])) < 1e-4
    y_ties = torch.from_numpy(z["y_pred_ties"]).to(device)      # massive ties: stable (index-order) rule
    assert cb.compute_HR(y_ties, yt, k) == float(z["HR_stable"])
    assert abs(cb.compute_NDCG(y_ties, yt, k) - float(z["NDCG_stable"])) < 1e-4
    yv = y.clone().requires_grad_(True)
    loss = cb.BinaryCrossEntropy().forward(yv, yt, m)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-5
    np.testing.assert_allclose(yv.grad.cpu().numpy(), z["dy"], rtol=1e-4, atol=1e-8)
    # the same three reductions in one launch (carca_eval_metrics: the per-batch body of evaluate()); called twice,
    # the kernel has to leave its scratch clean for the next call
    from carca_replication_b200 import ops
    o_x = (m != 0).to(torch.int32)
    stats = torch.zeros(4, dtype=torch.float64, device=device)
    for _ in range(2):
        ops.eval_metrics_(stats, y, yt, o_x, k)
    st = stats.cpu().numpy()
    assert st[0] == 2 * float(z["HR"]) and st[2] == 2 * y.shape[0]
    assert abs(st[1] - 2 * float(z["NDCG"])) < 2e-4 and abs(st[3] - 2 * float(z["loss"])) < 2e-5
    stats.zero_()
    ops.eval_metrics_(stats, y_ties, yt, o_x, k)
    assert float(stats[0]) == float(z["HR_stable"]) and abs(float(stats[1]) - float(z["NDCG_stable"])) < 1e-4
    # edge shapes: one row, one candidate; several labelled candidates in a row; strided views; int64 ids
    for Bn, Tn in ((1, 1), (3, 40), (37, 101)):
        g = torch.Generator().manual_seed(100 * Bn + Tn)
        yy = torch.rand((Bn, Tn + 3), generator=g).to(device)[:, 1:Tn + 1]            # non-contiguous rows
        tt = (torch.rand((Bn, Tn), generator=g) < 0.2).to(torch.int32).to(device)
        tt[:, 0] = 1
        xx = (torch.rand((Bn, Tn), generator=g) < 0.8).to(torch.int64).to(device) * 7
        xx[:, 0] = 5
        st2 = torch.zeros(4, dtype=torch.float64, device=device)
        ops.eval_metrics_(st2, yy, tt, xx, 10)
        acc = torch.zeros(3, dtype=torch.float64, device=device)
        ops.rank_metrics_(acc, yy, tt, 10)
        loss = cb.BinaryCrossEntropy().forward(yy.contiguous(), tt, cb.get_mask(xx))
        assert torch.allclose(st2[:3], acc, rtol=0, atol=1e-9), (Bn, Tn)
        assert abs(float(st2[3]) - float(loss)) < 1e-5 * max(1.0, abs(float(loss))), (Bn, Tn)


def check_knn(device):
    """KNN baseline (src/knn.py:8-21) against the fixture of the real reference."""
    z = np.load(f"{GOLDEN}/knn_ops.npz")
    p_a, o1, o2 = (torch.from_numpy(z[k]).to(device) for k in ("p_a", "o_a1", "o_a2"))
    zero = torch.zeros(1, device=device)
    y = cb.KNN().forward((zero, p_a, zero), [(zero, o1, zero), (zero, o2, zero)])
    np.testing.assert_allclose(y.cpu().numpy(), z["y"], rtol=1e-5, atol=1e-6)


def check_pickle_and_shapes(device):
    cfg, sd, z = load_case("single_user_ca")
    model = build_model(cfg, sd, device).eval()
    buf = io.BytesIO()
    torch.save(model, buf)                                   # src/train.py:124 pickles the module
    buf.seek(0)
    clone = torch.load(buf, weights_only=False)
    p_x, p_a, p_c, o_x, o_a, o_c, _ = [t.to(device) for t in batch_of(z, "eval")]
    with torch.no_grad():
        y1 = model.forward((p_x, p_a, p_c), [(o_x, o_a, o_c)])
        y2 = clone.forward((p_x, p_a, p_c), [(o_x, o_a, o_c)])
    assert tuple(y1.shape) == (1, cfg["T"])                  # [B, T] even for B == 1
    assert torch.equal(y1, y2)


def check_fused_vs_modular(device, shape_name="tiny", B=5, T=None, decoder="ca", all_valid=False, seed=11):
    """Fused one-kernel inference vs the per-op kernels on seeded synthetic data (any T, incl.
    candidate chunks beyond one 104-row tile and a ragged last user tile)."""
    import dataclasses

    from carca_replication_b200 import _native as N
    from carca_replication_b200 import fused, synth

    shape = synth.SHAPES[shape_name]
    if T is not None:
        shape = dataclasses.replace(shape, n_targets=T)
    model = synth.build_model(shape, decoder, p=0.5, seed=seed).to(device).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=seed).to(device))
    b = {k: v.to(device) for k, v in synth.make_eval_batch(shape, B, seed=seed, all_valid=all_valid).items()}
    assert fused.supported(model, shape.seq_len, shape.n_ctx)
    with torch.no_grad():
        y_f = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
        model.use_fused_eval = False
        y_m = model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])])
        model.use_fused_eval = True
    assert tuple(y_f.shape) == (B, shape.n_targets)
    assert rel_err(y_f.cpu().numpy(), y_m.cpu().numpy()) < FP32_RTOL
    assert topk_equal_up_to_ties(y_f.cpu().numpy(), y_m.cpu().numpy(), 10, tol=1e-6)
    return y_f


def check_fused_train_vs_per_op(device, B=7, decoder="ca", p=0.3, heads=2, n_tuples=2, all_valid=False, odd_masks=False,
                                shape_name="tiny", seed=5, gtol=2e-4):
    """Fused training core (csrc/fused_train.cuh: active positions packed into 64-row bins, one forward and one
    backward kernel) vs the per-op kernels on the same inputs, weights and Philox seed: probabilities, loss and
    every parameter gradient.  odd_masks punches holes into the windows (padding inside a profile, targets at
    padded profile positions and the reverse) — layouts the reference loader never builds but its model accepts."""
    import dataclasses

    from carca_replication_b200 import _native as N
    from carca_replication_b200 import synth

    shape = dataclasses.replace(synth.SHAPES[shape_name], n_heads=heads)
    L = shape.seq_len
    b = synth.make_train_batch(shape, B, seed=seed, all_valid=all_valid)
    if odd_masks:
        g = torch.Generator().manual_seed(seed)
        for key, frac in (("p_x", 0.15), ("o_x", 0.2)):
            hole = torch.rand(b[key].shape, generator=g) < frac
            b[key] = torch.where(hole, torch.zeros_like(b[key]), b[key])
        b["o_x"][0, :L] = torch.randint(1, shape.n_items, (L,), generator=g, dtype=torch.int32)   # targets everywhere
        b["p_x"][1] = 0                                                                          # empty profile
    b = {k: v.to(device) for k, v in b.items()}
    table = synth.make_attr_table(shape, seed=seed).to(device)
    results = []
    for fused in (True, False):
        model = synth.build_model(shape, decoder, p=p, seed=seed).to(device).train()
        model.embeds.set_attr_table(table)
        model.use_fused_train = fused
        tg = [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])][:n_tuples]
        ops.set_dropout_seed(1234 + seed)
        try:
            n0 = N.lib().carca_launch_count()
            y = model.forward((b["p_x"], None, b["p_c"]), tg)
            mask = cb.get_mask(b["o_x"][:, :n_tuples * L])
            loss = cb.BinaryCrossEntropy().forward(y, b["y_true"][:, :n_tuples * L], mask)
            loss.backward()
            launches = N.lib().carca_launch_count() - n0
        finally:
            ops.set_dropout_seed(None)
        results.append((y.detach().cpu().numpy(), loss.item(), {k: v.grad.cpu().numpy() for k, v in
                                                                 model.named_parameters() if v.grad is not None},
                        launches))
    (y_f, loss_f, g_f, n_f), (y_m, loss_m, g_m, n_m) = results
    assert n_f < n_m
    assert y_f.shape == y_m.shape == (B, n_tuples * L)
    assert rel_err(y_f, y_m) < FP32_RTOL
    assert abs(loss_f - loss_m) < 1e-5 * max(1.0, abs(loss_m))
    assert set(g_f) == set(g_m)
    for k in g_m:
        e = grad_err(g_f[k], g_m[k], grad_floor(k))
        assert e < gtol, (k, e)


def check_fused_adam(device, weight_decay=0.0, steps=5):
    """FusedAdam (one launch, csrc/optim.cuh) vs torch.optim.Adam with the constructor arguments of
    scripts/training.py:174 on tensors of awkward sizes (unaligned tails, > one 4096-element chunk)."""
    g = torch.Generator().manual_seed(3)
    shapes = [(5, 7), (4096,), (4097,), (3,), (64, 64, 1), (130, 64), (1,)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    ours = [t.detach().clone().to(device).requires_grad_(True) for t in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=weight_decay, betas=(0.9, 0.98))
    o_ours = cb.FusedAdam(ours, lr=1e-3, weight_decay=weight_decay, betas=(0.9, 0.98))
    for step in range(steps):
        for a, b in zip(ref, ours):
            gr = torch.randn(a.shape, generator=g) * (0.1 + step)
            a.grad = gr.clone()
            b.grad = None if (step == 1 and a.numel() == 3) else gr.to(device)    # a parameter without grad is skipped
            if b.grad is None:
                a.grad = None
        o_ref.step()
        o_ours.step()
    for a, b in zip(ref, ours):
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().numpy(), rtol=2e-5, atol=1e-7)
    sd = o_ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}


def check_fused_train_edges(device):
    """Boundary shapes of the fused training core against the per-op kernels: L = 64 with every position valid (a
    user fills a bin), L = 1, B = 1, a batch whose rows are all padding, 8 context features (the in-kernel
    embedding's limit) and 9 (embedding on the per-op kernels, core still fused)."""
    import dataclasses

    from carca_replication_b200 import synth

    base = synth.SHAPES["tiny"]
    cases = [dict(shape=dataclasses.replace(base, seq_len=64), B=3, all_valid=True, decoder="ca"),
             dict(shape=dataclasses.replace(base, seq_len=64), B=5, all_valid=False, decoder="dot"),
             dict(shape=dataclasses.replace(base, seq_len=1), B=4, all_valid=True, decoder="ca"),
             dict(shape=base, B=1, all_valid=False, decoder="ca"),
             dict(shape=base, B=4, all_valid=False, decoder="ca", empty=True),
             dict(shape=dataclasses.replace(base, n_ctx=8), B=4, all_valid=False, decoder="ca"),
             dict(shape=dataclasses.replace(base, n_ctx=9), B=4, all_valid=False, decoder="dot"),
             dict(shape=dataclasses.replace(base, n_blocks=1, n_heads=4), B=6, all_valid=False, decoder="ca"),
             # L > 64: fused as long as every user's active positions fit one 64-row bin
             dict(shape=dataclasses.replace(base, seq_len=100), B=6, all_valid=False, decoder="ca", max_valid=64),
             dict(shape=dataclasses.replace(base, seq_len=200), B=5, all_valid=False, decoder="dot", max_valid=40),
             dict(shape=dataclasses.replace(base, seq_len=100), B=4, all_valid=True, decoder="ca", expect_fused=False)]
    for case in cases:
        shape, B = case["shape"], case["B"]
        L = shape.seq_len
        b = synth.make_train_batch(shape, B, seed=9, all_valid=case["all_valid"])
        if case.get("max_valid"):                      # left-pad further: at most max_valid positions per window
            cut = L - case["max_valid"]
            b["p_x"][:, :cut] = 0
            b["o_x"][:, :cut] = 0
            b["o_x"][:, L:L + cut] = 0
            b["p_x"][0, cut:] = 7                      # one user uses the whole budget
            b["o_x"][0, cut:L] = 9
            b["o_x"][0, L + cut:] = 11
        if case.get("empty"):
            b["p_x"].zero_()
            b["o_x"].zero_()
        b = {k: v.to(device) for k, v in b.items()}
        table = synth.make_attr_table(shape, seed=9).to(device)
        out = []
        for fused in (True, False):
            model = synth.build_model(shape, case["decoder"], p=0.2, seed=9).to(device).train()
            model.embeds.set_attr_table(table)
            model.use_fused_train = fused
            tg = [(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])]
            assert model._fused_train_applies((b["p_x"], None, b["p_c"]), tg) == (fused and case.get("expect_fused", True))
            ops.set_dropout_seed(77)
            try:
                y = model.forward((b["p_x"], None, b["p_c"]), tg)
                (y * torch.linspace(0.5, 1.5, y.numel(), device=y.device).view_as(y)).sum().backward()
            finally:
                ops.set_dropout_seed(None)
            out.append((y.detach().cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in model.named_parameters()
                                                   if v.grad is not None}))
        (y_f, g_f), (y_m, g_m) = out
        assert rel_err(y_f, y_m) < FP32_RTOL, case
        assert set(g_f) == set(g_m)
        for k in g_m:
            e = grad_err(g_f[k], g_m[k], grad_floor(k))
            # d/d WK.bias is mathematically zero: both sides hold summation noise of this un-normalised loss
            assert e < (5e-3 if k.endswith("WK.bias") else 3e-4), (case, k, e)


def _tuple_batches(shape, n_batches, B, seed, device):
    """Eval batches as the 7-tuples a DataLoader over CARCADataset yields (p_a = o_a = None: device attr table)."""
    from carca_replication_b200 import synth

    out = []
    for i in range(n_batches):
        b = synth.make_eval_batch(shape, B, seed=seed + i)
        out.append(b)
    return out


def check_evaluate_vs_oracle(device, decoder="ca"):
    """evaluate() (src/train.py:35-53 signature; device accumulators, one host read) against the oracle's
    per-batch body on the same batches: HR@10 and NDCG@10 equal to 3 decimals, mean batch loss to 1e-4."""
    from carca_replication_b200 import synth
    from oracle import carca_oracle as O

    shape = synth.TINY
    model = synth.build_model(shape, decoder, p=0.5, seed=21)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=21)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder)
    batches = _tuple_batches(shape, 3, 11, 40, device)
    hits = ndcg = loss = 0.0
    users = 0
    for b in batches:
        h, n, l, u = O.eval_batch(sd, cfg, (b["p_x"], table.gather_dense(b["p_x"]), b["p_c"], b["o_x"],
                                            table.gather_dense(b["o_x"]), b["o_c"], b["y_true"]))
        hits, ndcg, loss, users = hits + h, ndcg + n, loss + l, users + u
    model = model.to(device)
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=21).to(device))
    loader = [(b["p_x"], None, b["p_c"], b["o_x"], None, b["o_c"], b["y_true"]) for b in batches]
    hr, nd, ls = cb.evaluate(model, loader, device, 10)
    assert round(hr, 3) == round(hits / users, 3)
    assert round(nd, 3) == round(ndcg / users, 3)
    assert abs(ls - loss / len(batches)) < 1e-4 * max(1.0, abs(loss))


def check_weights_epoch(device, d=64):
    """Caches derived from the weights (fused inference plan, AllEmbedding's folded item table) follow parameter
    updates that tensor version counters cannot see: FusedAdam writes through raw pointers.  After k optimizer steps
    the fused path, the folded per-op path and the unfolded per-op path must agree on the NEW weights."""
    import dataclasses

    from carca_replication_b200 import synth

    shape = dataclasses.replace(synth.TINY, d=d)
    model = synth.build_model(shape, "ca", p=0.0, seed=3).to(device)
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=3).to(device))
    b = {k: v.to(device) for k, v in synth.make_eval_batch(shape, 6, seed=3).items()}
    bt = {k: v.to(device) for k, v in synth.make_train_batch(shape, 6, seed=4).items()}
    L = shape.seq_len
    opt = cb.FusedAdam(model.parameters(), lr=5e-2, betas=(0.9, 0.98))

    def scores(fused_on, folded_on):
        model.eval()
        model.use_fused_eval, model.embeds.use_folded_eval = fused_on, folded_on
        with torch.no_grad():
            return model.forward((b["p_x"], None, b["p_c"]), [(b["o_x"], None, b["o_c"])]).cpu().numpy()

    before = scores(True, True)          # builds the plan / the folded table from the initial weights
    scores(False, True)
    versions = [p._version for p in model.parameters()]
    for _ in range(3):
        model.train()
        opt.zero_grad()
        y = model.forward((bt["p_x"], None, bt["p_c"]),
                          [(bt["o_x"][:, :L], None, bt["o_c"][:, :L]), (bt["o_x"][:, L:], None, bt["o_c"][:, L:])])
        cb.BinaryCrossEntropy().forward(y, bt["y_true"], cb.get_mask(bt["o_x"])).backward()
        opt.step()
    assert versions == [p._version for p in model.parameters()]      # the premise: versions did not move
    ref = scores(False, False)           # per-op kernels, unfolded embedding: reads the parameters directly
    assert rel_err(before, ref) > 1e-3   # the weights did move
    assert rel_err(scores(False, True), ref) < FP32_RTOL
    assert rel_err(scores(True, True), ref) < FP32_RTOL
    model.use_fused_eval, model.embeds.use_folded_eval = True, True


def check_train_loop_checkpoint(device, tmpdir):
    """train() (src/train.py:56-152 signature) end to end on a tiny synthetic set: CSV log format, best-NDCG
    checkpoint written as a state_dict file that the weights-only unpickler reads and that loads strictly into a
    fresh model (and into the reference's own key layout), returned model == checkpoint."""
    import os

    from carca_replication_b200 import synth
    from carca_replication_b200.train import CHECKPOINT_FORMAT, load_checkpoint, train

    shape = synth.TINY
    model = synth.build_model(shape, "dot", p=0.2, seed=8).to(device)
    table = synth.make_attr_table(shape, seed=8).to(device)
    model.embeds.set_attr_table(table)
    tr = [synth.make_train_batch(shape, 8, seed=60 + i) for i in range(3)]
    ev = [synth.make_eval_batch(shape, 8, seed=70 + i) for i in range(2)]
    as_tuple = lambda b: (b["p_x"], None, b["p_c"], b["o_x"], None, b["o_c"], b["y_true"])   # noqa: E731
    opt = cb.FusedAdam(model.parameters(), lr=1e-2, betas=(0.9, 0.98))
    datadir = os.path.join(str(tmpdir), "run")
    cwd = os.getcwd()
    os.chdir(str(tmpdir))
    try:
        out = train(model, [as_tuple(b) for b in tr], [as_tuple(b) for b in ev], [as_tuple(b) for b in ev], device, opt,
                    epochs=3, top_k=10, verbose=1, early_stop=5, datadir="run")
    finally:
        os.chdir(cwd)
    files = sorted(os.listdir(datadir))
    pth = [f for f in files if f.endswith(".pth")]
    assert len(pth) == 1 and len(pth[0].split("_")) == 3                       # {epoch:03d}_{HR:.4f}_{NDCG:.4f}.pth
    blob = torch.load(os.path.join(datadir, pth[0]), weights_only=True)        # no pickled code in the file
    assert blob["format"] == CHECKPOINT_FORMAT and int(pth[0][:3]) == blob["epoch"]
    assert set(blob["state_dict"]) == set(model.state_dict())
    fresh = synth.build_model(shape, "dot", p=0.2, seed=99).to(device)
    fresh.embeds.set_attr_table(table)
    load_checkpoint(fresh, os.path.join(datadir, pth[0]))
    for (k, a), (_, c) in zip(out.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a.cpu(), c.cpu()), k
    csv = [f for f in files if f.endswith(".csv")]
    assert len(csv) == 1
    rows = [r.split(";") for r in open(os.path.join(datadir, csv[0])).read().strip().split("\n")]
    assert all(len(r) == 6 for r in rows) and {r[2] for r in rows} == {"train", "val", "test"}   # time;epoch;split;loss;HR;NDCG


# ------------------------------------------------------------------------------- packed-rows pipeline (fp32 / bf16)
def _logit(p):
    p = np.asarray(p, dtype=np.float64)
    return np.log(np.clip(p, 1e-300, None)) - np.log(np.clip(1.0 - p, 1e-300, None))


def oracle_scores(shape, decoder, batch, seed):
    """Reference scores of a seeded synthetic model on `batch` (CPU oracle, dense attributes as the reference API)."""
    from carca_replication_b200 import synth
    from oracle import carca_oracle as O

    model = synth.build_model(shape, decoder, p=0.5, seed=seed)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    table = synth.make_attr_table(shape, seed=seed)
    cfg = O.OracleConfig(d=shape.d, n_heads=shape.n_heads, n_blocks=shape.n_blocks, decoder=decoder)
    with torch.no_grad():
        y = O.carca_forward(sd, cfg, (batch["p_x"], table.gather_dense(batch["p_x"]), batch["p_c"]),
                            [(batch["o_x"], table.gather_dense(batch["o_x"]), batch["o_c"])], training=False)
    return model, y.numpy()


def run_eval_path(model, shape, batch, device, seed, dtype="fp32", path=None, expand_ctx=False, no_sync=False):
    from carca_replication_b200 import synth

    model = model.to(device).eval()
    model.embeds.set_attr_table(synth.make_attr_table(shape, seed=seed).to(device))
    model.set_eval_dtype(dtype)
    model.force_eval_path = path
    d = {k: v.to(device) for k, v in batch.items()}
    o_c = d["o_c"][:, :1, :].contiguous().expand(-1, d["o_x"].shape[1], -1) if expand_ctx else d["o_c"]
    with torch.no_grad():
        y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, o_c)])          # builds plans / scratch
        if no_sync:
            # the decision which kernels run, and the run itself, never read the device: a host sync raises here
            torch.cuda.synchronize()
            torch.cuda.set_sync_debug_mode("error")
            try:
                y = model.forward((d["p_x"], None, d["p_c"]), [(d["o_x"], None, o_c)])
            finally:
                torch.cuda.set_sync_debug_mode("default")
    return model, y, d


def assert_fp32_parity(y, y_ref, d, B, k=10):
    from oracle import carca_oracle as O

    yn = y.cpu().numpy()
    assert rel_err(yn, y_ref) < FP32_RTOL
    assert topk_equal_up_to_ties(yn, y_ref, k, tol=1e-6)
    yr = torch.from_numpy(y_ref)
    assert cb.compute_HR(y, d["y_true"], k) == O.hit_count(yr, d["y_true"].cpu(), k)
    assert round(cb.compute_NDCG(y, d["y_true"], k) / B, 3) == round(O.ndcg_sum(yr, d["y_true"].cpu(), k) / B, 3)


def bf16_errors(y, y_ref):
    """(max |dp|, max |dlogit| over unsaturated candidates / max(1, max |logit_ref| of the batch), top-10 overlap)."""
    yn = np.asarray(y, dtype=np.float64)
    live = (y_ref > 1e-6) & (y_ref < 1 - 1e-6)           # candidates whose fp32 probability still resolves the logit
    lr = _logit(y_ref)
    dl = np.abs(_logit(yn) - lr)[live]
    finite = (y_ref > 0) & (y_ref < 1)                   # the batch's logit scale: every logit fp32 can represent
    scale = max(1.0, float(np.abs(lr[finite]).max())) if finite.any() else 1.0
    top = np.mean([len(set(np.argsort(-a, kind="stable")[:10]) & set(np.argsort(-r, kind="stable")[:10])) / 10.0
                   for a, r in zip(yn, y_ref)])
    return float(np.abs(yn - y_ref).max()), float(dl.max() / scale) if dl.size else 0.0, float(top)


# bf16 contract (north_star: "bf16 within 1e-2"; fp32 accumulation, softmax, LayerNorm, embedding and decoder):
#   probabilities within 1e-2 absolute for the cross-attention decoder (logits O(1));
#   for every decoder the logit error of unsaturated candidates within 1e-2 of the batch's logit scale;
#   top-10 lists overlap >= 98 %, HR@10 within max(2, 5 %) users of the batch, NDCG@10 within 1e-2
#   (the dot decoder's logits reach +-70 at d = 256: a 1e-3 relative logit error moves near-tied candidates).
BF16_TOL = 1e-2


def assert_bf16_parity(y, y_ref, d, B, decoder, k=10):
    from oracle import carca_oracle as O

    dp, dl, top = bf16_errors(y.cpu().numpy(), y_ref)
    assert np.isfinite(y.cpu().numpy()).all()
    if decoder == "ca":
        assert dp < BF16_TOL, dp
    assert dl < BF16_TOL, dl
    assert top >= 0.98, top
    yr = torch.from_numpy(y_ref)
    # HR / NDCG are compared only when fp32 probabilities still order the candidates: where several candidates of a
    # user saturate to exactly 1.0 (logits > 17, the synthetic dot decoder at d = 256) the reference's rank is decided
    # by its tie rule on saturated values, which no arithmetic with a 1e-3 relative logit error can reproduce
    saturated = bool(((y_ref >= 1.0).sum(1) > 1).any())
    if not saturated:
        assert abs(cb.compute_HR(y, d["y_true"], k) - O.hit_count(yr, d["y_true"].cpu(), k)) <= max(2, 0.05 * B)
        assert abs(cb.compute_NDCG(y, d["y_true"], k) - O.ndcg_sum(yr, d["y_true"].cpu(), k)) / B < 1e-2
    return dp, dl, top


def long_window_batch(shape, B, seed, n_long=3):
    """Eval batch with left-padded windows of `shape.seq_len` positions in which `n_long` users have MORE than 64
    valid positions (one of them every position) — the case a 64-row bin cannot hold."""
    from carca_replication_b200 import synth

    b = synth.make_eval_batch(shape, B, seed=seed)
    full = synth.make_eval_batch(shape, B, seed=seed + 1, all_valid=True)
    L = shape.seq_len
    for i in range(min(n_long, B)):
        keep = L if i == 0 else min(L, 65 + 17 * i)
        b["p_x"][i] = full["p_x"][i]
        b["p_c"][i] = full["p_c"][i]
        b["p_x"][i, : L - keep] = 0
        b["p_c"][i, : L - keep] = 0.0
    return b
