"""Generates tests/golden/*.npz by running the REAL reference implementation.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
The reference is imported, never copied.  Each fixture holds a config, the
reference model's state_dict, seeded synthetic inputs, and the outputs /
gradients the reference produced for them on CPU (torch fp32).  The fixtures pin
`oracle/carca_oracle.py` (tests/test_oracle_golden.py) and are also compared
directly with the CUDA path (tests/test_gpu_golden.py) on the GPU box, where
/root/reference does not exist.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("CARCA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from src.carca import (  # noqa: E402
    CARCA, AllEmbedding, AttrCtxEmbedding, AttrEmbedding, BinaryCrossEntropy, CrossAttentionBlock, DotProduct,
    IdEmbedding, IdentityEncoding, LearnableEncoding, MLPIdEmbedding, MultiHeadAttention, PositionalEncoding,
    SelfAttentionBlock, WeightedDotProduct,
)
from src.knn import KNN  # noqa: E402
from src.train import compute_HR, compute_NDCG  # noqa: E402
from src.utils import get_mask  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build_reference(cfg, seed):
    torch.manual_seed(seed)
    if cfg["encoding"] == "identity":
        enc = IdentityEncoding()
    elif cfg["encoding"] == "learnable":
        enc = LearnableEncoding(cfg["d"], cfg["L"])
    else:
        enc = PositionalEncoding(cfg["d"], cfg["L"])
    kind = cfg.get("embedding", "all")           # scripts/training.py:78-89
    if kind == "all":
        emb = AllEmbedding(cfg["n_items"], cfg["d"], cfg["g"], cfg["C"], cfg["A"], enc)
    elif kind == "id":
        emb = IdEmbedding(cfg["n_items"], cfg["d"], enc)
    elif kind == "mlpid":
        emb = MLPIdEmbedding(cfg["n_items"], cfg["d"], cfg["g"], enc)
    elif kind == "attr":
        emb = AttrEmbedding(cfg["d"], cfg["g"], cfg["A"], enc)
    else:
        emb = AttrCtxEmbedding(cfg["d"], cfg["g"], cfg["C"], cfg["A"], enc)
    blocks = nn.ModuleList([SelfAttentionBlock(cfg["d"], cfg["H"], cfg["p"], cfg["residual_sa"])
                            for _ in range(cfg["n_blocks"])])
    if cfg["decoder"] == "ca":
        dec = CrossAttentionBlock(cfg["d"], cfg["H"], cfg["p"], cfg["residual_ca"])
    elif cfg["decoder"] == "dot":
        dec = DotProduct()
    else:                                        # scripts/training.py:97-98
        dec = WeightedDotProduct(cfg["gamma"], cfg["L"], cfg["decoder"] == "wdot_norm", "cpu")
    model = CARCA(d=cfg["d"], p=cfg["p"], emb=emb, enc=blocks, dec=dec)
    # zero biases / unit LayerNorms would hide bias bugs: perturb them
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if name.endswith("bias") or ".norm" in name or name.startswith("norm."):
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
        if hasattr(model.embeds, "items_embed"):
            model.embeds.items_embed.weight[0].zero_()
    return model


def make_attr_table(cfg, rng):
    n, A = cfg["n_items"], cfg["A"]
    tab = np.zeros((n, A), dtype=np.float32)
    if cfg["attr_kind"] == "multihot":
        for i in range(1, n):
            nnz = int(np.clip(rng.poisson(4), 1, min(A, 9)))
            tab[i, rng.choice(A, size=nnz, replace=False)] = 1.0
    else:  # dense (image-feature-like)
        tab[1:] = rng.standard_normal((n - 1, A)).astype(np.float32)
    return tab


def make_batch(cfg, rng, tab, lens, mode):
    """Left-padded windows with the layout of src/data.py:90-192 (train: [B,2L] pos|neg)."""
    B, L, C, n = len(lens), cfg["L"], cfg["C"], cfg["n_items"]
    T = cfg["T"]
    p_x = np.zeros((B, L), np.int32)
    p_c = np.zeros((B, L, C), np.float32)
    if mode == "eval":
        o_x = np.zeros((B, T), np.int32)
        o_c = np.zeros((B, T, C), np.float32)
        y = np.zeros((B, T), np.int32)
    else:
        o_x = np.zeros((B, 2 * L), np.int32)
        o_c = np.zeros((B, 2 * L, C), np.float32)
        y = np.zeros((B, 2 * L), np.int32)
    for b, n_valid in enumerate(lens):
        items = rng.choice(np.arange(1, n), size=n_valid + 1, replace=False)
        ctx = rng.random((n_valid + 1, C)).astype(np.float32)
        if n_valid:
            p_x[b, L - n_valid:] = items[:n_valid]
            p_c[b, L - n_valid:] = ctx[:n_valid]
        if mode == "eval":
            o_x[b, 0] = items[n_valid]
            o_x[b, 1:] = rng.choice(np.setdiff1d(np.arange(1, n), items), size=T - 1, replace=False)
            o_c[b, :] = ctx[n_valid]            # negatives take the positive's context (data.py:185)
            y[b, 0] = 1
        else:
            if n_valid:
                o_x[b, L - n_valid:L] = items[1:n_valid + 1]
                o_c[b, L - n_valid:L] = ctx[1:n_valid + 1]
                o_x[b, 2 * L - n_valid:] = rng.choice(np.setdiff1d(np.arange(1, n), items), size=n_valid,
                                                      replace=False)
                o_c[b, 2 * L - n_valid:] = ctx[1:n_valid + 1]   # data.py:130
            y[b, :L] = (p_x[b] > 0)                             # data.py:134-135
    return dict(p_x=p_x, p_a=tab[p_x], p_c=p_c, o_x=o_x, o_a=tab[o_x], o_c=o_c, y_true=y)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def run_case(name, cfg, lens_eval, lens_train, seed):
    rng = np.random.default_rng(seed)
    model = build_reference(cfg, seed)
    tab = make_attr_table(cfg, rng)
    out = {"cfg": np.array(json.dumps(cfg)), "attr_table": tab}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.detach().numpy().copy()

    # ---- eval: forward + loss + HR/NDCG, src/train.py:41-51
    be = make_batch(cfg, rng, tab, lens_eval, "eval")
    model.eval()
    with torch.no_grad():
        y = model.forward(profile=(t(be["p_x"]), t(be["p_a"]), t(be["p_c"])),
                          targets=[(t(be["o_x"]), t(be["o_a"]), t(be["o_c"]))])
        y2 = y.reshape(len(lens_eval), -1)       # reference squeeze() collapses B==1 (carca.py:346)
        loss = BinaryCrossEntropy().forward(y2, t(be["y_true"]), get_mask(t(be["o_x"])))
        out["eval/HR"] = np.array(compute_HR(y2, t(be["y_true"]), cfg["k"]), np.float64)
        out["eval/NDCG"] = np.array(compute_NDCG(y2, t(be["y_true"]), cfg["k"]), np.float64)
    for k_, v in be.items():
        out["eval/in/" + k_] = v
    out["eval/y_pred"] = y2.numpy().copy()
    out["eval/loss"] = loss.numpy().copy()

    # ---- train (dropout p as configured; fixtures use p=0 so the pass is deterministic),
    #      src/train.py:84-95
    bt = make_batch(cfg, rng, tab, lens_train, "train")
    model.train()
    model.zero_grad()
    L = cfg["L"]
    o_x, o_a, o_c = t(bt["o_x"]), t(bt["o_a"]), t(bt["o_c"])
    y = model.forward(profile=(t(bt["p_x"]), t(bt["p_a"]), t(bt["p_c"])),
                      targets=[(o_x[:, :L], o_a[:, :L], o_c[:, :L]), (o_x[:, L:], o_a[:, L:], o_c[:, L:])])
    y2 = y.reshape(len(lens_train), -1)
    loss = BinaryCrossEntropy().forward(y2, t(bt["y_true"]), get_mask(o_x))
    loss.backward()
    for k_, v in bt.items():
        out["train/in/" + k_] = v
    out["train/y_pred"] = y2.detach().numpy().copy()
    out["train/loss"] = loss.detach().numpy().copy()
    for k_, prm in model.named_parameters():
        out["train/grad/" + k_] = (prm.grad if prm.grad is not None else torch.zeros_like(prm)).numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: eval loss {float(out['eval/loss']):.6f} HR {float(out['eval/HR'])} "
          f"NDCG {float(out['eval/NDCG']):.4f}  train loss {float(out['train/loss']):.6f}")


def run_mha_cases(seed):
    """MultiHeadAttention with return_w for the three mask regimes (carca.py:228-265)."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    out = {}
    for tag, (B, Lq, Lk, d, H, causal) in {
        "self": (4, 10, 10, 32, 2, 0),
        "cross_eval": (3, 13, 7, 32, 4, None),
        "cross_train": (3, 9, 9, 64, 2, -1),
    }.items():
        mha = MultiHeadAttention(d, H, 0.0)
        with torch.no_grad():
            for prm in mha.parameters():
                if prm.ndim == 1:
                    prm.add_(0.1 * torch.randn(prm.shape))
        q = torch.randn(B, Lq, d)
        kv = torch.randn(B, Lk, d)
        qm = t((rng.random((B, Lq)) > 0.3).astype(np.float32))
        km = t((rng.random((B, Lk)) > 0.3).astype(np.float32))
        qm[0, :] = 0.0          # a batch row with every query dead
        km[1, :] = 0.0          # a batch row with every key dead
        with torch.no_grad():
            w, o = mha.forward(q, kv, kv, qm, km, causal=causal, return_w=True)
        for k_, v in mha.state_dict().items():
            out[f"{tag}/sd/{k_}"] = v.numpy().copy()
        out[f"{tag}/cfg"] = np.array(json.dumps(dict(B=B, Lq=Lq, Lk=Lk, d=d, H=H, causal=causal)))
        out[f"{tag}/q"], out[f"{tag}/kv"] = q.numpy(), kv.numpy()
        out[f"{tag}/q_mask"], out[f"{tag}/k_mask"] = qm.numpy(), km.numpy()
        out[f"{tag}/w"], out[f"{tag}/out"] = w.numpy(), o.numpy()
    np.savez_compressed(os.path.join(OUT, "mha_ops.npz"), **out)
    print("mha_ops: ok")


def run_metric_cases(seed):
    """BCE + HR/NDCG incl. ties with the positive (train.py:15-32, carca.py:441-444).

    torch.sort is not stable by default, so the reference's order inside a group of tied scores
    is implementation-defined (observed here: a 4-way tie that includes the positive does NOT come
    out in index order).  Parity is therefore "up to ties" (north_star): `HR`/`NDCG` are pinned on
    tie-free scores, and the tie rule the kernels implement — the STABLE one, ties keep index
    order, so a positive in column 0 wins — is pinned by `*_stable`, the reference formulas run
    with sort(stable=True) on rows with small and massive tie groups.
    """
    rng = np.random.default_rng(seed)
    B, T, k = 64, 101, 10
    y = rng.random((B, T)).astype(np.float32)
    y[:8, 0] = 0.999                      # easy hits
    y[20:24, 0] = 0.0                     # positive last
    yt = np.zeros((B, T), np.int32)
    yt[:, 0] = 1
    yt[60:, 5] = 1                        # rows with two positives
    mask = (rng.random((B, T)) > 0.2).astype(np.float32)
    out = dict(y_pred=y, y_true=yt, mask=mask, k=np.array(k))
    out["HR"] = np.array(compute_HR(t(y), t(yt), k), np.float64)
    out["NDCG"] = np.array(compute_NDCG(t(y), t(yt), k), np.float64)
    yv = t(y).clone().requires_grad_(True)
    loss = BinaryCrossEntropy().forward(yv, t(yt), t(mask))
    loss.backward()
    out["loss"] = loss.detach().numpy().copy()
    out["dy"] = yv.grad.numpy().copy()

    y2 = y.copy()
    y2[8:16, 1:4] = y2[8:16, 0:1]         # 4-way ties with the positive
    y2[16:20] = 0.5                       # every candidate tied
    y2[24:28, :50] = 0.75                 # 50-way tie that includes the positive
    _, order = torch.sort(t(y2), descending=True, stable=True)
    top = torch.gather(t(yt), 1, order)[:, :k]
    out["y_pred_ties"] = y2
    out["HR_stable"] = np.array(top.sum().item(), np.float64)
    out["NDCG_stable"] = np.array((1.0 / torch.log2(torch.nonzero(top)[:, 1] + 2)).sum().item(), np.float64)
    np.savez_compressed(os.path.join(OUT, "metrics_ops.npz"), **out)
    print(f"metrics_ops: HR {float(out['HR'])} NDCG {float(out['NDCG']):.4f} loss {float(out['loss']):.6f} "
          f"HR_stable {float(out['HR_stable'])}")


BASE = dict(n_items=120, A=37, C=6, d=64, g=48, H=2, n_blocks=2, L=12, T=21, p=0.0, k=10,
            residual_sa=True, residual_ca=True, decoder="ca", encoding="identity", attr_kind="multihot")

CASES = {
    # name: (cfg overrides, eval profile lengths, train profile lengths)
    "beauty_ca": ({}, [12, 5, 1, 0, 9, 3], [12, 4, 1, 0, 7, 2]),
    "beauty_dot": (dict(decoder="dot"), [12, 5, 1, 0, 9, 3], [12, 4, 1, 0, 7, 2]),
    "men_ca": (dict(n_items=90, A=24, d=64, g=32, H=4, n_blocks=1, L=9, T=11, attr_kind="dense"),
               [9, 2, 6], [8, 3, 9]),
    "men_dot_d128": (dict(n_items=70, A=40, d=128, g=64, H=4, n_blocks=1, L=7, T=11, attr_kind="dense",
                          decoder="dot"), [7, 3], [6, 2]),
    "noresid_dot": (dict(d=32, g=16, H=1, n_blocks=1, residual_sa=False, decoder="dot"), [12, 6, 2], [11, 5, 1]),
    "noresid_ca": (dict(d=32, g=16, H=1, n_blocks=1, residual_sa=False, residual_ca=False), [12, 6, 2], [11, 5, 1]),
    "learnable_ca": (dict(encoding="learnable", n_blocks=1), [12, 3, 7], [10, 2, 5]),
    "sinus_dot": (dict(encoding="positional", n_blocks=1, decoder="dot"), [12, 3, 7], [10, 2, 5]),
    "single_user_ca": (dict(n_blocks=1), [5], [4]),
}

# module variants: written by `python tests/golden/make_golden.py variants`
VARIANT_CASES = {
    "idemb_dot": (dict(embedding="id", decoder="dot", encoding="learnable", n_blocks=1), [12, 5, 1, 0, 9], [12, 4, 1, 0, 7]),
    "mlpid_ca": (dict(embedding="mlpid", n_blocks=1), [12, 3, 7], [10, 2, 5]),
    "attr_dot": (dict(embedding="attr", decoder="dot", n_blocks=1), [12, 5, 0, 9], [12, 4, 0, 7]),
    "attrctx_ca": (dict(embedding="attrctx", encoding="positional", n_blocks=1), [12, 3, 7], [10, 2, 5]),
    "attrctx_dense_ca": (dict(embedding="attrctx", n_blocks=1, n_items=90, A=24, attr_kind="dense"),
                          [12, 2, 6], [8, 3, 9]),
    "wdot_all": (dict(decoder="wdot", gamma=0.3, d=32, g=16, H=1, n_blocks=1), [12, 5, 1, 9], [12, 4, 1, 7]),
    "wdotnorm_all": (dict(decoder="wdot_norm", gamma=0.8, n_blocks=1), [12, 5, 1, 9], [12, 4, 1, 7]),
}


def run_knn_case(seed):
    rng = np.random.default_rng(seed)
    B, L, T, A = 5, 7, 11, 13
    p_a = rng.random((B, L, A)).astype(np.float32)
    o_a1, o_a2 = rng.random((B, T, A)).astype(np.float32), rng.random((B, 3, A)).astype(np.float32)
    z = torch.zeros(1)
    with torch.no_grad():
        y = KNN().forward((z, t(p_a), z), [(z, t(o_a1), z), (z, t(o_a2), z)])
    np.savez_compressed(os.path.join(OUT, "knn_ops.npz"), p_a=p_a, o_a1=o_a1, o_a2=o_a2, y=y.numpy())
    print("knn_ops: ok")


if __name__ == "__main__":
    if "variants" in sys.argv:
        for i, (name, (over, le, lt)) in enumerate(VARIANT_CASES.items()):
            cfg = dict(BASE)
            cfg.update(over)
            run_case(name, cfg, le, lt, seed=4321 + i)
        run_knn_case(79)
    else:
        for i, (name, (over, le, lt)) in enumerate(CASES.items()):
            cfg = dict(BASE)
            cfg.update(over)
            run_case(name, cfg, le, lt, seed=1234 + i)
        run_mha_cases(77)
        run_metric_cases(78)
