"""Seeded synthetic Beauty- / Men-shaped data and weights.

There is no network and no dataset in the build or GPU boxes, so benchmarks and parity tests use
synthetic data with the shapes BASELINE.json names: item ids are 1-based with 0 = padding, windows
are LEFT-padded with the layout of the reference loader (src/data.py:90-192), negatives take the
positive's context (src/data.py:130,185), train labels are [p_x > 0 | 0...] (src/data.py:134-135).
Simplifications (they do not change the arithmetic per position): a user's items may repeat, and a
negative that collides with one of the user's items is re-drawn up to three times, not forever.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .attrs import ItemAttrTable


@dataclass(frozen=True)
class Shape:
    name: str
    n_users: int
    n_items: int          # rows of the item table, including the pad row 0
    n_attrs: int
    attr_kind: str        # "multihot" | "dense"
    n_ctx: int
    d: int
    g: int
    n_heads: int
    n_blocks: int
    seq_len: int
    n_targets: int        # eval candidates per user: 1 positive + 100 negatives


BEAUTY = Shape("beauty", 52_204, 57_290, 6_507, "multihot", 6, 64, 256, 2, 3, 50, 101)
MEN = Shape("men", 34_244, 110_637, 512, "dense", 6, 256, 256, 4, 3, 50, 101)
TINY = Shape("tiny", 500, 300, 41, "multihot", 6, 64, 32, 2, 2, 12, 21)
SHAPES = {s.name: s for s in (BEAUTY, MEN, TINY)}


def make_attr_table(shape: Shape, seed: int = 1234) -> ItemAttrTable:
    """Multi-hot: nnz ~ clip(Poisson(8),1,32) attribute ids ~ Zipf(1.1) over A, built as CSR
    directly (the dense 57K x 6.5K matrix is never materialised).  Dense: N(0,1) rows."""
    rng = np.random.default_rng(seed)
    n, A = shape.n_items, shape.n_attrs
    if shape.attr_kind == "dense":
        tab = rng.standard_normal((n, A), dtype=np.float32)
        tab[0] = 0.0
        return ItemAttrTable(n, A, dense=torch.from_numpy(tab))
    nnz = np.clip(rng.poisson(8, size=n), 1, min(32, A))
    nnz[0] = 0
    w = 1.0 / np.arange(1, A + 1) ** 1.1
    cdf = np.cumsum(w / w.sum())
    rowptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for i in range(n):
        c = np.unique(np.searchsorted(cdf, rng.random(int(nnz[i]))).clip(0, A - 1))
        cols.append(c)
        rowptr[i + 1] = rowptr[i] + c.size
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    return ItemAttrTable(n, A, rowptr=torch.from_numpy(rowptr), cols=torch.from_numpy(cols.astype(np.int64)),
                         vals=torch.ones(cols.size, dtype=torch.float32))


def _lengths(rng, B: int) -> np.ndarray:
    return np.clip(np.rint(rng.lognormal(1.8, 0.7, size=B)), 4, 300).astype(np.int64)


def _zipf_items(rng, shape: Shape, size) -> np.ndarray:
    n = shape.n_items - 1
    w = 1.0 / np.arange(1, n + 1)
    cdf = np.cumsum(w / w.sum())
    return (np.searchsorted(cdf, rng.random(size)).clip(0, n - 1) + 1).astype(np.int32)


def _negatives(rng, shape: Shape, own: np.ndarray, count: int) -> np.ndarray:
    """uniform over [1, n_items-1] avoiding the user's own items (src/data.py:77-87)."""
    B = own.shape[0]
    neg = rng.integers(1, shape.n_items, size=(B, count), dtype=np.int32)
    for _ in range(3):
        clash = (neg[:, :, None] == own[:, None, :]).any(-1)
        if not clash.any():
            break
        neg[clash] = rng.integers(1, shape.n_items, size=int(clash.sum()), dtype=np.int32)
    return neg


def make_eval_batch(shape: Shape, B: int, seed: int = 1234, all_valid: bool = False) -> Dict[str, torch.Tensor]:
    """One evaluate() batch (src/data.py:140-192): p_x [B,L], p_c [B,L,C], o_x [B,T] with the
    positive in column 0, o_c [B,T,C] = the positive's context, y_true [B,T]."""
    rng = np.random.default_rng(seed)
    L, T, C = shape.seq_len, shape.n_targets, shape.n_ctx
    n_valid = np.full(B, L) if all_valid else np.minimum(L, _lengths(rng, B) - 1)
    seq = _zipf_items(rng, shape, (B, L + 1))
    ctx = rng.random((B, L + 1, C), dtype=np.float32)
    valid = np.arange(L)[None, :] >= (L - n_valid)[:, None]
    p_x = np.where(valid, seq[:, :L], 0).astype(np.int32)
    p_c = np.where(valid[:, :, None], ctx[:, :L], 0.0).astype(np.float32)
    o_x = np.empty((B, T), np.int32)
    o_x[:, 0] = seq[:, L]
    o_x[:, 1:] = _negatives(rng, shape, np.concatenate([p_x, seq[:, L:]], 1), T - 1)
    o_c = np.broadcast_to(ctx[:, L:L + 1], (B, T, C)).copy()
    y = np.zeros((B, T), np.int32)
    y[:, 0] = 1
    return {k: torch.from_numpy(v) for k, v in dict(p_x=p_x, p_c=p_c, o_x=o_x, o_c=o_c, y_true=y).items()}


def make_train_batch(shape: Shape, B: int, seed: int = 1234, all_valid: bool = False) -> Dict[str, torch.Tensor]:
    """One train batch (src/data.py:90-137): o_x [B,2L] = next items | negatives."""
    rng = np.random.default_rng(seed)
    L, C = shape.seq_len, shape.n_ctx
    n_valid = np.full(B, L) if all_valid else np.clip(_lengths(rng, B) - 3, 1, L)
    seq = _zipf_items(rng, shape, (B, L + 1))
    ctx = rng.random((B, L + 1, C), dtype=np.float32)
    valid = np.arange(L)[None, :] >= (L - n_valid)[:, None]
    p_x = np.where(valid, seq[:, :L], 0).astype(np.int32)
    p_c = np.where(valid[:, :, None], ctx[:, :L], 0.0).astype(np.float32)
    pos = np.where(valid, seq[:, 1:], 0).astype(np.int32)
    neg = np.where(valid, _negatives(rng, shape, seq, L), 0).astype(np.int32)
    pos_c = np.where(valid[:, :, None], ctx[:, 1:], 0.0).astype(np.float32)
    o_x = np.concatenate([pos, neg], 1)
    o_c = np.concatenate([pos_c, pos_c], 1)
    y = np.concatenate([(p_x > 0).astype(np.int32), np.zeros((B, L), np.int32)], 1)
    return {k: torch.from_numpy(v) for k, v in dict(p_x=p_x, p_c=p_c, o_x=o_x, o_c=o_c, y_true=y).items()}


def build_model(shape: Shape, decoder: str = "ca", p: float = 0.5, seed: int = 1234, perturb: bool = True):
    """CARCA with the reference's construction order (scripts/training.py:165-172) and init, then
    (optionally) biases / LayerNorm parameters perturbed by N(0, 0.1) so zero-bias bugs cannot hide."""
    from . import carca as M

    torch.manual_seed(seed)
    enc = M.IdentityEncoding()
    emb = M.AllEmbedding(shape.n_items, shape.d, shape.g, shape.n_ctx, shape.n_attrs, enc)
    blocks = nn.ModuleList([M.SelfAttentionBlock(shape.d, shape.n_heads, p, True) for _ in range(shape.n_blocks)])
    dec = M.CrossAttentionBlock(shape.d, shape.n_heads, p, True) if decoder == "ca" else M.DotProduct()
    model = M.CARCA(d=shape.d, p=p, emb=emb, enc=blocks, dec=dec)
    if perturb:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for name, prm in model.named_parameters():
                if name.endswith("bias") or ".norm" in name or name.startswith("norm."):
                    prm.add_(0.1 * torch.randn(prm.shape, generator=g))
            model.embeds.items_embed.weight[0].zero_()
    return model
