"""Full-catalog candidate scoring with the item table sharded over ranks (BASELINE.json configs[3],
SURVEY.md §8e).

The reference scores 1 positive + 100 sampled negatives (src/data.py:140-192) and has no
full-catalog mode; its definition here is the reference API itself applied to all items: eval-mode
`model.forward(profile, targets=[chunk_1, ..., chunk_n])` over candidate chunks (every candidate is
scored independently, src/carca.py:339-340,362, and chunks concatenate, :431), every candidate taking
the positive's context (src/data.py:185), followed by compute_HR / compute_NDCG (src/train.py:15-32)
with the positive's column labelled.

Sharding: rank r owns the contiguous item ids [lo_r, hi_r) of [1, n_items).  Every rank encodes the
users it is given (the encoder is ~1% of the work at 57K candidates, so recomputing it per rank is
cheaper than exchanging K/V), scores them against its own shard, and counts per user how many of its
items a stable descending sort would place before the positive.  The ONE exchange of the path is the
all-reduce of those [B] int32 counts; rank < k gives the hit, 1/log2(rank+2) the NDCG term.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
from torch import Tensor

from . import _native as N
from .ops import as_f32, as_ids


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of the real item ids [1, n_items) owned by `rank` (id 0 is padding)."""
    n = n_items - 1
    return 1 + (n * rank) // world, 1 + (n * (rank + 1)) // world


def _table_only(p_a) -> None:
    if isinstance(p_a, Tensor):
        raise ValueError("full-catalog scoring needs device-resident item attributes (ItemAttrTable / "
                         "set_attr_table), not the dense per-position [B, N, A] tensor")


def score_items(model, profile, ctx_user: Tensor, item_lo: int, item_hi: int, chunk: int = 4096) -> Tensor:
    """Eval-mode scores of items [item_lo, item_hi) for every user -> [B, item_hi - item_lo].
    Fused kernels when the shape allows (one launch, no candidate tensors), else the per-op path
    over candidate chunks through CARCA.forward."""
    from . import fused

    p_x, p_a, p_c = profile
    _table_only(p_a)
    n = item_hi - item_lo
    B = p_x.shape[0]
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            mode = model._fused_eval_mode((p_x, p_a, p_c), [(p_x, p_a, p_c)])
            if mode == "tc":
                return fused.forward_catalog(model, profile, ctx_user, item_lo, n)
            if mode is not None:        # packed-rows pipeline (long windows, d = 256, bf16): same catalog mode
                return fused.forward_rows(model, profile, [], precision=mode[5:], cat_lo=item_lo, n_cand=n,
                                          ctx_user=ctx_user)
            out = torch.empty((B, n), dtype=torch.float32, device=p_x.device)
            for lo in range(item_lo, item_hi, chunk):
                hi = min(item_hi, lo + chunk)
                ids = torch.arange(lo, hi, dtype=torch.int32, device=p_x.device).unsqueeze(0).expand(B, hi - lo)
                ctx = as_f32(ctx_user).unsqueeze(1).expand(B, hi - lo, ctx_user.shape[-1])
                out[:, lo - item_lo:hi - item_lo] = model.forward(profile, [(ids.contiguous(), p_a, ctx.contiguous())])
            return out
    finally:
        model.train(was_training)


def catalog_ranks(model, profile, pos_item: Tensor, pos_ctx: Tensor, group=None,
                  shard: Optional[Tuple[int, int]] = None, user_chunk: int = 2048, reduce: bool = True,
                  use_tc: bool = True) -> Tensor:
    """Rank (0 = best) of each user's positive item among ALL items, int32 [B] on every rank.
    With torch.distributed initialised the item table is sharded over the group's ranks (`shard` overrides this
    rank's id range; `reduce=False` returns the local counts without the all-reduce; `use_tc=False` keeps the
    score-matrix path where the tensor-core catalog kernel would apply)."""
    N.require_device(pos_item, pos_ctx)
    emb = model.embeds
    n_items = emb.items_embed.weight.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard if shard is not None else shard_bounds(n_items, rank, world)
    p_x, p_a, p_c = profile
    _table_only(p_a)
    pos_item, pos_ctx = as_ids(pos_item), as_f32(pos_ctx)
    B = p_x.shape[0]
    counts = torch.zeros(B, dtype=torch.int32, device=p_x.device)
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            from . import fused

            if use_tc and model.use_fused_eval and N.is_device_tensor(p_x) and not isinstance(p_a, Tensor) and \
                    (p_a is not None or emb.attr_table is not None) and hasattr(emb, "joint_embed") and \
                    fused.catalog_tc_supported(model, p_x.shape[1], p_c.shape[-1]):
                # tensor-core path: no [users, shard] score matrix at all
                for u0 in range(0, B, 4096):
                    u1 = min(B, u0 + 4096)
                    fused.catalog_counts(model, (p_x[u0:u1], p_a, p_c[u0:u1]), pos_item[u0:u1], pos_ctx[u0:u1], lo, hi,
                                         counts[u0:u1])
                user_chunk = 0
            for u0 in (range(0, B, user_chunk) if user_chunk else ()):      # bounds the [users, shard] score matrix
                u1 = min(B, u0 + user_chunk)
                prof = (p_x[u0:u1], p_a, p_c[u0:u1])
                ctx = pos_ctx[u0:u1].contiguous()
                pos = pos_item[u0:u1].contiguous()
                # the positive's own score, computed by the same kernels as every other candidate
                y_pos = model.forward(prof, [(pos.unsqueeze(1), p_a, ctx.unsqueeze(1))])[:, 0].contiguous()
                y = score_items(model, prof, ctx, lo, hi)
                N.call("carca_catalog_rank_count", N.i32p(counts[u0:u1]), N.f32p(y), y.stride(0), N.f32p(y_pos),
                       N.i32p(pos), int(lo), u1 - u0, hi - lo, N.stream())
    finally:
        model.train(was_training)
    if world > 1 and reduce:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)      # the one exchange of the path
    return counts


def catalog_metrics(model, profile, pos_item: Tensor, pos_ctx: Tensor, k: int, group=None) -> Tuple[float, float, Tensor]:
    """(HR@k, NDCG@k, ranks) of the positives against the full catalog, averaged over the users given."""
    ranks = catalog_ranks(model, profile, pos_item, pos_ctx, group=group)
    r = ranks.cpu().numpy().astype("float64")           # the one device -> host read
    hit = r < k
    hr = float(hit.mean()) if r.size else 0.0
    ndcg = float((hit / np.log2(r + 2.0)).mean()) if r.size else 0.0
    return hr, ndcg, ranks
