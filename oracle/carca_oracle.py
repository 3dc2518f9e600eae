"""TEST INFRASTRUCTURE — CPU oracle for the CARCA hot path.  Not a product path.

A functional torch-fp32 restatement of the reference's forward pass, loss and
ranking metrics (r-papso/carca-replication, `src/carca.py`, `src/train.py`,
`src/utils.py`).  Gradients come from torch autograd over these functions.
It keeps the reference's ATen op sequence (dense `cat`, `addmm`, head
split/cat copies, two full sorts ...) so that timing it on the host cores is a
fair stand-in for the reference's CPU path (`bench.py --impl reference`,
`cpu_baseline.kind == "port"`).

Parity pin: `tests/golden/*.npz` hold inputs, weights, outputs and gradients
produced by importing the *real* reference from /root/reference in the build
container (`tests/golden/make_golden.py`, committed).  `tests/test_oracle_golden.py`
checks this file against every one of them, so the oracle is pinned to the
reference itself, not only to its author's reading of it.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Weights are passed as a flat dict with the reference's `state_dict` names
(SURVEY.md §8b): `embeds.items_embed.weight`, `embeds.feats_embed.{weight,bias}`,
`embeds.joint_embed.{weight,bias}`, `encoder.{i}.norm1/norm2.{weight,bias}`,
`encoder.{i}.attn.{WQ,WK,WV}.{weight,bias}`, `encoder.{i}.ffn_1/ffn_2.{weight,bias}`,
`norm.{weight,bias}`, `decoder.attn.{WQ,WK,WV}.{weight,bias}`, `decoder.ffn.{weight,bias}`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import philox

Tensor = torch.Tensor

NEG_FILL = -(2 ** 32) + 1.0  # src/carca.py:251

# dropout sites (shared numbering with carca_replication_b200/ops.py)
SITE_EMBED = 0
SITE_DECODER_ATTN = 1000


def site_attn(block: int) -> int:
    return 1 + 3 * block


def site_ffn1(block: int) -> int:
    return 2 + 3 * block


def site_ffn2(block: int) -> int:
    return 3 + 3 * block


@dataclass
class OracleConfig:
    d: int
    n_heads: int
    n_blocks: int
    decoder: str = "dot"          # "dot" | "ca"
    residual_sa: bool = True
    residual_ca: bool = True
    p_drop: float = 0.0
    seed: int = 0                 # dropout stream seed (ours, see oracle/philox.py)
    learnable_pos: bool = False   # embeds.enc.encoding.weight present
    sinus_pos: bool = False       # embeds.enc.pe present
    embedding: str = "all"        # "all" | "id" | "mlpid" | "attr" | "attrctx"  (scripts/training.py:78-89)
    gamma: float = 0.9            # WeightedDotProduct (decoder "wdot" / "wdot_norm")


class Dropper:
    """Applies the kernels' Philox keep/drop decisions with torch arithmetic."""

    def __init__(self, p: float, seed: int, active: bool):
        self.p, self.seed, self.active = float(p), int(seed), bool(active)

    def __call__(self, t: Tensor, site: int, attn_heads: int = 0) -> Tensor:
        if not self.active or self.p == 0.0:
            return t
        keep = philox.keep_mask(t.numel(), self.p, self.seed, site)
        scale = np.float32(1.0) / (np.float32(1.0) - np.float32(self.p))
        if attn_heads:
            # kernels index attention weights as ((b*H + h)*Lq + i)*Lk + j; the
            # reference stacks heads on the batch dim, i.e. [(h*B + b), i, j].
            HB, Lq, Lk = t.shape
            B = HB // attn_heads
            keep = keep.reshape(B, attn_heads, Lq, Lk).transpose(1, 0, 2, 3).reshape(HB, Lq, Lk)
        else:
            keep = keep.reshape(tuple(t.shape))
        m = torch.from_numpy(keep.astype(np.float32) * scale)
        return t * m


def padding_mask(ids: Tensor) -> Tensor:
    """src/utils.py:6-7 — 1.0 where id != 0, else 0.0 (fp32)."""
    return torch.where(ids == 0.0, 0.0, 1.0)


# ----------------------------------------------------------------------------- embedding
def embed_all(sd: Dict[str, Tensor], ids: Tensor, attrs: Tensor, ctx: Tensor, mask: Tensor,
              is_target: bool, cfg: OracleConfig) -> Tensor:
    """AllEmbedding.forward, src/carca.py:85-95.

    q = Lin_f([a | c]);  z = sqrt(d) * E[ids];  e = Lin_j([z | q]);
    (+ positional encoding for the profile only);  e *= mask.
    """
    d = cfg.d
    if cfg.embedding == "all":
        feats = torch.cat((attrs, ctx), dim=-1)                               # :86
        q = F.linear(feats, sd["embeds.feats_embed.weight"], sd["embeds.feats_embed.bias"])
        z = F.embedding(ids.long(), sd["embeds.items_embed.weight"])         # :87
        z = z * (d ** 0.5)                                                    # :88
        e = F.linear(torch.cat((z, q), dim=-1), sd["embeds.joint_embed.weight"],
                     sd["embeds.joint_embed.bias"])                           # :89
    elif cfg.embedding == "attrctx":                                          # AttrCtxEmbedding.forward, :112-114
        q = F.linear(torch.cat((attrs, ctx), dim=-1), sd["embeds.feats_embed.weight"], sd["embeds.feats_embed.bias"])
        e = F.linear(q, sd["embeds.joint_embed.weight"], sd["embeds.joint_embed.bias"])
    elif cfg.embedding == "attr":                                             # AttrEmbedding.forward, :139-141
        q = F.linear(attrs, sd["embeds.feats_embed.weight"], sd["embeds.feats_embed.bias"])
        e = F.linear(q, sd["embeds.joint_embed.weight"], sd["embeds.joint_embed.bias"])
    elif cfg.embedding == "id":                                               # IdEmbedding.forward, :163-165
        e = F.embedding(ids.long(), sd["embeds.items_embed.weight"]) * (d ** 0.5)
    elif cfg.embedding == "mlpid":                                            # MLPIdEmbedding.forward, :189-192
        z = F.embedding(ids.long(), sd["embeds.items_embed.weight"]) * (d ** 0.5)
        e = F.linear(z, sd["embeds.feats_embed.weight"], sd["embeds.feats_embed.bias"])
    else:
        raise ValueError(f"Unknown embedding type: {cfg.embedding}")
    if not is_target:                                                         # :91-92
        if cfg.learnable_pos:                                                 # :25-31
            e = e + sd["embeds.enc.encoding.weight"][: e.size(1)].unsqueeze(0)
        elif cfg.sinus_pos:                                                   # :54-60
            e = e + sd["embeds.enc.pe"][:, : e.size(1), :]
    return e * mask.unsqueeze(2)                                              # :94


# ----------------------------------------------------------------------------- attention
def multi_head_attention(sd: Dict[str, Tensor], prefix: str, query: Tensor, key: Tensor, value: Tensor,
                         q_mask: Tensor, k_mask: Tensor, n_heads: int, causal: Optional[int],
                         drop: Dropper, site: int) -> Tuple[Tensor, Tensor]:
    """MultiHeadAttention.forward, src/carca.py:228-265. Returns (weights, out)."""
    d = query.shape[-1]
    dh = d // n_heads
    Q = F.linear(query, sd[prefix + "WQ.weight"], sd[prefix + "WQ.bias"])     # :238
    K = F.linear(key, sd[prefix + "WK.weight"], sd[prefix + "WK.bias"])       # :239
    V = F.linear(value, sd[prefix + "WV.weight"], sd[prefix + "WV.bias"])     # :240
    Q = torch.cat(torch.split(Q, dh, dim=2), dim=0)                           # :242 head-major batch
    K = torch.cat(torch.split(K, dh, dim=2), dim=0)
    V = torch.cat(torch.split(V, dh, dim=2), dim=0)
    allow = torch.bmm(q_mask.unsqueeze(2), k_mask.unsqueeze(1)).bool()        # :246-248
    allow = torch.tile(allow, (n_heads, 1, 1))                                # :249
    if causal is not None:
        allow = torch.tril(allow, diagonal=causal)                            # :250
    add = torch.where(allow, 0.0, NEG_FILL)                                   # :251
    w = torch.baddbmm(add, Q, K.transpose(1, 2))                              # :253 mask BEFORE scaling
    w = w / (dh ** 0.5)                                                       # :254
    w = torch.softmax(w, dim=-1)                                              # :255
    w = w * allow                                                             # :256 dead rows -> exactly 0
    out = torch.bmm(drop(w, site, attn_heads=n_heads), V)                     # :258-259
    out = torch.cat(torch.split(out, out.shape[0] // n_heads, dim=0), dim=2)  # :260
    return w, out


def self_attention_block(sd: Dict[str, Tensor], i: int, x: Tensor, mask: Tensor, cfg: OracleConfig,
                         drop: Dropper) -> Tensor:
    """SelfAttentionBlock.forward, src/carca.py:297-318."""
    pre = f"encoder.{i}."
    d = cfg.d
    q = F.layer_norm(x, (d,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5)   # :298
    _, s = multi_head_attention(sd, pre + "attn.", q, x, x, mask, mask, cfg.n_heads, 0,
                                drop, site_attn(i))                                      # :299
    if cfg.residual_sa:
        s = s + q                                                                        # :302
    s = F.layer_norm(s, (d,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)   # :304
    f = s.transpose(1, 2).contiguous()                                                   # :305
    f = F.conv1d(f, sd[pre + "ffn_1.weight"], sd[pre + "ffn_1.bias"])                    # :307
    f = F.leaky_relu(f, 0.01)                                                            # :308
    f = drop(f.transpose(1, 2).contiguous(), site_ffn1(i)).transpose(1, 2).contiguous()  # :309
    f = F.conv1d(f, sd[pre + "ffn_2.weight"], sd[pre + "ffn_2.bias"])                    # :311
    f = f.transpose(1, 2).contiguous()                                                   # :313
    f = drop(f, site_ffn2(i))                                                            # :312
    if cfg.residual_sa:
        f = f + s                                                                        # :316
    return f


def cross_attention_scores(sd: Dict[str, Tensor], o: Tensor, o_mask: Tensor, p: Tensor, p_mask: Tensor,
                           cfg: OracleConfig, training: bool, drop: Dropper, site: int) -> Tensor:
    """CrossAttentionBlock.forward, src/carca.py:338-349.

    Deviation kept out of the oracle on purpose: the reference `squeeze()`s every
    unit dim (:346, collapses B==1); here only the trailing one is squeezed so the
    result is always [B, T].
    """
    causal = -1 if training else None                                         # :339
    _, s = multi_head_attention(sd, "decoder.attn.", o, p, p, o_mask, p_mask, cfg.n_heads, causal,
                                drop, site)                                   # :340
    if cfg.residual_ca:
        s = s + o                                                             # :343
    y = F.linear(s, sd["decoder.ffn.weight"], sd["decoder.ffn.bias"])         # :345
    return torch.sigmoid(y.squeeze(-1))                                       # :346-347


def dot_scores(o: Tensor, p: Tensor, training: bool) -> Tensor:
    """DotProduct.forward, src/carca.py:358-365."""
    if training:
        y = torch.sum(p * o, dim=-1)                                          # :360
    else:
        y = torch.sum(p[:, -1:, :] * o, dim=-1)                               # :362
    return torch.sigmoid(y)


def weighted_dot_scores(o: Tensor, p: Tensor, training: bool, gamma: float, normalize: bool) -> Tensor:
    """WeightedDotProduct.forward, src/carca.py:377-395 (the [L,L,1] weight of :373-374 rebuilt here)."""
    L = p.size(1)
    W = (gamma ** torch.arange(0, L).unsqueeze(0).repeat(L, 1)).tril().unsqueeze(-1)   # :373-374
    pw = p.unsqueeze(2).repeat(1, 1, L, 1)                                    # :378
    p = torch.sum(pw * W, dim=2)                                              # :379
    if normalize:
        p = F.normalize(p, dim=2)                                             # :382
        o = F.normalize(o, dim=2)                                             # :383
    y = torch.sum(p * o, dim=-1) if training else torch.sum(p[:, -1:, :] * o, dim=-1)   # :385-388
    return (y + 1.0) / 2.0 if normalize else torch.sigmoid(y)                 # :390-393


# ----------------------------------------------------------------------------- model
def encode_profile(sd, cfg: OracleConfig, profile, training: bool, drop: Dropper) -> Tuple[Tensor, Tensor]:
    """src/carca.py:412-421 — mask, embed, dropout, blocks, final LayerNorm."""
    p_x, p_a, p_c = profile
    p_mask = padding_mask(p_x)                                                # :413
    p_e = embed_all(sd, p_x, p_a, p_c, p_mask, False, cfg)                    # :415
    p_e = drop(p_e, SITE_EMBED)                                               # :416
    for i in range(cfg.n_blocks):                                             # :418-419
        p_e = self_attention_block(sd, i, p_e, p_mask, cfg, drop)
    p_e = F.layer_norm(p_e, (cfg.d,), sd["norm.weight"], sd["norm.bias"], 1e-5)  # :421
    return p_e, p_mask


def carca_forward(sd: Dict[str, Tensor], cfg: OracleConfig, profile: Tuple[Tensor, Tensor, Tensor],
                  targets: Sequence[Tuple[Tensor, Tensor, Tensor]], training: bool) -> Tensor:
    """CARCA.forward, src/carca.py:411-431 -> probabilities [B, sum(T)]."""
    drop = Dropper(cfg.p_drop, cfg.seed, training)
    p_e, p_mask = encode_profile(sd, cfg, profile, training, drop)
    ys: List[Tensor] = []
    for t_idx, (o_x, o_a, o_c) in enumerate(targets):                         # :424
        o_mask = padding_mask(o_x)                                            # :425
        o_e = embed_all(sd, o_x, o_a, o_c, o_mask, True, cfg)                 # :426
        if cfg.decoder == "ca":
            y = cross_attention_scores(sd, o_e, o_mask, p_e, p_mask, cfg, training, drop,
                                       SITE_DECODER_ATTN + t_idx)
        elif cfg.decoder == "dot":
            y = dot_scores(o_e, p_e, training)
        elif cfg.decoder in ("wdot", "wdot_norm"):
            y = weighted_dot_scores(o_e, p_e, training, cfg.gamma, cfg.decoder == "wdot_norm")
        else:
            raise ValueError(f"Unknown decoder type: {cfg.decoder}")
        ys.append(y)                                                          # :428-429
    return torch.cat(ys, dim=-1)                                              # :431


# ----------------------------------------------------------------------------- loss / metrics
def masked_bce(y_pred: Tensor, y_true: Tensor, mask: Tensor, eps: float = 1e-8) -> Tensor:
    """BinaryCrossEntropy.forward, src/carca.py:441-444 (eps on probabilities)."""
    ell = -(y_true * torch.log(y_pred + eps) + (1.0 - y_true) * torch.log(1.0 - y_pred + eps))
    return torch.sum(ell * mask) / torch.sum(mask)


def hit_count(y_pred: Tensor, y_true: Tensor, k: int) -> float:
    """compute_HR, src/train.py:15-21."""
    _, order = torch.sort(y_pred, descending=True)
    top = torch.gather(y_true, dim=1, index=order)[:, :k]
    return torch.sum(top).item()


def ndcg_sum(y_pred: Tensor, y_true: Tensor, k: int) -> float:
    """compute_NDCG, src/train.py:24-32."""
    _, order = torch.sort(y_pred, descending=True)
    top = torch.gather(y_true, dim=1, index=order)[:, :k]
    ranks = torch.nonzero(top)[:, 1]
    return torch.sum(1.0 / torch.log2(ranks + 2)).item()


def train_step_targets(o_x: Tensor, o_a: Tensor, o_c: Tensor):
    """src/train.py:86-88 — split the [B, 2L] target tensors into (pos, neg) halves."""
    h = o_x.shape[1] // 2
    return [(o_x[:, :h], o_a[:, :h], o_c[:, :h]), (o_x[:, h:], o_a[:, h:], o_c[:, h:])]


def eval_batch(sd, cfg: OracleConfig, batch, k: int = 10) -> Tuple[float, float, float, int]:
    """Body of evaluate(), src/train.py:41-51 -> (hits, ndcg_sum, loss, n_users)."""
    p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch
    with torch.no_grad():
        y = carca_forward(sd, cfg, (p_x, p_a, p_c), [(o_x, o_a, o_c)], training=False)
        loss = masked_bce(y, y_true, padding_mask(o_x)).item()
        return hit_count(y, y_true, k), ndcg_sum(y, y_true, k), loss, int(y_true.shape[0])


def train_batch_loss(sd, cfg: OracleConfig, batch) -> Tensor:
    """Forward half of the train step, src/train.py:84-93 -> scalar loss (autograd-enabled)."""
    p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch
    y = carca_forward(sd, cfg, (p_x, p_a, p_c), train_step_targets(o_x, o_a, o_c), training=True)
    return masked_bce(y, y_true, padding_mask(o_x))
