"""CARCA model classes backed by the sm_100a kernels.

Class names, constructor signatures, `forward` signatures and `state_dict` keys/shapes follow the
reference's `src/carca.py` so that `scripts/training.py:165-174` and `src/train.py` run against
them unchanged and checkpoints move both ways.  The torch `nn.Embedding / nn.Linear / nn.Conv1d /
nn.LayerNorm` objects below are parameter containers only (they give the reference's key names and
initialisation); their `forward` is never called — every op runs in libcarca_b200.so via `ops`.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .abstract import Decoder, Embedding, Encoder, Encoding, Model
from .attrs import ItemAttrTable
from .utils import get_mask


def _invalidate_plans() -> None:
    """.to() / .cuda() / .float() replace parameter tensors: cached inference plans (fused.py) must be rebuilt."""
    from . import fused

    fused.invalidate_plans()


def _xavier(layer: nn.Module, zero_bias: bool = True) -> nn.Module:
    nn.init.xavier_uniform_(layer.weight)
    if zero_bias and getattr(layer, "bias", None) is not None:
        nn.init.zeros_(layer.bias)
    return layer


# ------------------------------------------------------------------------------- positional encodings
class IdentityEncoding(Encoding):
    """No positional signal — the script default (src/carca.py:34-39)."""

    def forward(self, x: Tensor) -> Tensor:
        return x

    def table(self, seq_len: int) -> Optional[Tensor]:
        return None


class LearnableEncoding(Encoding):
    """Learned [max_len, d] table added to profile embeddings (src/carca.py:15-31).
    Inside AllEmbedding the add is fused into the embedding kernel's epilogue."""

    def __init__(self, d: int, max_len: int):
        super().__init__()
        self.max_len = max_len
        self.encoding = _xavier(nn.Embedding(max_len, d))

    def table(self, seq_len: int) -> Tensor:
        return self.encoding.weight

    def forward(self, x: Tensor) -> Tensor:
        return x + self.encoding.weight[: x.size(1)].unsqueeze(0)


class PositionalEncoding(Encoding):
    """Fixed sinusoidal table, buffer `pe` [1, max_len, d] (src/carca.py:43-60)."""

    def __init__(self, d_model: int, max_len: int):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(pos * freq)
        pe[0, :, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", pe)

    def table(self, seq_len: int) -> Tensor:
        return self.pe[0]

    def forward(self, x: Tensor) -> Tensor:
        return x + self.pe[:, : x.size(1), :]


# ------------------------------------------------------------------------------- embedding
class AllEmbedding(Embedding):
    """Item id + attribute + context embedding (src/carca.py:66-95), one fused op.

    `a` may be the dense [B, N, A] tensor of the reference API, an `ItemAttrTable`, or None when a
    table was attached with `set_attr_table` (device-resident attributes, ids + context only).
    """

    def __init__(self, n_items: int, d: int, g: int, n_ctx: int, n_attrs: int, enc: Encoding):
        super().__init__()
        self.d = d
        self.enc = enc
        self.items_embed = _xavier(nn.Embedding(n_items, d, padding_idx=0))
        self.feats_embed = _xavier(nn.Linear(n_ctx + n_attrs, g))
        self.joint_embed = _xavier(nn.Linear(g + d, d))
        with torch.no_grad():
            self.items_embed.weight[0].zero_()           # padding row (src/carca.py:81)
        self._attr_table: List[ItemAttrTable] = []       # kept out of _modules / state_dict

    def set_attr_table(self, table: Optional[ItemAttrTable]) -> "AllEmbedding":
        self._attr_table = [] if table is None else [table]
        return self

    @property
    def attr_table(self) -> Optional[ItemAttrTable]:
        return self._attr_table[0] if self._attr_table else None

    def _apply(self, fn, *args, **kwargs):
        _invalidate_plans()
        self.__dict__.pop("_fold_cache", None)
        for t in self._attr_table:                        # follow .to(device) / .cuda()
            t._apply(fn, *args, **kwargs)
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        if mode:
            self.__dict__.pop("_fold_cache", None)        # weights are about to change: drop the derived table
        return super().train(mode)

    # ---- inference through a folded item table (per-op path; the fused kernels keep their own plan) -------------
    use_folded_eval = True   # class default; set False on an instance to always run the unfolded op
    _FOLD_CHUNK = 16384      # item ids per call when the table is built

    def _folded(self, table: ItemAttrTable):
        """(T [n_items, d], McT [C, d]) for the current weights: T[i] = joint(sqrt(d) E[i], feats(attrs[i], 0)) is this
        module's own unfolded op applied to every item id with a zero context, McT = (Wj[:, d:] Wf[:, A:])^T.  Exact
        re-association of src/carca.py:86-89 (no nonlinearity between the two linears); rebuilt when a parameter
        or the attribute table changes."""
        ps = (self.items_embed.weight, self.feats_embed.weight, self.feats_embed.bias, self.joint_embed.weight,
              self.joint_embed.bias)
        from . import fused

        key = tuple((p.data_ptr(), p._version) for p in ps) + (id(table), fused.weights_epoch())
        hit = getattr(self, "_fold_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        E, Wf, bf, Wj, bj = ps
        n_items, d = E.shape
        Cn = Wf.shape[1] - table.n_attrs
        dev = E.device
        T = torch.empty((n_items, d), dtype=torch.float32, device=dev)
        for lo in range(0, n_items, self._FOLD_CHUNK):
            n = min(self._FOLD_CHUNK, n_items - lo)
            ids = torch.arange(lo, lo + n, dtype=torch.int32, device=dev).reshape(1, n)
            T[lo:lo + n] = ops.EmbedFn.apply(ids, torch.zeros((1, n, Cn), dtype=torch.float32, device=dev),
                                             torch.ones((1, n), dtype=torch.float32, device=dev), None, E, Wf, bf, Wj, bj,
                                             None, table, True)[0]
        A = table.n_attrs
        McT = ops.LinearFn.apply(Wf[:, A:].t().contiguous(), Wj[:, d:].contiguous(),
                                 torch.zeros(d, dtype=torch.float32, device=dev)) if Cn > 0 else \
            torch.zeros((0, d), dtype=torch.float32, device=dev)
        self._fold_cache = (key, T, McT.contiguous())
        return T, self._fold_cache[2]

    def __getstate__(self):
        state = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        state = dict(state)
        state.pop("_fold_cache", None)               # derived from the weights: never pickled with the module
        return state

    def _forward_folded(self, x: Tensor, c: Tensor, mask: Tensor, pos: Optional[Tensor], table: ItemAttrTable) -> Tensor:
        from . import _native as N_
        from .ops import as_f32, as_ids

        T, McT = self._folded(table)
        x, c, mask = as_ids(x), as_f32(c), as_f32(mask)
        n_rows, n_cols = x.shape
        d = T.shape[1]
        out = torch.empty((n_rows, n_cols, d), dtype=torch.float32, device=T.device)
        posc = None if pos is None else pos.contiguous()
        N_.call("carca_embed_folded_fwd", N_.f32p(out), N_.f32p(T), N_.f32p(McT), N_.i32p(x), N_.f32p(c),
                None if posc is None else N_.f32p(posc), N_.f32p(mask), n_rows, n_cols, d, int(c.shape[-1]), N_.stream())
        return out

    def forward(self, x: Tensor, a: Union[Tensor, ItemAttrTable, None], c: Tensor, mask: Tensor,
                target: bool) -> Tensor:
        table = a if isinstance(a, ItemAttrTable) else (self.attr_table if a is None else None)
        dense = a if isinstance(a, Tensor) else None
        pos, foreign_enc = None, False
        if not target:
            if hasattr(self.enc, "table"):
                pos = self.enc.table(x.shape[1])
            else:
                foreign_enc = True                        # a user-supplied Encoding plug-in
        kernel_mask = torch.ones_like(mask) if foreign_enc else mask
        if (self.use_folded_eval and not self.training and not torch.is_grad_enabled() and table is not None
                and dense is None and self.d % 4 == 0 and x.dim() == 2):
            e = self._forward_folded(x, c, kernel_mask, pos, table)
            if foreign_enc:
                e = self.enc.forward(e) * mask.unsqueeze(2)
            return e
        e = ops.EmbedFn.apply(x, c, kernel_mask, dense, self.items_embed.weight, self.feats_embed.weight,
                              self.feats_embed.bias, self.joint_embed.weight, self.joint_embed.bias, pos, table,
                              bool(target))
        if foreign_enc:
            e = self.enc.forward(e) * mask.unsqueeze(2)
        return e


class _VariantEmbedding(Embedding):
    """Shared tail of the ablation embeddings: positional encoding of profile rows, then the mask."""

    def _finish(self, e: Tensor, mask: Tensor, target: bool) -> Tensor:
        pos = None
        if not target:
            if hasattr(self.enc, "table"):
                pos = self.enc.table(e.shape[1])
            else:                                         # a user-supplied Encoding plug-in
                return self.enc.forward(ops.PosMaskFn.apply(e, None, torch.ones_like(mask))) * mask.unsqueeze(2)
        return ops.PosMaskFn.apply(e, pos, mask)


class _AttrSourceMixin:
    def set_attr_table(self, table: Optional[ItemAttrTable]):
        self._attr_table = [] if table is None else [table]
        return self

    @property
    def attr_table(self) -> Optional[ItemAttrTable]:
        return self._attr_table[0] if self._attr_table else None

    def _apply(self, fn, *args, **kwargs):
        for t in self._attr_table:
            t._apply(fn, *args, **kwargs)
        return super()._apply(fn, *args, **kwargs)

    def _split(self, a):
        table = a if isinstance(a, ItemAttrTable) else (self.attr_table if a is None else None)
        return table, (a if isinstance(a, Tensor) else None)


class AttrCtxEmbedding(_AttrSourceMixin, _VariantEmbedding):
    """Attributes + context only: Lin_{g->d}(Lin_{A+C->g}([a | c])) (src/carca.py:98-122)."""

    def __init__(self, d: int, g: int, n_ctx: int, n_attrs: int, enc: Encoding):
        super().__init__()
        self.d, self.enc = d, enc
        self.feats_embed = _xavier(nn.Linear(n_ctx + n_attrs, g))
        self.joint_embed = _xavier(nn.Linear(g, d))
        self._attr_table: List[ItemAttrTable] = []

    def forward(self, x: Tensor, a, c: Tensor, mask: Tensor, target: bool) -> Tensor:
        table, dense = self._split(a)
        q = ops.FeatsFn.apply(x, c, dense, self.feats_embed.weight, self.feats_embed.bias, table)
        e = ops.LinearFn.apply(q, self.joint_embed.weight, self.joint_embed.bias)
        return self._finish(e, mask, target)


class AttrEmbedding(_AttrSourceMixin, _VariantEmbedding):
    """Attributes only: Lin_{g->d}(Lin_{A->g}(a)) (src/carca.py:125-149)."""

    def __init__(self, d: int, g: int, n_attrs: int, enc: Encoding):
        super().__init__()
        self.d, self.enc = d, enc
        self.feats_embed = _xavier(nn.Linear(n_attrs, g))
        self.joint_embed = _xavier(nn.Linear(g, d))
        self._attr_table: List[ItemAttrTable] = []

    def forward(self, x: Tensor, a, c: Tensor, mask: Tensor, target: bool) -> Tensor:
        table, dense = self._split(a)
        q = ops.FeatsFn.apply(x, None, dense, self.feats_embed.weight, self.feats_embed.bias, table)
        e = ops.LinearFn.apply(q, self.joint_embed.weight, self.joint_embed.bias)
        return self._finish(e, mask, target)


class IdEmbedding(_VariantEmbedding):
    """Item ids only: sqrt(d) * Emb[x] (src/carca.py:152-171)."""

    def __init__(self, n_items: int, d: int, enc: Encoding):
        super().__init__()
        self.d, self.enc = d, enc
        self.items_embed = _xavier(nn.Embedding(n_items, d, padding_idx=0))
        with torch.no_grad():
            self.items_embed.weight[0].zero_()

    def forward(self, x: Tensor, a, c: Tensor, mask: Tensor, target: bool) -> Tensor:
        e = ops.GatherRowsFn.apply(x, self.items_embed.weight, self.d ** 0.5)
        return self._finish(e, mask, target)


class MLPIdEmbedding(_VariantEmbedding):
    """Lin_{g->d}(sqrt(d) * Emb_g[x]) (src/carca.py:174-198)."""

    def __init__(self, n_items: int, d: int, g: int, enc: Encoding):
        super().__init__()
        self.d, self.enc = d, enc
        self.items_embed = _xavier(nn.Embedding(n_items, g, padding_idx=0))
        self.feats_embed = _xavier(nn.Linear(g, d))
        with torch.no_grad():
            self.items_embed.weight[0].zero_()

    def forward(self, x: Tensor, a, c: Tensor, mask: Tensor, target: bool) -> Tensor:
        z = ops.GatherRowsFn.apply(x, self.items_embed.weight, self.d ** 0.5)
        e = ops.LinearFn.apply(z, self.feats_embed.weight, self.feats_embed.bias)
        return self._finish(e, mask, target)


# ------------------------------------------------------------------------------- attention
class MultiHeadAttention(nn.Module):
    """Masked multi-head attention without output projection (src/carca.py:204-265)."""

    def __init__(self, embed_dim: int, num_heads: int, dropout: float):
        super().__init__()
        assert embed_dim % num_heads == 0.0, "Embedding dim must be divisible by number of heads"
        self.d = embed_dim
        self.H = num_heads
        self.WQ = _xavier(nn.Linear(embed_dim, embed_dim))
        self.WK = _xavier(nn.Linear(embed_dim, embed_dim))
        self.WV = _xavier(nn.Linear(embed_dim, embed_dim))
        self.softmax = nn.Softmax(dim=-1)        # kept for module-tree parity; unused
        self.dropout = nn.Dropout(p=dropout)
        self._site = ops.SITE_DECODER_ATTN

    def forward(self, query: Tensor, key: Tensor, value: Tensor, q_mask: Tensor, k_mask: Tensor,
                causal: int = None, return_w: bool = False):
        Q = ops.LinearFn.apply(query, self.WQ.weight, self.WQ.bias)
        K = ops.LinearFn.apply(key, self.WK.weight, self.WK.bias)
        V = ops.LinearFn.apply(value, self.WV.weight, self.WV.bias)
        p = self.dropout.p if self.training else 0.0
        out, w = ops.AttentionCoreFn.apply(Q, K, V, q_mask, k_mask, self.H, causal, p, ops.current_seed(),
                                           self._site, bool(return_w))
        if return_w:
            B, H, Lq, Lk = w.shape                      # reference layout: heads stacked on the batch dim
            return w.permute(1, 0, 2, 3).reshape(H * B, Lq, Lk), out
        return out


# ------------------------------------------------------------------------------- encoder / decoders
class SelfAttentionBlock(Encoder):
    """LN -> causal MHA (+LN(x)) -> LN -> pointwise FFN (+) (src/carca.py:272-318), one C call."""

    def __init__(self, d: int, H: int, p: float, residual: bool):
        super().__init__()
        self.residual = residual
        self.norm1 = nn.LayerNorm(normalized_shape=d)
        self.attn = MultiHeadAttention(embed_dim=d, num_heads=H, dropout=p)
        self.norm2 = nn.LayerNorm(normalized_shape=d)
        self.ffn_1 = _xavier(nn.Conv1d(in_channels=d, out_channels=d, kernel_size=1))
        self.lrelu = nn.LeakyReLU()
        self.dropout1 = nn.Dropout(p=p)
        self.ffn_2 = _xavier(nn.Conv1d(in_channels=d, out_channels=d, kernel_size=1))
        self.dropout2 = nn.Dropout(p=p)
        self._block_index = 0                            # set by CARCA: selects the dropout sites

    def _apply(self, fn, *args, **kwargs):
        _invalidate_plans()
        return super()._apply(fn, *args, **kwargs)

    def _params(self) -> Tuple[Tensor, ...]:
        a = self.attn
        return (self.norm1.weight, self.norm1.bias, a.WQ.weight, a.WQ.bias, a.WK.weight, a.WK.bias, a.WV.weight,
                a.WV.bias, self.norm2.weight, self.norm2.bias, self.ffn_1.weight, self.ffn_1.bias,
                self.ffn_2.weight, self.ffn_2.bias)

    def forward(self, x: Tensor, mask: Tensor) -> Tensor:
        p = self.dropout1.p if self.training else 0.0
        return ops.SABlockFn.apply(x, mask, self.attn.H, bool(self.residual), p, ops.current_seed(),
                                   self._block_index, *self._params())


class CrossAttentionBlock(Decoder):
    """Targets attend over the encoded profile, then Linear(d->1) + sigmoid (src/carca.py:322-349).
    Always returns [B, T] (the reference's squeeze() drops the batch dim when B == 1, :346)."""

    def __init__(self, d: int, H: int, p: float, residual: bool):
        super().__init__()
        self.residual = residual
        self.attn = MultiHeadAttention(embed_dim=d, num_heads=H, dropout=p)
        self.ffn = _xavier(nn.Linear(in_features=d, out_features=1))
        self.sig = nn.Sigmoid()
        self._site = ops.SITE_DECODER_ATTN               # + target index, set by CARCA per call

    def forward(self, o: Tensor, o_mask: Tensor, p: Tensor, p_mask: Tensor) -> Tensor:
        a = self.attn
        p_drop = a.dropout.p if self.training else 0.0
        return ops.CrossScoreFn.apply(o, o_mask, p, p_mask, a.H, bool(self.residual), bool(self.training), p_drop,
                                      ops.current_seed(), self._site, a.WQ.weight, a.WQ.bias, a.WK.weight,
                                      a.WK.bias, a.WV.weight, a.WV.bias, self.ffn.weight, self.ffn.bias)


class DotProduct(Decoder):
    """sigmoid(<profile, target>): position-wise in training, last profile position in eval
    (src/carca.py:352-365) — the script's default decoder."""

    def __init__(self) -> None:
        super().__init__()
        self.sig = nn.Sigmoid()

    def forward(self, o: Tensor, o_mask: Tensor, p: Tensor, p_mask: Tensor) -> Tensor:
        return ops.DotScoreFn.apply(o, p, bool(self.training))


class WeightedDotProduct(Decoder):
    """Dot product with profile position i scaled by sum_{j<=i} gamma^j, optional L2 normalisation of both
    sides (src/carca.py:368-395; the reference's [L,L] weight tensor only ever yields that per-position
    scalar).  `device` is accepted for signature compatibility (scripts/training.py:98)."""

    def __init__(self, gamma: float, seq_len: int, normalize: bool, device: str = "cuda"):
        super().__init__()
        self.gamma, self.seq_len, self.norm = float(gamma), int(seq_len), bool(normalize)
        self.sig = nn.Sigmoid()

    def forward(self, o: Tensor, o_mask: Tensor, p: Tensor, p_mask: Tensor) -> Tensor:
        return ops.WDotScoreFn.apply(o, p, bool(self.training), self.gamma, self.norm)


# ------------------------------------------------------------------------------- model
class CARCA(Model):
    """mask -> embed(profile) -> dropout -> encoder blocks -> LayerNorm -> per target tuple
    embed + decode -> concat (src/carca.py:401-431)."""

    def __init__(self, d: int, p: float, emb: Embedding, enc: Iterable[Encoder], dec: Decoder):
        super().__init__()
        self.embeds = emb
        self.dropout = nn.Dropout(p=p)
        self.encoder = enc
        self.norm = nn.LayerNorm(normalized_shape=d)
        self.decoder = dec
        for i, blk in enumerate(enc):
            if isinstance(blk, SelfAttentionBlock):
                blk._block_index = i

    def _apply(self, fn, *args, **kwargs):
        _invalidate_plans()
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_eval_graph_steps", None)         # captured CUDA graphs of evaluate(): never pickled / deep-copied
        return state

    def encode(self, profile) -> Tuple[Tensor, Tensor]:
        p_x, p_a, p_c = profile
        p_mask = get_mask(p_x)
        p_e = self.embeds.forward(p_x, p_a, p_c, p_mask, False)
        if self.training and self.dropout.p > 0.0:
            p_e = ops.DropoutFn.apply(p_e, self.dropout.p, ops.current_seed(), ops.SITE_EMBED)
        for block in self.encoder:
            p_e = block.forward(p_e, p_mask)
        p_e = ops.LayerNormFn.apply(p_e, self.norm.weight, self.norm.bias)
        return p_e, p_mask

    use_fused_eval = True   # class default; set False on an instance to force the per-op path

    def _fused_eval_mode(self, profile, targets) -> Optional[str]:
        """Which whole-model inference path serves this call (eval mode, no autograd, device-resident attributes):
        "tc"   — the one-kernel fp32 tensor-core forward (d = 64, windows of at most one 64-row bin: L <= 64),
        "rows_fp32" — the packed-rows pipeline with 3xTF32 GEMMs (fp32 contract; any number of valid positions per
                 user, L <= 256, d in {32, 64, 128, 256}): longer windows and the widths the fused kernel does not cover,
        "rows_bf16" — the same pipeline in bf16 (eval_dtype == "bf16"),
        None   — the per-op kernels.  Nothing is decided by reading the device: no host sync on this path."""
        if self.training or torch.is_grad_enabled() or not self.use_fused_eval or not targets:
            return None
        emb = self.embeds
        if not isinstance(emb, AllEmbedding):
            return None
        for a in [profile[1]] + [t[1] for t in targets]:
            if isinstance(a, Tensor) or (a is None and emb.attr_table is None):
                return None
        if any(isinstance(t[1], ItemAttrTable) and t[1] is not (profile[1] or emb.attr_table) for t in targets):
            return None
        from . import fused
        from . import _native
        if not _native.is_device_tensor(profile[0]):
            return None
        L, n_ctx = profile[0].shape[1], profile[2].shape[-1]
        if self.eval_dtype == "bf16":
            if not fused.rows_supported(self, L, n_ctx, "bf16"):
                raise RuntimeError("CARCA.eval_dtype == 'bf16' needs AllEmbedding with a device attribute table, stock "
                                   "blocks / decoder, d in {64, 256}, head width 32 or 64, C <= 8, L <= 256")
            return "rows_bf16"
        if self.force_eval_path is not None:            # tests / benchmarks: "tc", "rows_fp32"
            return self.force_eval_path
        if L <= fused.BIN_ROWS and fused.supported(self, L, n_ctx):
            return "tc"
        if self.use_rows_eval and fused.rows_supported(self, L, n_ctx, "fp32") and not _native.is_emulated():
            return "rows_fp32"
        return None

    def _fused_eval_applies(self, profile, targets) -> bool:
        return self._fused_eval_mode(profile, targets) is not None

    force_eval_path: Optional[str] = None
    use_rows_eval = True    # class default; set False on an instance to keep shapes outside the fused kernel per-op

    # Arithmetic of the fused inference path: "fp32" (3xTF32 tensor-core kernel, scores within 1e-4 of the reference)
    # or "bf16" (BASELINE configs[1]: bf16 operands and tables, fp32 accumulation / softmax / LayerNorm, within 1e-2)
    eval_dtype = "fp32"

    def set_eval_dtype(self, dtype) -> "CARCA":
        name = {torch.float32: "fp32", torch.bfloat16: "bf16", "fp32": "fp32", "f32": "fp32", "bf16": "bf16"}.get(dtype)
        if name is None:
            raise ValueError(f"eval_dtype must be fp32 or bf16, got {dtype!r}")
        self.eval_dtype = name
        return self

    use_fused_train = True  # class default; set False on an instance to force the per-op training path

    def _fused_train_applies(self, profile, targets) -> bool:
        """Train mode, d == 64, L <= 256 with at most 64 active positions per user (always true for L <= 64),
        stock blocks / decoder, 1-2 target tuples of L positions each: the
        fused training core (csrc/fused_train.cuh) handles dropout -> blocks -> LayerNorm -> decoder."""
        if not (self.training and self.use_fused_train) or not 1 <= len(targets) <= 2:
            return False
        L = profile[0].shape[1]
        if self.norm.weight.shape[0] != 64 or L > 256 or profile[0].dim() != 2:
            return False
        blocks = list(self.encoder)
        if len(blocks) > 8 or not all(type(b) is SelfAttentionBlock for b in blocks):
            return False
        heads = {b.attn.H for b in blocks}
        resid = {bool(b.residual) for b in blocks}
        drops = {float(b.dropout1.p) for b in blocks} | {float(b.attn.dropout.p) for b in blocks} | {float(self.dropout.p)}
        dec = self.decoder
        if type(dec) is CrossAttentionBlock:
            heads.add(dec.attn.H)
            drops.add(float(dec.attn.dropout.p))
        elif type(dec) is not DotProduct:
            return False
        if len(heads) > 1 or len(resid) > 1 or len(drops) != 1 or (heads and next(iter(heads)) not in (1, 2, 4)):
            return False
        if not all(t[0].dim() == 2 and t[0].shape[1] == L for t in targets):
            return False
        if L > 64:
            # a user's ACTIVE positions must fit one 64-row bin: one device reduction + host read per step
            # (not possible while a CUDA graph is being captured)
            if self._fits_override is not None:        # GraphedTrainStep checked the batch before replaying
                return self._fits_override
            if profile[0].is_cuda and torch.cuda.is_current_stream_capturing():
                return False
            return self.max_active_positions(profile[0], [t[0] for t in targets]) <= 64
        return True

    _fits_override: Optional[bool] = None

    @staticmethod
    def max_active_positions(p_x: Tensor, o_xs: List[Tensor]) -> int:
        """Largest number of positions of one user whose profile id or any target id is non-zero (host read)."""
        active = p_x != 0
        for o_x in o_xs:
            active = active | (o_x != 0)
        return int(active.sum(dim=1).max().item()) if active.numel() else 0

    def _folded_embedding_table(self, profile, targets) -> Optional[ItemAttrTable]:
        """The sparse item->attribute table when AllEmbedding can run inside the fused training kernels
        (device-resident CSR attributes for every row set, built-in positional encodings), else None."""
        emb = self.embeds
        if type(emb) is not AllEmbedding or not hasattr(emb.enc, "table") or profile[2].shape[-1] > 8:
            return None
        tables = [a if isinstance(a, ItemAttrTable) else (emb.attr_table if a is None else False)
                  for a in [profile[1]] + [t[1] for t in targets]]
        first = tables[0]
        if not isinstance(first, ItemAttrTable) or not first.is_sparse or any(t is not first for t in tables):
            return None
        return first

    def _forward_fused_train(self, profile, targets) -> Tensor:
        p_x, p_a, p_c = profile
        with ops.forward_seed():
            blocks = list(self.encoder)
            dec = self.decoder
            is_ca = type(dec) is CrossAttentionBlock
            H = blocks[0].attn.H if blocks else (dec.attn.H if is_ca else 1)
            params = [t for b in blocks for t in b._params()] + [self.norm.weight, self.norm.bias]
            if is_ca:
                a = dec.attn
                params += [a.WQ.weight, a.WQ.bias, a.WK.weight, a.WK.bias, a.WV.weight, a.WV.bias, dec.ffn.weight,
                           dec.ffn.bias]
            cfg = (H, len(blocks), 1 if is_ca else 0, bool(blocks[0].residual) if blocks else True,
                   bool(dec.residual) if is_ca else True, float(self.dropout.p), ops.current_seed())
            o_x1 = targets[1][0] if len(targets) > 1 else None
            table = self._folded_embedding_table(profile, targets)
            if table is not None:
                # AllEmbedding runs inside the kernels too (folded tables): ids + context in, probabilities out
                emb = self.embeds
                pos = emb.enc.table(p_x.shape[1])
                eparams = [emb.items_embed.weight, emb.feats_embed.weight, emb.feats_embed.bias,
                           emb.joint_embed.weight, emb.joint_embed.bias] + ([pos] if pos is not None else [])
                o_c1 = targets[1][2] if len(targets) > 1 else None
                return ops.TrainCoreFn.apply(None, None, None, p_x, targets[0][0], o_x1,
                                             (table, p_c, targets[0][2], o_c1, pos is not None), cfg, *eparams, *params)
            p_mask = get_mask(p_x)
            p_e = self.embeds.forward(p_x, p_a, p_c, p_mask, False)
            o_es = [self.embeds.forward(o_x, o_a, o_c, get_mask(o_x), True) for (o_x, o_a, o_c) in targets]
            o_e1 = o_es[1] if len(o_es) > 1 else None
            return ops.TrainCoreFn.apply(p_e, o_es[0], o_e1, p_x, targets[0][0], o_x1, None, cfg, *params)

    def forward(self, profile: Tuple[Tensor, Tensor, Tensor],
                targets: List[Tuple[Tensor, Tensor, Tensor]]) -> Tensor:
        mode = self._fused_eval_mode(profile, targets)
        if mode is not None:
            from . import fused
            if mode == "tc":
                return fused.forward(self, profile, targets)
            return fused.forward_rows(self, profile, targets, precision=mode[5:])
        if self._fused_train_applies(profile, targets):
            return self._forward_fused_train(profile, targets)
        with ops.forward_seed():
            p_e, p_mask = self.encode(profile)
            y_preds = []
            for t_idx, (o_x, o_a, o_c) in enumerate(targets):
                o_mask = get_mask(o_x)
                o_e = self.embeds.forward(o_x, o_a, o_c, o_mask, True)
                if isinstance(self.decoder, CrossAttentionBlock):
                    self.decoder._site = ops.SITE_DECODER_ATTN + t_idx
                y_preds.append(self.decoder.forward(o_e, o_mask, p_e, p_mask))
            return y_preds[0] if len(y_preds) == 1 else torch.cat(y_preds, dim=-1)


# ------------------------------------------------------------------------------- loss
class BinaryCrossEntropy(nn.Module):
    """Masked mean of -(t log(y+eps) + (1-t) log(1-y+eps)) (src/carca.py:437-444).

    `reduce_sums` (set by the data-parallel wrapper) all-reduces the two partial sums so every rank
    normalises by the global mask count, as a single-process run on the whole batch would.
    """

    def __init__(self):
        super().__init__()
        self.reduce_sums = None

    def forward(self, y_pred: Tensor, y_true: Tensor, mask: Tensor, eps: float = 1e-8) -> Tensor:
        return ops.BCEFn.apply(y_pred, y_true, mask, eps, self.reduce_sums)
