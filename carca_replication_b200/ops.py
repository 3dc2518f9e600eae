"""torch.autograd glue over the C ABI (include/carca_b200.h).

Each Function allocates its outputs / saved activations with torch (device memory only), passes
raw pointers to libcarca_b200.so on the current CUDA stream, and keeps no Python-side math.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _native as N
from .attrs import ItemAttrTable

SITE_EMBED = 0
SITE_DECODER_ATTN = 1000


# ------------------------------------------------------------------------------- dropout seeds
class _SeedState(threading.local):
    def __init__(self):
        self.counter = 0
        self.forced: Optional[int] = None     # tests pin the Philox seed to compare with the oracle
        self.active: Optional[int] = None     # seed of the CARCA.forward currently running
        self.salt = 0                         # data-parallel rank: independent dropout masks per rank


_seeds = _SeedState()


def set_dropout_seed(seed: Optional[int]) -> None:
    """Pin (or with None, release) the Philox seed used by every following forward pass."""
    _seeds.forced = None if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF


def set_seed_salt(salt: int) -> None:
    """Mixed into every fresh seed: UserDataParallel passes its rank, so that ranks (which all start from the same
    torch.initial_seed() and counter) do not draw identical dropout masks for the same row / site indices."""
    _seeds.salt = int(salt)


def fresh_seed() -> int:
    if _seeds.forced is not None:
        return _seeds.forced
    _seeds.counter += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _seeds.counter * 0xD1B54A32D192ED03
            + _seeds.salt * 0xA24BAED4963EE407) & 0xFFFFFFFFFFFFFFFF


class forward_seed:
    """Context: one Philox seed shared by all dropout sites of one model forward."""

    def __enter__(self):
        self.prev = _seeds.active
        _seeds.active = fresh_seed()
        return _seeds.active

    def __exit__(self, *exc):
        _seeds.active = self.prev
        return False


def current_seed() -> int:
    return _seeds.active if _seeds.active is not None else fresh_seed()


# ------------------------------------------------------------------------------- small helpers
def _c(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def as_ids(x: Tensor) -> Tensor:
    """int32 contiguous ids (nn.Embedding accepts int32 and int64; the kernels take int32)."""
    return _c(x if x.dtype == torch.int32 else x.to(torch.int32))


def as_f32(t: Tensor) -> Tensor:
    return _c(t if t.dtype == torch.float32 else t.to(torch.float32))


def padding_mask(ids: Tensor) -> Tensor:
    """get_mask (src/utils.py:6-7) on device: integer ids, or any floating input (compared as float32)."""
    N.require_device(ids)
    if ids.dtype.is_floating_point:
        x = as_f32(ids)
        m = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        N.call("carca_padding_mask_f32", N.f32p(m), N.f32p(x), x.numel(), N.stream())
        return m
    x = as_ids(ids)
    m = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    N.call("carca_padding_mask", N.f32p(m), N.i32p(x), x.numel(), N.stream())
    return m


def _struct(cls, names: Sequence[str], tensors: Sequence[Optional[Tensor]]):
    s = cls()
    for n, t in zip(names, tensors):
        setattr(s, n, None if t is None else N.f32p(t))
    return s


# ------------------------------------------------------------------------------- embedding
def _embed_params(E, Wf, WfT, bf, Wj, bj, pos, A, Cn) -> N.EmbedParams:
    p = N.EmbedParams()
    p.n_items, p.d = E.shape
    p.g = Wf.shape[0]
    p.n_attrs, p.n_ctx = A, Cn
    p.items_embed, p.feats_w, p.feats_b = N.f32p(E), N.f32p(Wf), N.f32p(bf)
    p.feats_wT = None if WfT is None else N.f32p(WfT)
    p.joint_w, p.joint_b = N.f32p(Wj), N.f32p(bj)
    p.pos = None if pos is None else N.f32p(pos)
    p.pos_len = 0 if pos is None else pos.shape[0]
    return p


def _attr_source(table: Optional[ItemAttrTable], dense: Optional[Tensor]) -> N.AttrSource:
    s = N.AttrSource()
    if dense is not None:
        s.kind, s.dense = 2, N.f32p(dense)
    elif table.is_sparse:
        s.kind = 0
        s.csr_rowptr, s.csr_cols, s.csr_vals = N.i32p(table.rowptr), N.i32p(table.cols), N.f32p(table.vals)
    else:
        s.kind, s.dense = 1, N.f32p(table.dense)
    return s


DENSE_SCAN_MIN_ATTRS = 1024


class EmbedFn(torch.autograd.Function):
    """AllEmbedding.forward (src/carca.py:85-95) -> carca_embed_fwd / carca_embed_bwd."""

    @staticmethod
    def forward(ctx, x, c, mask, a_dense, E, Wf, bf, Wj, bj, pos, table, is_target):
        N.require_device(x, c, mask, a_dense, E)
        n_rows, n_cols = x.shape
        P = n_rows * n_cols
        d, g = E.shape[1], Wf.shape[0]
        Cn = c.shape[-1]
        A = Wf.shape[1] - Cn
        x, c, mask = as_ids(x), as_f32(c), as_f32(mask)
        a_dense = None if a_dense is None else as_f32(a_dense)
        if a_dense is None and table is None:
            raise RuntimeError("AllEmbedding: attributes missing (pass the dense tensor or register an ItemAttrTable)")
        if a_dense is not None and a_dense.shape[-1] != A:
            raise RuntimeError(f"AllEmbedding: a has {a_dense.shape[-1]} attributes, weights expect {A}")
        sparse = a_dense is None and table.is_sparse
        WfT = None
        # (dense rows of a large vocabulary — the reference API's multi-hot tensors — are projected by a scan + gather-sum
        # kernel that reads the transposed weights, as the CSR path does: carca_embed_fwd picks it when feats_wT is given)
        if sparse or (a_dense is not None and A >= DENSE_SCAN_MIN_ATTRS and g <= 256):
            WfT = torch.empty((A + Cn, g), dtype=torch.float32, device=E.device)
            N.call("carca_transpose", N.f32p(WfT), N.f32p(_c(Wf)), g, A + Cn, 0, N.stream())
        e = torch.empty((n_rows, n_cols, d), dtype=torch.float32, device=E.device)
        q = torch.empty((P, g), dtype=torch.float32, device=E.device)
        Ec, Wfc, bfc, Wjc, bjc = _c(E), _c(Wf), _c(bf), _c(Wj), _c(bj)
        posc = None if pos is None else _c(pos)
        prm = _embed_params(Ec, Wfc, WfT, bfc, Wjc, bjc, posc, A, Cn)
        src = _attr_source(table, a_dense)
        N.call("carca_embed_fwd", N.f32p(e), N.f32p(q), C.byref(prm), C.byref(src), N.i32p(x), N.f32p(c),
               N.f32p(mask), n_rows, n_cols, int(bool(is_target)), N.stream())
        ctx.save_for_backward(x, c, mask, a_dense, Ec, Wfc, bfc, Wjc, bjc, posc, q)
        ctx.table, ctx.is_target, ctx.sparse = table, bool(is_target), sparse
        return e

    @staticmethod
    def backward(ctx, de):
        x, c, mask, a_dense, E, Wf, bf, Wj, bj, pos, q = ctx.saved_tensors
        n_rows, n_cols = x.shape
        P = n_rows * n_cols
        d, g = E.shape[1], Wf.shape[0]
        Cn = c.shape[-1]
        A = Wf.shape[1] - Cn
        dev = E.device
        de = as_f32(de)
        gE, gWf, gbf, gWj, gbj = (torch.zeros_like(t) for t in (E, Wf, bf, Wj, bj))
        gpos = torch.zeros_like(pos) if (pos is not None and not ctx.is_target) else None
        grads = _struct(N.EmbedGrads, ("items_embed", "feats_w", "feats_b", "joint_w", "joint_b", "pos"),
                        (gE, gWf, gbf, gWj, gbj, gpos))
        WfT = None
        prm = _embed_params(E, Wf, WfT, bf, Wj, bj, pos, A, Cn)
        src = _attr_source(ctx.table, a_dense)
        s_pd = torch.empty((2 * P, d), dtype=torch.float32, device=dev)
        s_pg = torch.empty((P, g), dtype=torch.float32, device=dev)
        s_wT = torch.zeros((A, g), dtype=torch.float32, device=dev) if ctx.sparse else None
        N.call("carca_embed_bwd", C.byref(grads), N.f32p(de), N.f32p(q), C.byref(prm), C.byref(src), N.i32p(x),
               N.f32p(c), N.f32p(mask), n_rows, n_cols, int(ctx.is_target), N.f32p(s_pd), N.f32p(s_pg),
               N.f32p(s_wT), N.stream())
        return None, None, None, None, gE, gWf, gbf, gWj, gbj, gpos, None, None


# ------------------------------------------------------------------------------- dropout / LN / linear
class DropoutFn(torch.autograd.Function):
    """nn.Dropout with the Philox stream (src/carca.py:416)."""

    @staticmethod
    def forward(ctx, x, p, seed, site):
        N.require_device(x)
        x = as_f32(x)
        y = torch.empty_like(x)
        N.call("carca_dropout", N.f32p(y), N.f32p(x), x.numel(), float(p), int(seed), int(site), N.stream())
        ctx.cfg = (float(p), int(seed), int(site))
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed, site = ctx.cfg
        dy = as_f32(dy)
        dx = torch.empty_like(dy)
        N.call("carca_dropout", N.f32p(dx), N.f32p(dy), dy.numel(), p, seed, site, N.stream())
        return dx, None, None, None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(d) (src/carca.py:408,421)."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        N.require_device(x, gamma)
        x = as_f32(x)
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        gamma, beta = _c(gamma), _c(beta)
        N.call("carca_layernorm_fwd", N.f32p(y), N.f32p(mean), N.f32p(rstd), N.f32p(x), N.f32p(gamma),
               N.f32p(beta), rows, d, N.stream())
        ctx.save_for_backward(x, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        d = x.shape[-1]
        rows = x.numel() // d
        dy = as_f32(dy)
        dx = torch.empty_like(x)
        dg, db = torch.zeros_like(gamma), torch.zeros_like(gamma)
        N.call("carca_layernorm_bwd", N.f32p(dx), N.f32p(dg), N.f32p(db), N.f32p(dy), N.f32p(x), N.f32p(mean),
               N.f32p(rstd), N.f32p(gamma), rows, d, 0, N.stream())
        return dx, dg, db


class LinearFn(torch.autograd.Function):
    """nn.Linear (src/carca.py:238-240)."""

    @staticmethod
    def forward(ctx, x, w, b):
        N.require_device(x, w)
        x, w, b = as_f32(x), _c(w), _c(b)
        K = x.shape[-1]
        M = x.numel() // K
        Nn = w.shape[0]
        y = torch.empty((*x.shape[:-1], Nn), dtype=torch.float32, device=x.device)
        N.call("carca_linear_fwd", N.f32p(y), N.f32p(x), N.f32p(w), N.f32p(b), M, Nn, K, 0, N.stream())
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        K = x.shape[-1]
        M = x.numel() // K
        Nn = w.shape[0]
        dy = as_f32(dy)
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        db = torch.zeros(Nn, dtype=torch.float32, device=x.device)
        N.call("carca_linear_bwd_input", N.f32p(dx), N.f32p(dy), N.f32p(w), M, Nn, K, 0, N.stream())
        N.call("carca_linear_bwd_weight", N.f32p(dw), N.f32p(db), N.f32p(dy), N.f32p(x), M, Nn, K, N.stream())
        return dx, dw, db


class AttentionCoreFn(torch.autograd.Function):
    """src/carca.py:242-260 after the projections -> carca_attention_fwd / _bwd."""

    @staticmethod
    def forward(ctx, Q, K, V, q_mask, k_mask, H, causal, p, seed, site, want_w):
        N.require_device(Q, K, V, q_mask, k_mask)
        Q, K, V, q_mask, k_mask = (as_f32(t) for t in (Q, K, V, q_mask, k_mask))
        B, Lq, d = Q.shape
        Lk = K.shape[1]
        O = torch.empty_like(Q)
        W = torch.empty((B, H, Lq, Lk), dtype=torch.float32, device=Q.device) if want_w else None
        cz = (0, 0) if causal is None else (1, int(causal))
        N.call("carca_attention_fwd", N.f32p(O), N.f32p(W), N.f32p(Q), N.f32p(K), N.f32p(V), N.f32p(q_mask),
               N.f32p(k_mask), B, H, Lq, Lk, d, cz[0], cz[1], float(p), int(seed), int(site), N.stream())
        ctx.save_for_backward(Q, K, V, q_mask, k_mask)
        ctx.cfg = (H, cz, float(p), int(seed), int(site))
        if want_w:
            ctx.mark_non_differentiable(W)
            return O, W
        return O, None

    @staticmethod
    def backward(ctx, dO, _dW):
        Q, K, V, q_mask, k_mask = ctx.saved_tensors
        H, cz, p, seed, site = ctx.cfg
        B, Lq, d = Q.shape
        Lk = K.shape[1]
        dO = as_f32(dO)
        dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
        N.call("carca_attention_bwd", N.f32p(dQ), N.f32p(dK), N.f32p(dV), N.f32p(dO), N.f32p(Q), N.f32p(K),
               N.f32p(V), N.f32p(q_mask), N.f32p(k_mask), B, H, Lq, Lk, d, cz[0], cz[1], p, seed, site, N.stream())
        return dQ, dK, dV, None, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------- encoder block
class SABlockFn(torch.autograd.Function):
    """SelfAttentionBlock.forward (src/carca.py:297-318) -> carca_sa_block_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, mask, H, residual, p, seed, block_index, *params):
        N.require_device(x, mask, *params)
        x, mask = as_f32(x), as_f32(mask)
        params = tuple(_c(t) for t in params)
        B, L, d = x.shape
        dev = x.device
        P = B * L

        def big():
            return torch.empty((P, d), dtype=torch.float32, device=dev)

        def small():
            return torch.empty(P, dtype=torch.float32, device=dev)

        saved = dict(qn=big(), mean1=small(), rstd1=small(), Q=big(), K=big(), V=big(), s=big(), mean2=small(),
                     rstd2=small(), s2=big(), a1=big())
        sv = _struct(N.BlockSaved, N.BLOCK_SAVED_NAMES, [saved[n] for n in N.BLOCK_SAVED_NAMES])
        w = _struct(N.BlockParams, N.BLOCK_PARAM_NAMES, params)
        out = torch.empty_like(x)
        N.call("carca_sa_block_fwd", N.f32p(out), C.byref(sv), N.f32p(x), N.f32p(mask), C.byref(w), B, L, d, int(H),
               int(bool(residual)), float(p), int(seed), int(block_index), N.stream())
        ctx.save_for_backward(x, mask, *params, *[saved[n] for n in N.BLOCK_SAVED_NAMES])
        ctx.cfg = (int(H), int(bool(residual)), float(p), int(seed), int(block_index))
        return out

    @staticmethod
    def backward(ctx, dout):
        H, residual, p, seed, block_index = ctx.cfg
        st = ctx.saved_tensors
        x, mask = st[0], st[1]
        params = st[2:2 + len(N.BLOCK_PARAM_NAMES)]
        saved = st[2 + len(N.BLOCK_PARAM_NAMES):]
        B, L, d = x.shape
        dout = as_f32(dout)
        grads = tuple(torch.zeros_like(t) for t in params)
        gs = _struct(N.BlockParams, N.BLOCK_PARAM_NAMES, grads)
        w = _struct(N.BlockParams, N.BLOCK_PARAM_NAMES, params)
        sv = _struct(N.BlockSaved, N.BLOCK_SAVED_NAMES, saved)
        dx = torch.empty_like(x)
        scratch = torch.empty((4, B * L, d), dtype=torch.float32, device=x.device)
        N.call("carca_sa_block_bwd", N.f32p(dx), C.byref(gs), N.f32p(dout), N.f32p(x), N.f32p(mask), C.byref(w),
               C.byref(sv), B, L, d, H, residual, p, seed, block_index, N.f32p(scratch), N.stream())
        return (dx, None, None, None, None, None, None, *grads)


# ------------------------------------------------------------------------------- decoders
class DotScoreFn(torch.autograd.Function):
    """DotProduct.forward (src/carca.py:358-365)."""

    @staticmethod
    def forward(ctx, o, p, per_position):
        N.require_device(o, p)
        o, p = as_f32(o), as_f32(p)
        B, T, d = o.shape
        Lp = p.shape[1]
        y = torch.empty((B, T), dtype=torch.float32, device=o.device)
        N.call("carca_dot_score_fwd", N.f32p(y), N.f32p(p), N.f32p(o), B, T, Lp, d, int(bool(per_position)), T, 0,
               N.stream())
        ctx.save_for_backward(o, p, y)
        ctx.per_position = int(bool(per_position))
        return y

    @staticmethod
    def backward(ctx, dy):
        o, p, y = ctx.saved_tensors
        B, T, d = o.shape
        Lp = p.shape[1]
        dy = as_f32(dy)
        d_o = torch.empty_like(o)
        d_p = torch.zeros_like(p)
        N.call("carca_dot_score_bwd", N.f32p(d_o), N.f32p(d_p), N.f32p(dy), N.f32p(y), N.f32p(p), N.f32p(o), B, T,
               Lp, d, ctx.per_position, T, 0, N.stream())
        return d_o, d_p, None


class CrossScoreFn(torch.autograd.Function):
    """CrossAttentionBlock.forward (src/carca.py:338-349) -> carca_cross_score_fwd / _bwd."""

    @staticmethod
    def forward(ctx, o, o_mask, p, p_mask, H, residual, training, p_drop, seed, site, *params):
        N.require_device(o, o_mask, p, p_mask, *params)
        o, o_mask, p, p_mask = (as_f32(t) for t in (o, o_mask, p, p_mask))
        params = tuple(_c(t) for t in params)
        B, T, d = o.shape
        Lp = p.shape[1]
        dev = o.device
        saved = dict(Q=torch.empty((B * T, d), dtype=torch.float32, device=dev),
                     K=torch.empty((B * Lp, d), dtype=torch.float32, device=dev),
                     V=torch.empty((B * Lp, d), dtype=torch.float32, device=dev),
                     s=torch.empty((B * T, d), dtype=torch.float32, device=dev))
        sv = _struct(N.CrossSaved, N.CROSS_SAVED_NAMES, [saved[n] for n in N.CROSS_SAVED_NAMES])
        w = _struct(N.CrossParams, N.CROSS_PARAM_NAMES, params)
        y = torch.empty((B, T), dtype=torch.float32, device=dev)
        N.call("carca_cross_score_fwd", N.f32p(y), C.byref(sv), N.f32p(o), N.f32p(o_mask), N.f32p(p), N.f32p(p_mask),
               C.byref(w), B, T, Lp, d, int(H), int(bool(residual)), int(bool(training)), float(p_drop), int(seed),
               int(site), T, 0, N.stream())
        ctx.save_for_backward(o, o_mask, p, p_mask, y, *params, *[saved[n] for n in N.CROSS_SAVED_NAMES])
        ctx.cfg = (int(H), int(bool(residual)), int(bool(training)), float(p_drop), int(seed), int(site))
        return y

    @staticmethod
    def backward(ctx, dy):
        H, residual, training, p_drop, seed, site = ctx.cfg
        st = ctx.saved_tensors
        o, o_mask, p, p_mask, y = st[:5]
        params = st[5:5 + len(N.CROSS_PARAM_NAMES)]
        saved = st[5 + len(N.CROSS_PARAM_NAMES):]
        B, T, d = o.shape
        Lp = p.shape[1]
        dy = as_f32(dy)
        grads = tuple(torch.zeros_like(t) for t in params)
        gs = _struct(N.CrossParams, N.CROSS_PARAM_NAMES, grads)
        w = _struct(N.CrossParams, N.CROSS_PARAM_NAMES, params)
        sv = _struct(N.CrossSaved, N.CROSS_SAVED_NAMES, saved)
        d_o = torch.empty_like(o)
        d_p = torch.zeros_like(p)
        scratch = torch.empty((4, max(B * T, B * Lp), d), dtype=torch.float32, device=o.device)
        N.call("carca_cross_score_bwd", N.f32p(d_o), N.f32p(d_p), C.byref(gs), N.f32p(dy), N.f32p(y), C.byref(sv),
               N.f32p(o), N.f32p(o_mask), N.f32p(p), N.f32p(p_mask), C.byref(w), B, T, Lp, d, H, residual, training,
               p_drop, seed, site, T, 0, N.f32p(scratch), N.stream())
        return (d_o, None, d_p, None, None, None, None, None, None, None, *grads)


# ------------------------------------------------------------------------------- fused training core
# Optional provider of the flat gradient buffer: (n_floats, device, parameter tensors) -> float32 tensor or None.  The data-parallel wrapper
# installs one that returns peer-mapped communication memory (parallel.PeerAllReduce.buffer).
FLAT_GRAD_ALLOC = None


class _FlatZeros:
    """Carves zero-initialised gradient tensors out of ONE zero-filled buffer (one fill launch instead of one
    per parameter)."""

    def __init__(self, shapes, device, params=()):
        sizes = [(int(torch.Size(sh).numel()) + 3) // 4 * 4 for sh in shapes]      # 16-byte aligned pieces
        flat = FLAT_GRAD_ALLOC(sum(sizes), device, params) if FLAT_GRAD_ALLOC is not None else None
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device) if flat is None else flat.zero_()
        self.views, off = [], 0
        for sh, n in zip(shapes, sizes):
            self.views.append(self.flat[off:off + int(torch.Size(sh).numel())].view(sh))
            off += n


class TrainCoreFn(torch.autograd.Function):
    """src/carca.py:415-431 in train mode (dropout, encoder blocks, final LayerNorm, decoder per target tuple, and
    optionally AllEmbedding itself through folded tables) -> carca_train_core_fwd / _bwd: one forward and one
    backward kernel over the ACTIVE positions only.

    emb is None: p_e / o_e* are the embedded rows (any Embedding module) and params = block params, norm, decoder.
    emb = (table, p_c, o_c0, o_c1, has_pos): p_e / o_e* are None, the kernels embed the rows themselves and params
    starts with AllEmbedding's E, Wf, bf, Wj, bj (, pos)."""

    @staticmethod
    def forward(ctx, p_e, o_e0, o_e1, p_x, o_x0, o_x1, emb, cfg, *params):
        N.require_device(p_e, o_e0, o_e1, p_x, o_x0, o_x1, *params)
        H, n_blocks, decoder_kind, residual_sa, residual_ca, p_drop, seed = cfg
        n_tuples = 1 if o_x1 is None else 2
        p_x, o_x0 = as_ids(p_x), as_ids(o_x0)
        o_x1 = None if o_x1 is None else as_ids(o_x1)
        params = tuple(_c(t) for t in params)
        B, L = p_x.shape
        dev = params[0].device
        L_ = N.lib()
        c = N.TrainCore()
        keep = []
        n_emb = 0
        if emb is not None:
            table, p_c, o_c0, o_c1, has_pos = emb
            N.require_device(p_c, o_c0, o_c1)
            n_emb = 6 if has_pos else 5
            E, Wf, bf, Wj, bj = params[:5]
            pos = params[5] if has_pos else None
            Cn = p_c.shape[-1]
            prm = _embed_params(E, Wf, None, bf, Wj, bj, pos, Wf.shape[1] - Cn, Cn)
            src = _attr_source(table, None)
            p_c, o_c0 = as_f32(p_c), as_f32(o_c0)
            o_c1 = None if o_c1 is None else as_f32(o_c1)
            fold = torch.empty(L_.carca_train_core_fold_floats(C.byref(prm)), dtype=torch.float32, device=dev)
            c.embed, c.attrs = C.pointer(prm), C.pointer(src)
            c.p_c, c.fold = N.f32p(p_c), fold.data_ptr()
            c.o_c[0] = N.f32p(o_c0)
            if n_tuples == 2:
                c.o_c[1] = N.f32p(o_c1)
            keep += [prm, src, table]
            tensors = (p_c, o_c0, o_c1, fold)
        else:
            p_e, o_e0 = as_f32(p_e), as_f32(o_e0)
            o_e1 = None if o_e1 is None else as_f32(o_e1)
            c.p_e = N.f32p(p_e)
            c.o_e[0] = N.f32p(o_e0)
            if n_tuples == 2:
                c.o_e[1] = N.f32p(o_e1)
            tensors = (p_e, o_e0, o_e1, None)
        core = params[n_emb:]
        nbp = len(N.BLOCK_PARAM_NAMES)
        blocks = (N.BlockParams * max(n_blocks, 1))()
        for b in range(n_blocks):
            blocks[b] = _struct(N.BlockParams, N.BLOCK_PARAM_NAMES, core[b * nbp:(b + 1) * nbp])
        rest = core[n_blocks * nbp:]
        c.B, c.L, c.n_heads, c.n_blocks, c.n_tuples = B, L, int(H), int(n_blocks), n_tuples
        c.decoder_kind, c.residual_sa, c.residual_ca = int(decoder_kind), int(bool(residual_sa)), int(bool(residual_ca))
        c.p_drop, c.seed = float(p_drop), int(seed)
        c.p_x = N.i32p(p_x)
        c.o_x[0] = N.i32p(o_x0)
        if n_tuples == 2:
            c.o_x[1] = N.i32p(o_x1)
        c.blocks = blocks
        c.norm_g, c.norm_b = N.f32p(rest[0]), N.f32p(rest[1])
        if decoder_kind == 1:
            c.cross = _struct(N.CrossParams, N.CROSS_PARAM_NAMES, rest[2:2 + len(N.CROSS_PARAM_NAMES)])
        rows = torch.empty(L_.carca_train_core_rows_ints(B), dtype=torch.int32, device=dev)
        saved = torch.empty(L_.carca_train_core_saved_floats(B, int(n_blocks), n_tuples), dtype=torch.float32,
                            device=dev)
        c.rows, c.saved = rows.data_ptr(), saved.data_ptr()
        y = torch.empty((B, n_tuples * L), dtype=torch.float32, device=dev)
        N.call("carca_train_core_fwd", N.f32p(y), n_tuples * L, C.byref(c), N.stream())
        ctx.save_for_backward(p_x, o_x0, o_x1, rows, saved, *tensors, *params)
        ctx.core, ctx.keep = c, (blocks, keep)
        ctx.cfg = (n_tuples, int(n_blocks), int(decoder_kind), n_emb, emb is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        st = ctx.saved_tensors
        t0, t1, t2 = st[5], st[6], st[7]          # p_e, o_e0, o_e1  or  p_c, o_c0, o_c1
        params = st[9:]
        c = ctx.core
        n_tuples, n_blocks, decoder_kind, n_emb, embed_mode = ctx.cfg
        nbp = len(N.BLOCK_PARAM_NAMES)
        dy = as_f32(dy)
        dev = dy.device
        # every gradient comes out of ONE zero-filled buffer (one fill launch; atomics accumulate into it), the
        # parameter gradients first and contiguous so data parallelism can all-reduce the buffer in place
        core = params[n_emb:]
        shapes = [t.shape for t in params[:n_emb]] + [t.shape for t in core]
        if embed_mode:
            shapes.append((int(N.lib().carca_train_core_fold_floats(c.embed)),))
        else:
            shapes += [t0.shape, t1.shape] + ([t2.shape] if t2 is not None else [])
        z = _FlatZeros(shapes, dev, params)
        egrads = z.views[:n_emb]
        grads = z.views[n_emb:n_emb + len(core)]
        extra = z.views[n_emb + len(core):]
        gblocks = (N.BlockParams * max(n_blocks, 1))()
        for b in range(n_blocks):
            gblocks[b] = _struct(N.BlockParams, N.BLOCK_PARAM_NAMES, grads[b * nbp:(b + 1) * nbp])
        rest = grads[n_blocks * nbp:]
        gcross = None
        if decoder_kind == 1:
            gcross = C.byref(_struct(N.CrossParams, N.CROSS_PARAM_NAMES, rest[2:2 + len(N.CROSS_PARAM_NAMES)]))
        if embed_mode:
            gE, gWf, gbf, gWj, gbj = egrads[:5]
            gpos = egrads[5] if n_emb == 6 else None
            d_fold = extra[0]
            gemb = _struct(N.EmbedGrads, ("items_embed", "feats_w", "feats_b", "joint_w", "joint_b", "pos"),
                           (gE, gWf, gbf, gWj, gbj, gpos))
            N.call("carca_train_core_bwd", None, None, None, gblocks, N.f32p(rest[0]), N.f32p(rest[1]), gcross,
                   C.byref(gemb), N.f32p(d_fold), N.f32p(dy), dy.shape[1], C.byref(c), N.stream())
            return (None,) * 8 + tuple(egrads) + tuple(grads)
        d_pe, d_o0 = extra[0], extra[1]
        d_o1 = extra[2] if t2 is not None else None
        N.call("carca_train_core_bwd", N.f32p(d_pe), N.f32p(d_o0), N.f32p(d_o1), gblocks, N.f32p(rest[0]),
               N.f32p(rest[1]), gcross, None, None, N.f32p(dy), dy.shape[1], C.byref(c), N.stream())
        return (d_pe, d_o0, d_o1, None, None, None, None, None, *grads)


# ------------------------------------------------------------------------------- loss / metrics
class BCEFn(torch.autograd.Function):
    """BinaryCrossEntropy.forward (src/carca.py:441-444); `reduce_sums` lets data-parallel runs
    all-reduce (sum ell*mask, sum mask) so the loss is the global masked mean."""

    @staticmethod
    def forward(ctx, y_pred, y_true, mask, eps, reduce_sums):
        N.require_device(y_pred, y_true, mask)
        y_pred, mask = as_f32(y_pred), as_f32(mask)
        y_true = as_ids(y_true)
        n = y_pred.numel()
        sums = torch.zeros(2, dtype=torch.float32, device=y_pred.device)
        loss = torch.empty((), dtype=torch.float32, device=y_pred.device)
        N.call("carca_bce_sums", N.f32p(sums), N.f32p(y_pred), N.i32p(y_true), N.f32p(mask), n, float(eps),
               N.stream())
        if reduce_sums is not None:
            reduce_sums(sums)
        N.call("carca_bce_finalize", N.f32p(loss), N.f32p(sums), N.stream())
        ctx.save_for_backward(y_pred, y_true, mask, sums)
        ctx.eps = float(eps)
        return loss

    @staticmethod
    def backward(ctx, g):
        y_pred, y_true, mask, sums = ctx.saved_tensors
        g = as_f32(g).reshape(1)
        dy = torch.empty_like(y_pred)
        N.call("carca_bce_bwd", N.f32p(dy), N.f32p(g), N.f32p(sums), N.f32p(y_pred), N.i32p(y_true), N.f32p(mask),
               y_pred.numel(), ctx.eps, N.stream())
        return dy, None, None, None, None


def rank_metrics_(acc: Tensor, y_pred: Tensor, y_true: Tensor, k: int, first_rank: Optional[Tensor] = None) -> None:
    """acc (fp64[3], device) += [hits@k, sum 1/log2(rank+2), rows]; src/train.py:15-32."""
    N.require_device(acc, y_pred, y_true)
    y_pred = y_pred if y_pred.dtype == torch.float32 else y_pred.float()
    y_true = y_true if y_true.dtype == torch.int32 else y_true.to(torch.int32)
    if y_pred.stride(-1) != 1:
        y_pred = y_pred.contiguous()
    if y_true.stride(-1) != 1:
        y_true = y_true.contiguous()
    B, T = y_pred.shape
    N.call("carca_rank_metrics", acc.data_ptr(), None if first_rank is None else N.i32p(first_rank),
           y_pred.data_ptr(), y_true.data_ptr(), B, T, y_pred.stride(0), y_true.stride(0), int(k), N.stream())


_eval_work: dict = {}


def eval_metrics_(stats: Tensor, y_pred: Tensor, y_true: Tensor, o_x: Tensor, k: int, eps: float = 1e-8) -> None:
    """stats (fp64[4], device) += [hits@k, sum 1/log2(rank+2), rows, masked BCE of the batch]: the per-batch
    reductions of evaluate() (src/train.py:44-50) in one launch; mask = get_mask(o_x) (src/utils.py:6-7)."""
    N.require_device(stats, y_pred, y_true, o_x)
    if stats.dtype != torch.float64 or stats.numel() != 4 or not stats.is_contiguous():
        raise ValueError("stats must be a contiguous float64[4] device tensor")
    y_pred = y_pred if y_pred.dtype == torch.float32 else y_pred.float()
    y_true = y_true if y_true.dtype == torch.int32 else y_true.to(torch.int32)
    o_x = as_ids(o_x)
    if y_pred.stride(-1) != 1:
        y_pred = y_pred.contiguous()
    if y_true.stride(-1) != 1:
        y_true = y_true.contiguous()
    if o_x.stride(-1) != 1:
        o_x = o_x.contiguous()
    B, T = y_pred.shape
    key = str(stats.device)
    work = _eval_work.get(key)
    if work is None:      # scratch of the kernel (it leaves it zero again); one per device, calls on a stream are ordered
        work = _eval_work[key] = torch.zeros(3, dtype=torch.float64, device=stats.device)
    N.call("carca_eval_metrics", stats.data_ptr(), work.data_ptr(), y_pred.data_ptr(), y_true.data_ptr(), o_x.data_ptr(),
           B, T, y_pred.stride(0), y_true.stride(0), o_x.stride(0), int(k), float(eps), N.stream())


# ------------------------------------------------------------------------------- module variants
class FeatsFn(torch.autograd.Function):
    """q = Wf [a | c] + bf (src/carca.py:113, :138): dense `a` tensor or the device-resident ItemAttrTable."""

    @staticmethod
    def forward(ctx, x, c, a_dense, Wf, bf, table):
        N.require_device(x, c, a_dense, Wf)
        P = x.numel()
        g = Wf.shape[0]
        Cn = 0 if c is None else c.shape[-1]
        A = Wf.shape[1] - Cn
        x = as_ids(x)
        c = None if c is None else as_f32(c)
        a_dense = None if a_dense is None else as_f32(a_dense)
        if a_dense is None and table is None:
            raise RuntimeError("attribute embedding: attributes missing (pass the dense tensor or register an ItemAttrTable)")
        sparse = a_dense is None and table.is_sparse
        Wfc, bfc = _c(Wf), _c(bf)
        WfT = None
        if sparse:
            WfT = torch.empty((A + Cn, g), dtype=torch.float32, device=Wf.device)
            N.call("carca_transpose", N.f32p(WfT), N.f32p(Wfc), g, A + Cn, 0, N.stream())
        q = torch.empty((*x.shape, g), dtype=torch.float32, device=Wf.device)
        prm = N.EmbedParams()
        prm.g, prm.n_attrs, prm.n_ctx = g, A, Cn
        prm.feats_w, prm.feats_b = N.f32p(Wfc), N.f32p(bfc)
        prm.feats_wT = None if WfT is None else N.f32p(WfT)
        src = _attr_source(table, a_dense)
        N.call("carca_feats_fwd", N.f32p(q), C.byref(prm), C.byref(src), N.i32p(x), N.f32p(c), P, N.stream())
        ctx.save_for_backward(x, c, a_dense, Wfc, bfc)
        ctx.table, ctx.sparse = table, sparse
        return q

    @staticmethod
    def backward(ctx, dq):
        x, c, a_dense, Wf, bf = ctx.saved_tensors
        P = x.numel()
        g = Wf.shape[0]
        Cn = 0 if c is None else c.shape[-1]
        A = Wf.shape[1] - Cn
        dq = as_f32(dq)
        gWf, gbf = torch.zeros_like(Wf), torch.zeros_like(bf)
        prm = N.EmbedParams()
        prm.g, prm.n_attrs, prm.n_ctx = g, A, Cn
        prm.feats_w, prm.feats_b = N.f32p(Wf), N.f32p(bf)
        src = _attr_source(ctx.table, a_dense)
        s_wT = torch.zeros((A, g), dtype=torch.float32, device=Wf.device) if ctx.sparse else None
        N.call("carca_feats_bwd", N.f32p(gWf), N.f32p(gbf), N.f32p(dq), C.byref(prm), C.byref(src), N.i32p(x),
               N.f32p(c), P, N.f32p(s_wT), N.stream())
        return None, None, None, gWf, gbf, None


class GatherRowsFn(torch.autograd.Function):
    """alpha * nn.Embedding(x) with padding_idx = 0 (src/carca.py:161-162, :187-188)."""

    @staticmethod
    def forward(ctx, x, E, alpha):
        N.require_device(x, E)
        x, Ec = as_ids(x), _c(E)
        d = E.shape[1]
        out = torch.empty((*x.shape, d), dtype=torch.float32, device=E.device)
        N.call("carca_gather_rows_fwd", N.f32p(out), N.f32p(Ec), N.i32p(x), float(alpha), x.numel(), d, N.stream())
        ctx.save_for_backward(x)
        ctx.shape, ctx.alpha = tuple(E.shape), float(alpha)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        dout = as_f32(dout)
        gE = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        N.call("carca_gather_rows_bwd", N.f32p(gE), N.f32p(dout), N.i32p(x), ctx.alpha, x.numel(), ctx.shape[1],
               N.stream())
        return None, gE, None


class PosMaskFn(torch.autograd.Function):
    """(e + pos[position]) * mask: positional encoding of profile rows + the final mask of every embedding."""

    @staticmethod
    def forward(ctx, e, pos, mask):
        N.require_device(e, mask)
        e, mask = as_f32(e), as_f32(mask)
        n_rows, n_cols, d = e.shape
        posc = None if pos is None else _c(pos)
        if posc is not None and posc.shape[0] < n_cols:
            raise RuntimeError(f"sequence length {n_cols} > positional table {posc.shape[0]}")
        out = torch.empty_like(e)
        N.call("carca_pos_mask_fwd", N.f32p(out), N.f32p(e), N.f32p(posc), N.f32p(mask), n_rows, n_cols, d, N.stream())
        ctx.save_for_backward(mask)
        ctx.pos_shape = None if pos is None else tuple(pos.shape)
        ctx.dims = (n_rows, n_cols, d)
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        n_rows, n_cols, d = ctx.dims
        dout = as_f32(dout)
        de = torch.empty_like(dout)
        gpos = None if ctx.pos_shape is None else torch.zeros(ctx.pos_shape, dtype=torch.float32, device=dout.device)
        N.call("carca_pos_mask_bwd", N.f32p(de), N.f32p(gpos), N.f32p(dout), N.f32p(mask), n_rows, n_cols, d, N.stream())
        return de, gpos, None


class WDotScoreFn(torch.autograd.Function):
    """WeightedDotProduct.forward (src/carca.py:377-395)."""

    @staticmethod
    def forward(ctx, o, p, per_position, gamma, normalize):
        N.require_device(o, p)
        o, p = as_f32(o), as_f32(p)
        B, T, d = o.shape
        Lp = p.shape[1]
        y = torch.empty((B, T), dtype=torch.float32, device=o.device)
        N.call("carca_wdot_score_fwd", N.f32p(y), N.f32p(p), N.f32p(o), B, T, Lp, d, int(bool(per_position)),
               float(gamma), int(bool(normalize)), T, 0, N.stream())
        ctx.save_for_backward(o, p, y)
        ctx.cfg = (int(bool(per_position)), float(gamma), int(bool(normalize)))
        return y

    @staticmethod
    def backward(ctx, dy):
        o, p, y = ctx.saved_tensors
        B, T, d = o.shape
        Lp = p.shape[1]
        per_position, gamma, normalize = ctx.cfg
        dy = as_f32(dy)
        d_o = torch.empty_like(o)
        d_p = torch.zeros_like(p)
        N.call("carca_wdot_score_bwd", N.f32p(d_o), N.f32p(d_p), N.f32p(dy), N.f32p(y), N.f32p(p), N.f32p(o), B, T, Lp,
               d, per_position, gamma, normalize, T, 0, N.stream())
        return d_o, d_p, None, None, None


def knn_scores(p_a: Tensor, o_a: Tensor) -> Tensor:
    """KNN.forward (src/knn.py:14-21) for one target tuple: <attributes of the last profile item, candidate attributes>."""
    N.require_device(p_a, o_a)
    p_a, o_a = as_f32(p_a), as_f32(o_a)
    B, T, A = o_a.shape
    y = torch.empty((B, T), dtype=torch.float32, device=o_a.device)
    N.call("carca_knn_score", N.f32p(y), N.f32p(p_a), N.f32p(o_a), B, T, p_a.shape[1], A, T, 0, N.stream())
    return y
