"""Whole-model inference path: carca_eval_prepare + carca_eval_forward (one fused kernel).

Used by CARCA.forward when the model is in eval mode under torch.no_grad(), every sub-module is one
of this package's classes and attributes come from a device-resident ItemAttrTable: the one-kernel
tensor-core forward (`forward`: d = 64, L <= 64), the packed-rows pipeline (`forward_rows`: any window up
to 256 positions, d in {32, 64, 128, 256}, fp32 or bf16) and the tensor-core full-catalog kernel
(`catalog_counts`).  Anything else keeps the per-op entry points.  The "plan" (folded item table + packed
projection weights) is rebuilt only when a parameter changes (tensor versions + `weights_epoch`, which
raw-pointer optimizers bump), e.g. once per evaluate() during training.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _native as N
from .attrs import ItemAttrTable
from .ops import _attr_source, _c, _embed_params, _struct, as_f32, as_ids

_plans: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()

# kernel variant: 0 = best available (tensor-core kernel when the shape allows), 1 = fp32 FFMA
# kernel, 2 = tcgen05 tensor-core kernel (two-head cross-attention decoder as an fp32 loop: one candidate row per
# thread inside the kernel, or — catalog mode — a separate decoder kernel over all (user, candidate pair) items),
# 3 = the same kernel with that decoder on tcgen05 score MMAs over 128-row tiles, 4 / 5 = the in-kernel row / pair
# loop forced, 6 = the separate decoder kernel forced (needs one context row per user; otherwise as 2).  Variants 0 / 2
# switch to the tcgen05 decoder on the device when the batch's profiles are dense (> 24 valid positions per user).
# CARCA_FUSED_VARIANT overrides (benchmark comparisons).
VARIANT = int(__import__("os").environ.get("CARCA_FUSED_VARIANT", "0"))

MAX_L, MAX_L_TC, MAX_CTX, MAX_BLOCKS, WIDTH = 52, 256, 8, 8, 64
BIN_ROWS = 64      # rows of one packed bin of the tensor-core kernel: the most valid positions a user may have


def supported(model, seq_len: int, n_ctx: int) -> bool:
    from . import carca as M

    emb, dec = model.embeds, model.decoder
    if not isinstance(emb, M.AllEmbedding) or not hasattr(emb.enc, "table"):
        return False
    blocks = list(model.encoder)
    if not blocks or len(blocks) > MAX_BLOCKS or not all(isinstance(b, M.SelfAttentionBlock) for b in blocks):
        return False
    H = blocks[0].attn.H
    if any(b.attn.H != H or bool(b.residual) != bool(blocks[0].residual) for b in blocks):
        return False
    if isinstance(dec, M.CrossAttentionBlock):
        if dec.attn.H != H:
            return False
    elif not isinstance(dec, M.DotProduct):
        return False
    d = emb.d
    if d != WIDTH or n_ctx > MAX_CTX or d % H != 0:
        return False
    ffma = seq_len <= MAX_L and (d // H) % 4 == 0
    tc = seq_len <= MAX_L_TC and H in (2, 4) and N.is_device_tensor(emb.items_embed.weight)
    return ffma or tc


def _model_params(model, table: ItemAttrTable, WfT: Optional[Tensor], n_ctx: int):
    """carca_model_params + the Python objects that must stay alive while it is used."""
    from . import carca as M

    emb = model.embeds
    E, Wf, bf = _c(emb.items_embed.weight), _c(emb.feats_embed.weight), _c(emb.feats_embed.bias)
    Wj, bj = _c(emb.joint_embed.weight), _c(emb.joint_embed.bias)
    A = Wf.shape[1] - n_ctx
    pos = emb.enc.table(MAX_L_TC)
    pos = None if pos is None else _c(pos)
    blocks = list(model.encoder)
    arr = (N.BlockParams * len(blocks))()
    keep: List = [E, Wf, bf, Wj, bj, pos, WfT, arr]
    for i, b in enumerate(blocks):
        ps = tuple(_c(t) for t in b._params())
        keep.append(ps)
        for name, t in zip(N.BLOCK_PARAM_NAMES, ps):
            setattr(arr[i], name, N.f32p(t))
    m = N.ModelParams()
    m.embed = _embed_params(E, Wf, WfT, bf, Wj, bj, pos, A, n_ctx)
    m.n_blocks, m.n_heads = len(blocks), blocks[0].attn.H
    m.residual_sa = int(bool(blocks[0].residual))
    m.blocks = C.cast(arr, C.POINTER(N.BlockParams))
    ng, nb = _c(model.norm.weight), _c(model.norm.bias)
    keep += [ng, nb]
    m.norm_g, m.norm_b = N.f32p(ng), N.f32p(nb)
    dec = model.decoder
    if isinstance(dec, M.CrossAttentionBlock):
        a = dec.attn
        cp = tuple(_c(t) for t in (a.WQ.weight, a.WQ.bias, a.WK.weight, a.WK.bias, a.WV.weight, a.WV.bias,
                                   dec.ffn.weight, dec.ffn.bias))
        keep.append(cp)
        m.decoder_kind, m.residual_ca = 1, int(bool(dec.residual))
        m.cross = _struct(N.CrossParams, N.CROSS_PARAM_NAMES, cp)
    else:
        m.decoder_kind, m.residual_ca = 0, 0
    return m, keep


# Bumped whenever one of this package's modules is moved / cast (`Module._apply`, i.e. .to() / .cuda() / .float()):
# those REPLACE parameter tensors, which the cached tensor lists below would not notice.  In-place updates
# (optimizer steps, load_state_dict) are caught through the tensors' version counters.
_struct_epoch = [0]


def invalidate_plans() -> None:
    _struct_epoch[0] += 1


# Weights epoch: bumped by everything that writes parameters through RAW POINTERS, which tensor version counters do
# not see — FusedAdam.step() (csrc/optim.cuh) and every GraphedTrainStep replay.  Part of every cache key derived
# from the weights (inference plans here, AllEmbedding's folded table in carca.py).
_weights_epoch = [0]


def bump_weights_epoch() -> None:
    _weights_epoch[0] += 1


def weights_epoch() -> int:
    return _weights_epoch[0]


class _Entry:
    __slots__ = ("epoch", "tensors", "key", "plan", "status", "m", "keep", "n_ctx", "rows", "rows_spare")


def _tensors_of(model) -> list:
    return list(model.parameters()) + list(model.buffers())


def _version_key(tensors, table: ItemAttrTable) -> Tuple:
    return (id(table), _weights_epoch[0]) + tuple(t._version for t in tensors)


def plan_is_current(model) -> bool:
    """True when the cached inference plan of `model` was built from its current weights."""
    ent = _plans.get(model)
    return (ent is not None and ent.epoch == _struct_epoch[0]
            and ent.key[1:] == (_weights_epoch[0],) + tuple(t._version for t in ent.tensors))


def eval_plan(model, table: ItemAttrTable, n_ctx: int):
    """Inference plan for the model's current weights (cached until a parameter changes) and the argument block of
    the forward calls (device pointers of every parameter; rebuilt together with the plan)."""
    ent = _plans.get(model)
    same_struct = ent is not None and ent.epoch == _struct_epoch[0] and ent.n_ctx == n_ctx
    if same_struct and ent.key == _version_key(ent.tensors, table):
        return ent
    # same parameter tensors, new values: the plan is rebuilt INTO THE SAME BUFFER, so that CUDA graphs which captured
    # its address (GraphedEvalStep) score with the new weights after a refresh
    reuse = ent.plan if same_struct and ent.key[0] == id(table) else None
    tensors = _tensors_of(model)
    key = _version_key(tensors, table)
    emb = model.embeds
    dev = emb.items_embed.weight.device
    Wf = _c(emb.feats_embed.weight)
    g, AC = Wf.shape
    WfT = None
    if table.is_sparse:
        WfT = torch.empty((AC, g), dtype=torch.float32, device=dev)
        N.call("carca_transpose", N.f32p(WfT), N.f32p(Wf), g, AC, 0, N.stream())
    m, keep = _model_params(model, table, WfT, n_ctx)
    n_floats = N.lib().carca_eval_plan_floats(C.byref(m))
    plan = reuse if reuse is not None and reuse.numel() == n_floats and reuse.device == dev else \
        torch.empty(n_floats, dtype=torch.float32, device=dev)
    scratch = torch.empty((emb.items_embed.weight.shape[0], g), dtype=torch.float32, device=dev)
    src = _attr_source(table, None)
    N.call("carca_eval_prepare", N.f32p(plan), N.f32p(scratch), C.byref(m), C.byref(src), N.stream())
    old = _plans.get(model)
    ent = _Entry()
    ent.rows = None                                  # bf16 plan (rows_plan), built on first use
    ent.rows_spare = None if (old is None or reuse is None) else (old.rows if old.rows is not None else old.rows_spare)
    ent.epoch, ent.tensors, ent.key, ent.n_ctx = _struct_epoch[0], tensors, key, n_ctx
    ent.plan = plan
    ent.status = old.status if reuse is not None and plan is reuse else torch.zeros(1, dtype=torch.int32, device=dev)
    ent.m, ent.keep = _model_params(model, table, None, n_ctx)     # the forward calls' argument block
    _plans[model] = ent
    return ent


_scratch_cache: dict = {}


def _scratch(B: int, device) -> Tensor:
    """Row-packing scratch of the tensor-core kernel (carca_eval_scratch_bytes), reused per batch size.
    Calls on one stream are ordered, so a single buffer per (device, B) is enough."""
    key = (str(device), int(B))
    buf = _scratch_cache.get(key)
    if buf is None:
        if len(_scratch_cache) > 8:
            _scratch_cache.clear()
        nbytes = int(N.lib().carca_eval_scratch_bytes(int(B)))
        buf = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)
        _scratch_cache[key] = buf
    return buf


def status_word(model, device) -> Tensor:
    """float64[1] device tensor: the status word of the model's inference plan (bit 0: a tcgen05 completion wait
    timed out), 0 when the model has no plan.  No host sync — evaluate() appends it to its accumulators."""
    hit = _plans.get(model)
    if hit is None:
        return torch.zeros(1, dtype=torch.float64, device=device)
    return hit.status.to(torch.float64)


def mma_timed_out(model) -> bool:
    """True if a tensor-core completion wait ever timed out for this model's plan (device sync)."""
    hit = _plans.get(model)
    return bool(hit is not None and (int(hit.status.item()) & 1) != 0)


def forward(model, profile, targets: Sequence, variant: Optional[int] = None, dbg: Optional[Tensor] = None,
            dbg_stage: int = 0) -> Tensor:
    """CARCA.forward in eval mode (src/carca.py:411-431) through the fused kernel -> [B, sum(T)]."""
    p_x, p_a, p_c = profile
    table = p_a if isinstance(p_a, ItemAttrTable) else model.embeds.attr_table
    N.require_device(p_x, p_c)
    p_x, p_c = as_ids(p_x), as_f32(p_c)
    B, L = p_x.shape
    n_ctx = p_c.shape[-1]
    per_user_ctx = False
    if len(targets) == 1:
        o_x, o_c = as_ids(targets[0][0]), targets[0][2]
        if o_c.dim() == 3 and o_c.stride(2) == 1 and (o_c.shape[1] == 1 or o_c.stride(1) == 0):
            # [B,T,C] expanded from one context row per user (the positive's context given to every
            # sampled negative, src/data.py:185): read the [B,C] base, never materialise the copies
            o_c, per_user_ctx = as_f32(o_c[:, 0, :]), True
        else:
            o_c = as_f32(o_c)
    else:   # eval-mode decoders score every candidate independently, so tuples simply concatenate
        o_x = torch.cat([as_ids(t[0]) for t in targets], dim=1)
        o_c = torch.cat([as_f32(t[2]) for t in targets], dim=1)
    T = o_x.shape[1]
    ent = eval_plan(model, table, n_ctx)
    plan, status, m = ent.plan, ent.status, ent.m
    y = torch.empty((B, T), dtype=torch.float32, device=p_x.device)
    v = (VARIANT if variant is None else int(variant)) | (0x100 if per_user_ctx else 0)
    N.call("carca_eval_forward_opts", N.f32p(y), T, 0, N.f32p(plan), C.byref(m), N.i32p(p_x), N.f32p(p_c),
           N.i32p(o_x), N.f32p(o_c), B, L, T, v, N.i32p(status),
           None if dbg is None else N.f32p(dbg), int(dbg_stage), _scratch(B, p_x.device).data_ptr(), N.stream())
    return y


def forward_catalog(model, profile, ctx_user: Tensor, item_lo: int, n_cand: int,
                    variant: Optional[int] = None) -> Tensor:
    """Scores of the contiguous item range [item_lo, item_lo + n_cand) for every user -> [B, n_cand]
    (carca_eval_forward_catalog; eval-mode CARCA.forward over candidate chunks, src/carca.py:424-431)."""
    p_x, p_a, p_c = profile
    table = p_a if isinstance(p_a, ItemAttrTable) else model.embeds.attr_table
    N.require_device(p_x, p_c, ctx_user)
    p_x, p_c, ctx_user = as_ids(p_x), as_f32(p_c), as_f32(ctx_user)
    B, L = p_x.shape
    n_ctx = p_c.shape[-1]
    ent = eval_plan(model, table, n_ctx)
    plan, status, m = ent.plan, ent.status, ent.m
    y = torch.empty((B, n_cand), dtype=torch.float32, device=p_x.device)
    N.call("carca_eval_forward_catalog", N.f32p(y), n_cand, 0, N.f32p(plan), C.byref(m), N.i32p(p_x), N.f32p(p_c),
           N.f32p(ctx_user), int(item_lo), int(n_cand), B, L, VARIANT if variant is None else int(variant),
           N.i32p(status), _scratch(B, p_x.device).data_ptr(), N.stream())
    return y


# ------------------------------------------------------------------------------- bf16 packed-rows pipeline
ROWS_MAX_L = 256
ROWS_SHAPES = {"bf16": ((64, 256), (32, 64)), "fp32": ((32, 64, 128, 256), (16, 32, 64))}   # widths, head widths


def rows_supported(model, seq_len: int, n_ctx: int, precision: str = "bf16") -> bool:
    """Shape range of the packed-rows pipeline (csrc/rows_bf16.cuh): L <= 256 with any number of valid positions per
    user, stock blocks, cross-attention or dot-product decoder; bf16 flavour d in {64, 256} with head width 32 / 64,
    fp32 flavour d in {32, 64, 128, 256} with head width 16 / 32 / 64."""
    from . import carca as M

    emb, dec = model.embeds, model.decoder
    if not isinstance(emb, M.AllEmbedding) or not hasattr(emb.enc, "table"):
        return False
    blocks = list(model.encoder)
    if not blocks or len(blocks) > MAX_BLOCKS or not all(isinstance(b, M.SelfAttentionBlock) for b in blocks):
        return False
    H = blocks[0].attn.H
    if any(b.attn.H != H or bool(b.residual) != bool(blocks[0].residual) for b in blocks):
        return False
    if isinstance(dec, M.CrossAttentionBlock):
        if dec.attn.H != H:
            return False
    elif not isinstance(dec, M.DotProduct):
        return False
    d = emb.d
    widths, head_widths = ROWS_SHAPES[precision]
    if (d, H) == (256, 16):
        return False
    return (d in widths and d % H == 0 and d // H in head_widths and n_ctx <= MAX_CTX
            and seq_len <= ROWS_MAX_L and N.is_device_tensor(emb.items_embed.weight))


def rows_plan(model, table: ItemAttrTable, n_ctx: int):
    """(entry, bf16 plan buffer) for the model's current weights: the fp32 plan's folded tables converted / folded
    further for the bf16 pipeline (carca_rows_prepare).  Rebuilt with the fp32 plan, into the same buffer."""
    ent = eval_plan(model, table, n_ctx)
    if ent.rows is None:
        nbytes = int(N.lib().carca_rows_plan_bytes(C.byref(ent.m)))
        dev = ent.plan.device
        buf = ent.rows_spare
        if buf is None or buf.numel() != nbytes or buf.device != dev:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        N.call("carca_rows_prepare", buf.data_ptr(), N.f32p(ent.plan), C.byref(ent.m), N.stream())
        ent.rows, ent.rows_spare = buf, None
    return ent


_rows_scratch_cache: dict = {}


def _rows_scratch(ent, B: int, L: int, device, lane: int = 0) -> Tensor:
    key = (str(device), int(B), int(L), int(ent.m.embed.d), int(ent.m.n_heads), int(lane))
    buf = _rows_scratch_cache.get(key)
    if buf is None:
        if len(_rows_scratch_cache) > 12:
            _rows_scratch_cache.clear()
        nbytes = int(N.lib().carca_rows_scratch_bytes(C.byref(ent.m), int(B), int(L)))
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _rows_scratch_cache[key] = buf
    return buf


def forward_rows(model, profile, targets: Sequence, precision: str = "bf16", cat_lo: int = 0, n_cand: int = 0,
                 ctx_user: Optional[Tensor] = None) -> Tensor:
    """CARCA.forward in eval mode (src/carca.py:411-431) through the packed-rows pipeline -> [B, sum(T)]; precision
    "bf16" or "fp32".  With cat_lo > 0: catalog mode, scores of items [cat_lo, cat_lo + n_cand) with one context row
    per user."""
    p_x, p_a, p_c = profile
    table = p_a if isinstance(p_a, ItemAttrTable) else model.embeds.attr_table
    N.require_device(p_x, p_c)
    p_x, p_c = as_ids(p_x), as_f32(p_c)
    B, L = p_x.shape
    n_ctx = p_c.shape[-1]
    per_user_ctx = False
    if cat_lo > 0:
        o_x, o_c, per_user_ctx, T = None, as_f32(ctx_user), True, int(n_cand)
    elif len(targets) == 1:
        o_x, o_c = as_ids(targets[0][0]), targets[0][2]
        if o_c.dim() == 3 and o_c.stride(2) == 1 and (o_c.shape[1] == 1 or o_c.stride(1) == 0):
            o_c, per_user_ctx = as_f32(o_c[:, 0, :]), True      # expanded view of one context row per user
        else:
            o_c = as_f32(o_c)
        T = o_x.shape[1]
    else:
        o_x = torch.cat([as_ids(t[0]) for t in targets], dim=1)
        o_c = torch.cat([as_f32(t[2]) for t in targets], dim=1)
        T = o_x.shape[1]
    ent = rows_plan(model, table, n_ctx)
    y = torch.empty((B, T), dtype=torch.float32, device=p_x.device)
    prec = {"bf16": 0, "fp32": 1}[precision]

    def run(b0: int, b1: int, lane: int) -> None:
        N.call("carca_rows_eval_forward", N.f32p(y[b0:b1]), T, 0, ent.rows.data_ptr(), N.f32p(ent.plan), C.byref(ent.m),
               N.i32p(p_x[b0:b1]), N.f32p(p_c[b0:b1]), None if o_x is None else N.i32p(o_x[b0:b1]), N.f32p(o_c[b0:b1]),
               b1 - b0, L, T, int(per_user_ctx), int(cat_lo), prec, N.i32p(ent.status),
               _rows_scratch(ent, b1 - b0, L, p_x.device, lane).data_ptr(), N.stream())

    lanes = rows_lanes(B, L)
    if lanes <= 1:
        run(0, B, 0)
        return y
    # Users are independent, and with a few thousand users every stage of the pipeline is a short kernel bound by its
    # own launch / fill / drain latency, not by throughput: slices of the batch run the whole pipeline side by side
    # on separate streams (fork / join by events; capturable), each with its own scratch.
    main = torch.cuda.current_stream(p_x.device)
    fork = torch.cuda.Event()
    fork.record(main)
    step = -(-B // lanes)
    for i in range(lanes):
        b0, b1 = i * step, min(B, (i + 1) * step)
        if b0 >= b1:
            break
        if i == 0:
            run(b0, b1, 0)
            continue
        side = _lane_stream(p_x.device, i)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            run(b0, b1, i)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)
    return y


ROWS_LANES = 1            # slices of a large batch that run the rows pipeline concurrently (see forward_rows)
ROWS_LANE_MIN_USERS = 1024
_lane_streams: dict = {}


def rows_lanes(B: int, L: int) -> int:
    return max(1, min(int(ROWS_LANES), B // ROWS_LANE_MIN_USERS))


def _lane_stream(device, i: int):
    key = (str(device), i)
    st = _lane_streams.get(key)
    if st is None:
        st = _lane_streams[key] = torch.cuda.Stream(device=device)
    return st


def catalog_tc_supported(model, seq_len: int, n_ctx: int) -> bool:
    """Tensor-core full-catalog kernel (csrc/catalog_tc.cuh): d = 64, L <= 128, fp32, dot decoder or two-head
    cross-attention."""
    from . import carca as M

    if not rows_supported(model, seq_len, n_ctx, "fp32") or model.embeds.d != 64 or seq_len > 128:
        return False
    if model.eval_dtype != "fp32" or N.is_emulated():
        return False
    dec = model.decoder
    return isinstance(dec, M.DotProduct) or dec.attn.H == 2


_cat_scratch_cache: dict = {}


def catalog_counts(model, profile, pos_item: Tensor, pos_ctx: Tensor, item_lo: int, item_hi: int, counts: Tensor) -> None:
    """counts[b] += items of [item_lo, item_hi) ranked before user b's positive (carca_rows_catalog_counts): encoder on
    the packed-rows fp32 pipeline, scores + softmax + sigmoid + comparison in one tcgen05 kernel, no score matrix."""
    p_x, p_a, p_c = profile
    table = p_a if isinstance(p_a, ItemAttrTable) else model.embeds.attr_table
    N.require_device(p_x, p_c, pos_item, pos_ctx, counts)
    p_x, p_c, pos_item, pos_ctx = as_ids(p_x), as_f32(p_c), as_ids(pos_item), as_f32(pos_ctx)
    B, L = p_x.shape
    ent = rows_plan(model, table, p_c.shape[-1])
    key = (str(p_x.device), B, L)
    buf = _cat_scratch_cache.get(key)
    if buf is None:
        if len(_cat_scratch_cache) > 4:
            _cat_scratch_cache.clear()
        nbytes = int(N.lib().carca_rows_catalog_scratch_bytes(C.byref(ent.m), B, L))
        buf = _cat_scratch_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=p_x.device)
    N.call("carca_rows_catalog_counts", N.i32p(counts), ent.rows.data_ptr(), N.f32p(ent.plan), C.byref(ent.m), N.i32p(p_x),
           N.f32p(p_c), N.f32p(pos_ctx), N.i32p(pos_item), int(item_lo), int(item_hi - item_lo), B, L, N.i32p(ent.status),
           buf.data_ptr(), N.stream())
