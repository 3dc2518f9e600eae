"""ctypes binding of libcarca_b200.so (C ABI declared in include/carca_b200.h).

The library is the product: there is no Python/torch fallback.  If it is missing, or a tensor
is not on a CUDA device, the call raises.  torch supplies device memory and the current stream
only; tensors cross the boundary as raw pointers + sizes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcarca_b200.so")

_LIB = None

vp, i32, i64, u32, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_float


class AttrSource(C.Structure):
    _fields_ = [("kind", i32), ("csr_rowptr", vp), ("csr_cols", vp), ("csr_vals", vp), ("dense", vp)]


class EmbedParams(C.Structure):
    _fields_ = [("n_items", i32), ("d", i32), ("g", i32), ("n_attrs", i32), ("n_ctx", i32),
                ("items_embed", vp), ("feats_w", vp), ("feats_wT", vp), ("feats_b", vp),
                ("joint_w", vp), ("joint_b", vp), ("pos", vp), ("pos_len", i32)]


class EmbedGrads(C.Structure):
    _fields_ = [(n, vp) for n in ("items_embed", "feats_w", "feats_b", "joint_w", "joint_b", "pos")]


BLOCK_PARAM_NAMES = ("ln1_g", "ln1_b", "wq", "bq", "wk", "bk", "wv", "bv", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")
BLOCK_SAVED_NAMES = ("qn", "mean1", "rstd1", "Q", "K", "V", "s", "mean2", "rstd2", "s2", "a1")
CROSS_PARAM_NAMES = ("wq", "bq", "wk", "bk", "wv", "bv", "wf", "bf")
CROSS_SAVED_NAMES = ("Q", "K", "V", "s")


class BlockParams(C.Structure):
    _fields_ = [(n, vp) for n in BLOCK_PARAM_NAMES]


class BlockSaved(C.Structure):
    _fields_ = [(n, vp) for n in BLOCK_SAVED_NAMES]


class CrossParams(C.Structure):
    _fields_ = [(n, vp) for n in CROSS_PARAM_NAMES]


class CrossSaved(C.Structure):
    _fields_ = [(n, vp) for n in CROSS_SAVED_NAMES]


class ModelParams(C.Structure):
    _fields_ = [("embed", EmbedParams), ("n_blocks", i32), ("n_heads", i32), ("residual_sa", i32),
                ("residual_ca", i32), ("decoder_kind", i32), ("blocks", C.POINTER(BlockParams)), ("norm_g", vp),
                ("norm_b", vp), ("cross", CrossParams)]


class TrainCore(C.Structure):
    _fields_ = [("B", i32), ("L", i32), ("n_heads", i32), ("n_blocks", i32), ("n_tuples", i32), ("decoder_kind", i32),
                ("residual_sa", i32), ("residual_ca", i32), ("p_drop", f32), ("seed", u64), ("p_x", vp),
                ("o_x", vp * 2), ("p_e", vp), ("o_e", vp * 2), ("blocks", C.POINTER(BlockParams)), ("norm_g", vp),
                ("norm_b", vp), ("cross", CrossParams), ("rows", vp), ("saved", vp), ("embed", C.POINTER(EmbedParams)),
                ("attrs", C.POINTER(AttrSource)), ("p_c", vp), ("o_c", vp * 2), ("fold", vp)]


class AdamTensor(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("step", vp), ("numel", i64)]


class Interactions(C.Structure):
    _fields_ = [("rowptr", vp), ("items", vp), ("ctx", vp), ("n_users", i32), ("n_ctx", i32)]


P = C.POINTER

# name -> argtypes (restype is always int unless noted); mirrors include/carca_b200.h one to one
SIGNATURES = {
    "carca_abi_version": [],
    "carca_launch_count": [],
    "carca_set_seed_source": [vp],
    "carca_transpose": [vp, vp, i32, i32, i32, vp],
    "carca_padding_mask": [vp, vp, i64, vp],
    "carca_padding_mask_f32": [vp, vp, i64, vp],
    "carca_embed_fwd": [vp, vp, P(EmbedParams), P(AttrSource), vp, vp, vp, i32, i32, i32, vp],
    "carca_embed_bwd": [P(EmbedGrads), vp, vp, P(EmbedParams), P(AttrSource), vp, vp, vp, i32, i32, i32,
                        vp, vp, vp, vp],
    "carca_feats_fwd": [vp, P(EmbedParams), P(AttrSource), vp, vp, i32, vp],
    "carca_feats_bwd": [vp, vp, vp, P(EmbedParams), P(AttrSource), vp, vp, i32, vp, vp],
    "carca_gather_rows_fwd": [vp, vp, vp, f32, i32, i32, vp],
    "carca_gather_rows_bwd": [vp, vp, vp, f32, i32, i32, vp],
    "carca_pos_mask_fwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_embed_folded_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "carca_pos_mask_bwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_wdot_score_fwd": [vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, i64, i32, vp],
    "carca_wdot_score_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, i64, i32, vp],
    "carca_knn_score": [vp, vp, vp, i32, i32, i32, i32, i64, i32, vp],
    "carca_dropout": [vp, vp, i64, f32, u64, u32, vp],
    "carca_layernorm_fwd": [vp, vp, vp, vp, vp, vp, i32, i32, vp],
    "carca_layernorm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_linear_fwd": [vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "carca_linear_bwd_input": [vp, vp, vp, i32, i32, i32, i32, vp],
    "carca_linear_bwd_weight": [vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_attention_fwd": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, u64, u32, vp],
    "carca_attention_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, u64, u32,
                            vp],
    "carca_sa_block_fwd": [vp, P(BlockSaved), vp, vp, P(BlockParams), i32, i32, i32, i32, i32, f32, u64, i32, vp],
    "carca_sa_block_bwd": [vp, P(BlockParams), vp, vp, vp, P(BlockParams), P(BlockSaved), i32, i32, i32, i32, i32,
                           f32, u64, i32, vp, vp],
    "carca_dot_score_fwd": [vp, vp, vp, i32, i32, i32, i32, i32, i64, i32, vp],
    "carca_dot_score_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i32, vp],
    "carca_cross_score_fwd": [vp, P(CrossSaved), vp, vp, vp, vp, P(CrossParams), i32, i32, i32, i32, i32, i32, i32,
                              f32, u64, u32, i64, i32, vp],
    "carca_cross_score_bwd": [vp, vp, P(CrossParams), vp, vp, P(CrossSaved), vp, vp, vp, vp, P(CrossParams), i32,
                              i32, i32, i32, i32, i32, i32, f32, u64, u32, i64, i32, vp, vp],
    "carca_train_core_set_ticks": [vp, i32],
    "carca_train_core_rows_ints": [i32],
    "carca_train_core_saved_floats": [i32, i32, i32],
    "carca_train_core_fwd": [vp, i64, P(TrainCore), vp],
    "carca_train_core_fold_floats": [P(EmbedParams)],
    "carca_train_core_bwd": [vp, vp, vp, P(BlockParams), vp, vp, P(CrossParams), P(EmbedGrads), vp, vp, i64,
                             P(TrainCore), vp],
    "carca_bce_sums": [vp, vp, vp, vp, i64, f32, vp],
    "carca_bce_finalize": [vp, vp, vp],
    "carca_bce_bwd": [vp, vp, vp, vp, vp, vp, i64, f32, vp],
    "carca_rank_metrics": [vp, vp, vp, vp, i32, i32, i64, i64, i32, vp],
    "carca_eval_metrics": [vp, vp, vp, vp, vp, i32, i32, i64, i64, i64, i32, f32, vp],
    "carca_adam_step": [P(AdamTensor), i32, f32, f32, f32, f32, f32, vp],
    "carca_build_eval_batch": [vp, vp, vp, vp, vp, P(Interactions), vp, i32, i32, i32, i32, i32, i32, u64, vp],
    "carca_build_train_batch": [vp, vp, vp, vp, vp, P(Interactions), vp, i32, i32, i32, i32, u64, vp],
    "carca_unpack_windows": [vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_eval_plan_floats": [P(ModelParams)],
    "carca_eval_prepare": [vp, vp, P(ModelParams), P(AttrSource), vp],
    "carca_eval_forward": [vp, i64, i32, vp, P(ModelParams), vp, vp, vp, vp, i32, i32, i32, vp],
    "carca_eval_forward_opts": [vp, i64, i32, vp, P(ModelParams), vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp,
                                vp],
    "carca_eval_scratch_bytes": [i32],
    "carca_eval_forward_catalog": [vp, i64, i32, vp, P(ModelParams), vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp],
    "carca_rows_plan_bytes": [P(ModelParams)],
    "carca_rows_prepare": [vp, vp, P(ModelParams), vp],
    "carca_rows_scratch_bytes": [P(ModelParams), i32, i32],
    "carca_rows_eval_forward": [vp, i64, i32, vp, vp, P(ModelParams), vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp,
                                vp, vp],
    "carca_rows_set_stage_events": [vp, i32],
    "carca_rows_stage_ids": [vp, i32],
    "carca_rows_catalog_scratch_bytes": [P(ModelParams), i32, i32],
    "carca_rows_catalog_counts": [vp, vp, vp, P(ModelParams), vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp],
    "carca_catalog_rank_count": [vp, vp, i64, vp, vp, i32, i32, i32, vp],
    "carca_peer_buffer_bytes": [i64],
    "carca_peer_data_offset": [],
    "carca_peer_alloc": [vp, i64],
    "carca_peer_free": [vp],
    "carca_peer_export": [vp, vp],
    "carca_peer_open": [vp, vp],
    "carca_peer_close": [vp],
    "carca_peer_allreduce": [vp, i64, vp, i32, i32, vp, vp],
    "carca_umma_selftest": [vp, vp, vp, i32, i32, i32, vp, vp],
    "carca_umma_probe": [vp, vp, i32, vp, i32, i32, i32, u32, u32, u32, u32, u32, u32, u32, vp, vp],
}


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restypes; raises AttributeError if the library lacks a declared symbol."""
    lib.carca_last_error.restype = C.c_char_p
    lib.carca_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.carca_launch_count.restype = C.c_int64
    lib.carca_set_seed_source.restype = None
    lib.carca_eval_plan_floats.restype = C.c_int64
    lib.carca_eval_scratch_bytes.restype = C.c_int64
    lib.carca_rows_plan_bytes.restype = C.c_int64
    lib.carca_peer_buffer_bytes.restype = C.c_int64
    lib.carca_peer_data_offset.restype = C.c_int64
    lib.carca_rows_scratch_bytes.restype = C.c_int64
    lib.carca_rows_catalog_scratch_bytes.restype = C.c_int64
    lib.carca_train_core_set_ticks.restype = None
    lib.carca_train_core_rows_ints.restype = C.c_int64
    lib.carca_train_core_saved_floats.restype = C.c_int64
    lib.carca_train_core_fold_floats.restype = C.c_int64
    return lib


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m carca_replication_b200.build` "
                "(nvcc, sm_100a). carca_replication_b200 has no CPU or PyTorch fallback.")
        loaded = bind(C.CDLL(LIB_PATH))
        got = loaded.carca_abi_version()
        if got != 2:
            raise RuntimeError(f"libcarca_b200.so ABI version {got}, expected 2")
        _LIB = loaded
    return _LIB


def call(name: str, *args) -> None:
    """Invoke an entry point; non-zero return -> RuntimeError with the library's message."""
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.carca_last_error().decode()}")


def require_device(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("carca_replication_b200 runs on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")


def is_device_tensor(t: torch.Tensor) -> bool:
    return t.is_cuda


def is_emulated() -> bool:
    """True only under the development emulator fixture of the CPU tests (which has no tcgen05 pipeline)."""
    return False


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def f32p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"expected a contiguous float32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def i32p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if t.dtype != torch.int32 or not t.is_contiguous():
        raise RuntimeError(f"expected a contiguous int32 tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return t.data_ptr()


def anyp(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_contiguous():
        raise RuntimeError("expected a contiguous tensor")
    return t.data_ptr()
