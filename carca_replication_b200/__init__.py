"""carca_replication_b200 — B200 (sm_100a) implementation of CARCA's hot path.

Drop-in for the reference's `src` package on that path: same module / function names
(`abstract`, `carca`, `train`, `utils`), same signatures, same `state_dict` layout.  All compute
runs in libcarca_b200.so (hand-written CUDA, C ABI in include/carca_b200.h); there is no CPU or
PyTorch fallback.
"""
from . import abstract, carca, train, utils  # noqa: F401
from .attrs import ItemAttrTable  # noqa: F401
from .carca import (  # noqa: F401
    CARCA, AllEmbedding, AttrCtxEmbedding, AttrEmbedding, BinaryCrossEntropy, CrossAttentionBlock, DotProduct,
    IdEmbedding, IdentityEncoding, LearnableEncoding, MLPIdEmbedding, MultiHeadAttention, PositionalEncoding,
    SelfAttentionBlock, WeightedDotProduct,
)
from .knn import KNN  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .train import compute_HR, compute_NDCG, evaluate, load_checkpoint, save_checkpoint  # noqa: F401
from .utils import get_mask, to  # noqa: F401

__all__ = [
    "CARCA", "AllEmbedding", "AttrCtxEmbedding", "AttrEmbedding", "IdEmbedding", "MLPIdEmbedding",
    "WeightedDotProduct", "KNN", "BinaryCrossEntropy", "CrossAttentionBlock", "DotProduct", "IdentityEncoding",
    "LearnableEncoding", "MultiHeadAttention", "PositionalEncoding", "SelfAttentionBlock", "ItemAttrTable",
    "compute_HR", "compute_NDCG", "evaluate", "get_mask", "to", "FusedAdam",
]
