"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-user batch construction
(r-papso/carca-replication, src/data.py), without the attribute tensors (those stay in the item
table on our side) and with the sampled negatives passed in, so every output is deterministic.

Pinned by tests/golden/data_sequences.npz, which tests/golden/make_golden_data.py produced by
calling the REAL get_train_sequences / get_test_sequences of /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def pad_profile(n: int, max_len: int, mode: str, test: bool) -> range:
    """src/data.py:53-74 on the profile LENGTH: the index range it selects."""
    if mode not in ["train", "val", "test"]:
        raise ValueError(f"Invalid mode: {mode}")
    start, end = 0, 0
    if mode == "train" and n > 1:
        ex = 2 if test else 1
        start, end = max(0, n - ex - max_len - 1), max(1, n - ex)
    if mode == "val" and n > 2:
        ex = 1 if test else 0
        start, end = max(0, n - ex - max_len - 1), max(2, n - ex)
    if mode == "test" and n > 3:
        start, end = max(0, n - max_len - 1), max(3, n)
    return range(start, end)


def test_sequences(profile: Sequence[int], ctx_rows: np.ndarray, seq_len: int, negs: Sequence[int], mode: str,
                   test: bool):
    """src/data.py:140-192 -> p_x [L], p_c [L,C], o_x [1+len(negs)], o_c [1+len(negs), C], y_true.
    ctx_rows[j] is ctx[(user, profile[j])]."""
    C = ctx_rows.shape[1]
    T = len(negs) + 1
    p_x, p_c = np.zeros(seq_len, np.int32), np.zeros((seq_len, C), np.float32)
    o_x, o_c = np.zeros(T, np.int32), np.zeros((T, C), np.float32)
    idxs = list(pad_profile(len(profile), seq_len, mode, test))
    one_out = idxs[-1]                                                       # :162
    o_x[0], o_c[0] = profile[one_out], ctx_rows[one_out]
    for i, pi in enumerate(reversed(idxs[:-1])):                             # :172
        p_x[seq_len - i - 1], p_c[seq_len - i - 1] = profile[pi], ctx_rows[pi]
    for i, oi in enumerate(negs, start=1):                                   # :180
        o_x[i], o_c[i] = oi, ctx_rows[one_out]                               # :185
    y = np.zeros(T, np.int32)
    y[0] = 1
    return p_x, p_c, o_x, o_c, y


def train_sequences(profile: Sequence[int], ctx_rows: np.ndarray, seq_len: int, negs: Sequence[int], test: bool):
    """src/data.py:90-137 -> p_x [L], p_c [L,C], o_x [2L], o_c [2L,C], y_true [2L]."""
    C = ctx_rows.shape[1]
    p_x, p_c = np.zeros(seq_len, np.int32), np.zeros((seq_len, C), np.float32)
    o_x, o_c = np.zeros(2 * seq_len, np.int32), np.zeros((2 * seq_len, C), np.float32)
    idxs = list(pad_profile(len(profile), seq_len, "train", test))
    for i, pi in enumerate(reversed(idxs[:-1])):                             # :111
        idx = seq_len - i - 1
        p_x[idx], o_x[idx], o_x[seq_len + idx] = profile[pi], profile[pi + 1], negs[i]
        p_c[idx], o_c[idx], o_c[seq_len + idx] = ctx_rows[pi], ctx_rows[pi + 1], ctx_rows[pi + 1]   # :130
    y = np.zeros(2 * seq_len, np.int32)
    y[np.where(p_x > 0)] = 1                                                 # :134-135
    return p_x, p_c, o_x, o_c, y
