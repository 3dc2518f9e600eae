"""Device-resident item -> attribute table.

The reference materialises attributes per sequence position on the host (`attrs[item]` copied
into dense [B, N, A] arrays, src/data.py:119-131; 1 GB per 256-user Beauty batch) and passes
that tensor as `a` to Embedding.forward (src/abstract.py:22).  The table the copies come from,
`load_attrs()` (src/data.py:28-35: [n_items, A] with a zero row for the <pad> item), is static,
so here it lives on the GPU once — CSR for multi-hot / sparse attributes (Beauty: ~8 of 6,507
set), dense for image-feature-like attributes (Men: 512 floats) — and the embedding kernel reads
attribute rows by item id.  Both forms give the same result as the dense per-position tensor.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn


class ItemAttrTable(nn.Module):
    """Buffers are non-persistent: the reference's state_dict keys stay untouched."""

    def __init__(self, n_items: int, n_attrs: int, rowptr: Optional[torch.Tensor] = None,
                 cols: Optional[torch.Tensor] = None, vals: Optional[torch.Tensor] = None,
                 dense: Optional[torch.Tensor] = None):
        super().__init__()
        self.n_items, self.n_attrs = int(n_items), int(n_attrs)
        self.is_sparse = dense is None
        if self.is_sparse:
            if rowptr is None or cols is None or vals is None:
                raise ValueError("ItemAttrTable: CSR form needs rowptr, cols and vals")
            if rowptr.numel() != n_items + 1:
                raise ValueError("ItemAttrTable: rowptr must have n_items + 1 entries")
            self.register_buffer("rowptr", rowptr.to(torch.int32).contiguous(), persistent=False)
            self.register_buffer("cols", cols.to(torch.int32).contiguous(), persistent=False)
            self.register_buffer("vals", vals.to(torch.float32).contiguous(), persistent=False)
            self.dense = None
        else:
            if tuple(dense.shape) != (n_items, n_attrs):
                raise ValueError("ItemAttrTable: dense table must be [n_items, n_attrs]")
            self.register_buffer("dense", dense.to(torch.float32).contiguous(), persistent=False)
            self.rowptr = self.cols = self.vals = None

    @property
    def nnz(self) -> int:
        return int(self.cols.numel()) if self.is_sparse else self.n_items * self.n_attrs

    @classmethod
    def from_dense(cls, attrs, sparse: Optional[bool] = None) -> "ItemAttrTable":
        """attrs: [n_items, A] array as returned by the reference's load_attrs (row 0 = pad)."""
        a = attrs.detach().cpu().numpy() if isinstance(attrs, torch.Tensor) else np.asarray(attrs)
        a = np.ascontiguousarray(a, dtype=np.float32)
        n_items, n_attrs = a.shape
        if sparse is None:
            sparse = np.count_nonzero(a) < 0.25 * a.size
        if not sparse:
            return cls(n_items, n_attrs, dense=torch.from_numpy(a))
        rows, cols = np.nonzero(a)
        rowptr = np.zeros(n_items + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr)
        return cls(n_items, n_attrs, rowptr=torch.from_numpy(rowptr), cols=torch.from_numpy(cols.astype(np.int64)),
                   vals=torch.from_numpy(a[rows, cols]))

    @classmethod
    def from_csr(cls, n_items: int, n_attrs: int, rowptr, cols, vals) -> "ItemAttrTable":
        return cls(n_items, n_attrs, rowptr=torch.as_tensor(rowptr), cols=torch.as_tensor(cols),
                   vals=torch.as_tensor(vals))

    def gather_dense(self, ids: torch.Tensor) -> torch.Tensor:
        """Host/test helper: the dense [..., A] tensor the reference loader would have built."""
        flat = ids.reshape(-1).long().cpu()
        if not self.is_sparse:
            return self.dense.cpu()[flat].reshape(*ids.shape, self.n_attrs)
        out = torch.zeros((flat.numel(), self.n_attrs), dtype=torch.float32)
        rp, cl, vl = self.rowptr.cpu().long(), self.cols.cpu().long(), self.vals.cpu()
        starts = rp[flat]
        counts = rp[flat + 1] - starts
        total = int(counts.sum())
        if total:
            row = torch.repeat_interleave(torch.arange(flat.numel()), counts)
            within = torch.arange(total) - torch.repeat_interleave(counts.cumsum(0) - counts, counts)
            src = torch.repeat_interleave(starts, counts) + within
            out[row, cl[src]] = vl[src]
        return out.reshape(*ids.shape, self.n_attrs)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:  # pragma: no cover - convenience only
        return self.gather_dense(ids)
