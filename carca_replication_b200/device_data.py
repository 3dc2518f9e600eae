"""Device-side batch construction.

The reference builds every batch on the host: per user, `pad_profile`, rejection-sampled negatives
with Python's `random`, and a per-position copy of `attrs[item]` / `ctx[(user, item)]` into dense
arrays (src/data.py:90-192, 1 GB per 256-user Beauty batch, ~500 users/s).  Here the interaction
log lives on the GPU once (CSR over users) and a kernel emits, per batch, exactly what the B200
path consumes: ids + context, no attribute tensors.

    log = DeviceInteractions.from_profiles(user_ids, profiles, ctx).to(device)
    for batch in DeviceLoader(log, n_items, profile_seq_len=50, target_seq_len=100, mode="val", batch_size=8192):
        p_x, p_a, p_c, o_x, o_a, o_c, y_true = batch          # p_a = o_a = None -> ItemAttrTable
    evaluate(model, DeviceLoader(...), device, k=10)          # drops into src/train.py's loops
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import _native as N

MODES = {"train": 0, "val": 1, "test": 2}


class DeviceInteractions(nn.Module):
    """Interaction log as CSR over users: `items[rowptr[u]:rowptr[u+1]]` chronological, `ctx` row per interaction."""

    def __init__(self, rowptr: Tensor, items: Tensor, ctx: Tensor):
        super().__init__()
        if rowptr.dim() != 1 or items.dim() != 1 or ctx.dim() != 2 or ctx.shape[0] != items.shape[0]:
            raise ValueError("DeviceInteractions: rowptr [U+1], items [nnz], ctx [nnz, C] expected")
        self.register_buffer("rowptr", rowptr.to(torch.int32).contiguous(), persistent=False)
        self.register_buffer("items", items.to(torch.int32).contiguous(), persistent=False)
        self.register_buffer("ctx", ctx.to(torch.float32).contiguous(), persistent=False)

    @property
    def n_users(self) -> int:
        return self.rowptr.numel() - 1

    @classmethod
    def from_profiles(cls, user_ids: Sequence[int], profiles: Dict[int, List[int]],
                      ctx: Dict[Tuple[int, int], np.ndarray]) -> "DeviceInteractions":
        """From what load_profiles / load_ctx return (src/data.py:17-50).  Row u of the log is user_ids[u]."""
        lens = [len(profiles[u]) for u in user_ids]
        rowptr = np.zeros(len(user_ids) + 1, dtype=np.int64)
        np.cumsum(lens, out=rowptr[1:])
        items = np.fromiter((i for u in user_ids for i in profiles[u]), dtype=np.int32, count=int(rowptr[-1]))
        c_len = next(iter(ctx.values())).shape[0]
        cx = np.zeros((int(rowptr[-1]), c_len), dtype=np.float32)
        j = 0
        for u in user_ids:
            for i in profiles[u]:
                cx[j] = ctx[(u, i)]
                j += 1
        return cls(torch.from_numpy(rowptr), torch.from_numpy(items), torch.from_numpy(cx))

    def lengths(self) -> Tensor:
        return (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)

    def valid_users(self, mode: str) -> Tensor:
        """Rows whose profile is long enough for `mode` (CARCADataset.valid_user_ids, src/data.py:247-248)."""
        return (self.lengths() > {"train": 1, "val": 2, "test": 3}[mode]).nonzero()[:, 0].to(torch.int32)

    def _struct(self) -> N.Interactions:
        s = N.Interactions()
        s.rowptr, s.items, s.ctx = N.i32p(self.rowptr), N.i32p(self.items), N.f32p(self.ctx)
        s.n_users, s.n_ctx = self.n_users, self.ctx.shape[1]
        return s

    def eval_batch(self, users: Tensor, n_items: int, profile_seq_len: int, target_seq_len: int, mode: str,
                   test: bool = True, seed: int = 0):
        """The 7-tuple of get_test_sequences for `users` (rows of the log): o_c is an expanded [B,T,C] view of
        one context row per user; p_a / o_a are None (attributes come from the model's ItemAttrTable)."""
        if mode not in ("val", "test"):
            raise ValueError(f"Invalid mode: {mode}")                     # src/data.py:54-55
        N.require_device(users, self.items)
        users = users.to(torch.int32).contiguous()
        B, L, T, Cn = users.numel(), int(profile_seq_len), int(target_seq_len) + 1, self.ctx.shape[1]
        dev = self.items.device
        p_x = torch.empty((B, L), dtype=torch.int32, device=dev)
        p_c = torch.empty((B, L, Cn), dtype=torch.float32, device=dev)
        o_x = torch.empty((B, T), dtype=torch.int32, device=dev)
        o_cu = torch.empty((B, Cn), dtype=torch.float32, device=dev)
        y = torch.empty((B, T), dtype=torch.int32, device=dev)
        s = self._struct()
        N.call("carca_build_eval_batch", N.i32p(p_x), N.f32p(p_c), N.i32p(o_x), N.f32p(o_cu), N.i32p(y), C.byref(s),
               N.i32p(users), B, L, T, int(n_items), MODES[mode], int(bool(test)), int(seed) & (2 ** 64 - 1), N.stream())
        return p_x, None, p_c, o_x, None, o_cu.unsqueeze(1).expand(B, T, Cn), y

    def train_batch(self, users: Tensor, n_items: int, seq_len: int, test: bool = True, seed: int = 0):
        """The 7-tuple of get_train_sequences for `users`."""
        N.require_device(users, self.items)
        users = users.to(torch.int32).contiguous()
        B, L, Cn = users.numel(), int(seq_len), self.ctx.shape[1]
        dev = self.items.device
        p_x = torch.empty((B, L), dtype=torch.int32, device=dev)
        p_c = torch.empty((B, L, Cn), dtype=torch.float32, device=dev)
        o_x = torch.empty((B, 2 * L), dtype=torch.int32, device=dev)
        o_c = torch.empty((B, 2 * L, Cn), dtype=torch.float32, device=dev)
        y = torch.empty((B, 2 * L), dtype=torch.int32, device=dev)
        s = self._struct()
        N.call("carca_build_train_batch", N.i32p(p_x), N.f32p(p_c), N.i32p(o_x), N.f32p(o_c), N.i32p(y), C.byref(s),
               N.i32p(users), B, L, int(n_items), int(bool(test)), int(seed) & (2 ** 64 - 1), N.stream())
        return p_x, None, p_c, o_x, None, o_c, y


class DeviceLoader:
    """Stands in for DataLoader(CARCADataset(...)) (scripts/training.py:120-163) in train()/evaluate():
    iterates the valid users of `mode` in batches built on the device; `shuffle` permutes them per epoch."""

    def __init__(self, log: DeviceInteractions, n_items: int, profile_seq_len: int, target_seq_len: int, mode: str,
                 batch_size: int, test: bool = True, shuffle: bool = False, seed: int = 0,
                 users: Optional[Tensor] = None):
        if mode not in MODES:
            raise ValueError(f"Invalid mode: {mode}")
        self.log, self.n_items, self.L, self.T = log, int(n_items), int(profile_seq_len), int(target_seq_len)
        self.mode, self.batch_size, self.test, self.shuffle, self.seed = mode, int(batch_size), test, shuffle, int(seed)
        self.users = log.valid_users(mode) if users is None else users.to(torch.int32)
        self.epoch = 0

    def __len__(self) -> int:
        return (self.users.numel() + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[tuple]:
        users = self.users
        if self.shuffle:
            g = torch.Generator(device=users.device).manual_seed(self.seed + self.epoch)
            users = users[torch.randperm(users.numel(), device=users.device, generator=g)]
        for i in range(len(self)):
            u = users[i * self.batch_size:(i + 1) * self.batch_size]
            seed = (self.seed * 1_000_003 + self.epoch) * 1_000_003 + i
            if self.mode == "train":
                yield self.log.train_batch(u, self.n_items, self.L, self.test, seed)
            else:
                yield self.log.eval_batch(u, self.n_items, self.L, self.T, self.mode, self.test, seed)
        self.epoch += 1


# ------------------------------------------------------------------------------- packed host -> device batches
class PackedEvalLayout:
    """Byte layout of one packed evaluate() batch inside a pinned host arena / its device mirror:

        [ offs int32 [B + 1] | o_x int32 [B, T] | o_c_user fp32 [B, C] | rows fp32 [R, 1 + C] ]

    `rows` comes last, so a batch occupies a PREFIX of the arena and one copy of `used_bytes(R)` bytes moves it.  What
    is not sent: the padding of p_x / p_c (Beauty-shaped windows are ~86 % padding), the [B, T, C] copies of the
    positive's context (src/data.py:185) and y_true, which is the constant row [1, 0, ..., 0] (src/data.py:189-190)."""

    def __init__(self, B: int, L: int, T: int, C: int):
        self.B, self.L, self.T, self.C = int(B), int(L), int(T), int(C)
        al = lambda x: (x + 255) // 256 * 256                       # noqa: E731
        self.o_offs = 0
        self.o_ox = al(4 * (self.B + 1))
        self.o_oc = al(self.o_ox + 4 * self.B * self.T)
        self.o_rows = al(self.o_oc + 4 * self.B * self.C)
        self.capacity = self.o_rows + 4 * self.B * self.L * (1 + self.C)

    def used_bytes(self, n_rows: int) -> int:
        return self.o_rows + 4 * int(n_rows) * (1 + self.C)

    def dense_bytes(self) -> int:
        """What the same batch takes as the reference loader's dense tensors (p_x, p_c, o_x, o_c, y_true)."""
        return 4 * self.B * (self.L * (1 + self.C) + self.T * (2 + self.C))


def pack_eval_batch(layout: PackedEvalLayout, arena: Tensor, p_x: Tensor, p_c: Tensor, o_x: Tensor, o_c: Tensor) -> int:
    """Host side (a collate_fn's job): writes the batch into `arena` (uint8, ideally pinned) in PackedEvalLayout and
    returns the bytes to copy.  p_x / p_c must be left-padded windows (what pad_profile builds, src/data.py:53-74);
    o_c is [B, T, C] with the positive's context in every row, or [B, C]."""
    B, L, T, Cn = layout.B, layout.L, layout.T, layout.C
    valid = p_x != 0
    lens = valid.sum(1)
    if not bool((valid == (torch.arange(L).unsqueeze(0) >= (L - lens).unsqueeze(1))).all()):
        raise ValueError("pack_eval_batch: windows must be left-padded (no padding between valid positions)")
    offs = torch.zeros(B + 1, dtype=torch.int32)
    offs[1:] = torch.cumsum(lens, 0).to(torch.int32)
    R = int(offs[-1])
    arena[layout.o_offs:layout.o_offs + 4 * (B + 1)].view(torch.int32).copy_(offs)
    arena[layout.o_ox:layout.o_ox + 4 * B * T].view(torch.int32).view(B, T).copy_(o_x)
    ocu = o_c[:, 0, :] if o_c.dim() == 3 else o_c
    arena[layout.o_oc:layout.o_oc + 4 * B * Cn].view(torch.float32).view(B, Cn).copy_(ocu)
    rows = arena[layout.o_rows:layout.o_rows + 4 * R * (1 + Cn)].view(torch.float32).view(R, 1 + Cn)
    rows[:, 0].copy_(p_x[valid].to(torch.int32).view(torch.float32))
    rows[:, 1:].copy_(p_c[valid])
    return layout.used_bytes(R)


def unpack_eval_batch(layout: PackedEvalLayout, dev_arena: Tensor, out: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """Device side: the evaluate() batch as the tensors CARCA.forward / the metrics take (p_x, p_c rebuilt by one
    kernel; o_x a view of the arena; o_c and y_true expanded views — nothing else is materialised).  `out` (from a
    previous call) reuses the p_x / p_c buffers, so the call can sit inside a CUDA graph."""
    B, L, T, Cn = layout.B, layout.L, layout.T, layout.C
    N.require_device(dev_arena)
    dev = dev_arena.device
    offs = dev_arena[layout.o_offs:layout.o_offs + 4 * (B + 1)].view(torch.int32)
    o_x = dev_arena[layout.o_ox:layout.o_ox + 4 * B * T].view(torch.int32).view(B, T)
    o_cu = dev_arena[layout.o_oc:layout.o_oc + 4 * B * Cn].view(torch.float32).view(B, Cn)
    rows = dev_arena[layout.o_rows:layout.capacity].view(torch.float32)
    if out is None:
        label = torch.zeros((1, T), dtype=torch.int32, device=dev)
        label[0, 0] = 1
        out = {"p_x": torch.empty((B, L), dtype=torch.int32, device=dev),
               "p_c": torch.empty((B, L, Cn), dtype=torch.float32, device=dev),
               "o_x": o_x, "o_c": o_cu.unsqueeze(1).expand(B, T, Cn), "y_true": label.expand(B, T)}
    N.call("carca_unpack_windows", N.i32p(out["p_x"]), N.f32p(out["p_c"]), N.i32p(offs), N.f32p(rows), B, L, Cn, N.stream())
    return out
