"""Whole-train-step CUDA graph (B200: launch-bound inner loops belong in graphs).

The reference's step body (src/train.py:90-96: zero_grad, forward, BCE, backward, Adam) is ~170
kernel launches of this library plus the optimizer's; at the reference batch size (256) the Python /
ctypes / autograd dispatch of those launches takes longer than the kernels run.  `GraphedTrainStep`
captures the whole body once (static input buffers, capturable Adam) and replays it per batch:

    step = GraphedTrainStep(model, optim, batch_like)      # warm-up + capture
    loss = step(batch)                                     # device scalar, no host sync

Dropout masks still change every step: the captured kernels XOR a device-resident seed word into
their Philox seed (carca_set_seed_source) and the graph increments that word on every replay.
Semantics are those of the eager step on the same data (tests/test_gpu_graph.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
from torch import Tensor

from . import _native as N
from .carca import BinaryCrossEntropy
from .utils import get_mask


class GraphedTrainStep:
    KEYS = ("p_x", "p_c", "o_x", "o_c", "y_true")

    def __init__(self, model, optim: torch.optim.Optimizer, batch: Dict[str, Tensor],
                 loss_fn: Optional[BinaryCrossEntropy] = None, warmup: int = 3):
        for g in optim.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedTrainStep needs an optimizer built with capturable=True")
        self.model, self.optim = model, optim
        self.loss_fn = loss_fn or BinaryCrossEntropy()
        dev = batch["p_x"].device
        self.static = {k: batch[k].clone() for k in self.KEYS}
        self.seed = torch.zeros(1, dtype=torch.int64, device=dev)      # device seed word, bumped per replay
        self.graph = torch.cuda.CUDAGraph()
        # windows longer than 64: the fused training kernels need every user's active positions to fit a 64-row bin,
        # which cannot be checked inside a capture — check the example batch here and every batch before its replay
        self.long_windows = batch["p_x"].shape[1] > 64 and hasattr(model, "max_active_positions")
        if self.long_windows:
            model._fits_override = self._fits(batch)
        N.lib().carca_set_seed_source(self.seed.data_ptr())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        optim.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        N.lib().carca_set_seed_source(None)
        self.graph_is_fused = bool(getattr(model, "_fits_override", None))
        if self.long_windows:
            model._fits_override = None

    def _fits(self, batch: Dict[str, Tensor]) -> bool:
        L = batch["p_x"].shape[1]
        return self.model.max_active_positions(batch["p_x"], [batch["o_x"][:, :L], batch["o_x"][:, L:]]) <= 64

    def _body(self) -> Tensor:
        b = self.static
        L = b["p_x"].shape[1]
        self.seed.add_(1)
        self.optim.zero_grad(set_to_none=True)
        y = self.model.forward(profile=(b["p_x"], None, b["p_c"]),
                               targets=[(b["o_x"][:, :L], None, b["o_c"][:, :L]), (b["o_x"][:, L:], None, b["o_c"][:, L:])])
        loss = self.loss_fn.forward(y, b["y_true"], get_mask(b["o_x"]))
        loss.backward()
        self.optim.step()
        return loss.detach()

    def __call__(self, batch: Dict[str, Tensor]) -> Tensor:
        for k in self.KEYS:
            self.static[k].copy_(batch[k], non_blocking=True)
        if self.long_windows and self.graph_is_fused and not self._fits(batch):
            # a user with more than 64 active positions: this step runs eagerly on the per-op kernels
            N.lib().carca_set_seed_source(self.seed.data_ptr())
            try:
                return self._body()
            finally:
                N.lib().carca_set_seed_source(None)
        self.graph.replay()
        return self.loss
