"""Whole-train-step and whole-eval-step CUDA graphs (B200: launch-bound inner loops belong in graphs).

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
        from . import fused

        for k in self.KEYS:
            self.static[k].copy_(batch[k], non_blocking=True)
        fused.bump_weights_epoch()      # the replayed Adam launch writes parameters through raw pointers
        if self.long_windows and self.graph_is_fused and not self._fits(batch):
            # a user with more than 64 active positions: this step runs eagerly on the per-op kernels
            N.lib().carca_set_seed_source(self.seed.data_ptr())
            try:
                return self._body()
            finally:
                N.lib().carca_set_seed_source(None)
        self.graph.replay()
        return self.loss


class GraphedEvalStep:
    """The per-batch body of `evaluate` (src/train.py:44-50: forward, BCE loss, HR@k / NDCG@k accumulation)
    captured once as a CUDA graph and replayed per batch — the eager body is ~8 launches whose Python / ctypes
    dispatch (~0.2 ms) is as long as the kernels run for a Beauty-sized batch.

        stats = torch.zeros(4, dtype=torch.float64, device=dev)   # hits@k, sum 1/log2(rank+2), users, sum of batch losses
        step = GraphedEvalStep(model, first_batch, k=10, stats=stats)
        for batch in loader: step(batch)                          # copies the batch into the static buffers, replays
        hr, ndcg, loss = stats[0] / stats[2], stats[1] / stats[2], stats[3] / n_batches     # one read at the end

    `batch`: dict of p_x, p_c, o_x, o_c, y_true on the device.  With `static_inputs=True` the tensors of `batch` ARE
    the graph's input buffers (e.g. views into a device arena that the caller refills with one H2D copy per step) and
    `replay()` runs the step on whatever they hold.  `result`: optional pinned host fp64[4]; the graph then ends with
    an asynchronous device-to-host copy of `stats` into it (read it after synchronising on an event recorded behind
    the replay).  Several steps (one per input slot) may share one `stats` tensor.
    """
    KEYS = ("p_x", "p_c", "o_x", "o_c", "y_true")

    def __init__(self, model, batch: Dict[str, Tensor], k: int = 10, stats: Optional[Tensor] = None,
                 result: Optional[Tensor] = None, static_inputs: bool = False,
                 loss_fn: Optional[BinaryCrossEntropy] = None, warmup: int = 2, prologue=None):
        from . import ops

        if model.training:
            raise ValueError("GraphedEvalStep captures the eval-mode forward: call model.eval() first")
        dev = batch["p_x"].device
        self.model, self.k = model, int(k)
        self.loss_fn = loss_fn or BinaryCrossEntropy()
        self.static = {key: (batch[key] if static_inputs else batch[key].clone()) for key in self.KEYS}
        self.stats = stats if stats is not None else torch.zeros(4, dtype=torch.float64, device=dev)
        if self.stats.dtype != torch.float64 or self.stats.numel() != 4 or not self.stats.is_contiguous():
            raise ValueError("stats must be a contiguous float64[4] device tensor")
        if result is not None and not (result.is_pinned() and result.dtype == torch.float64 and result.numel() == 4):
            raise ValueError("result must be a pinned float64[4] host tensor")
        self.result = result
        self.prologue = prologue          # optional callable run at the start of every step, inside the graph (e.g.
                                          # device_data.unpack_eval_batch rebuilding the windows from a packed arena)
        self._metrics = ops.eval_metrics_
        keep = self.stats.clone()
        # (windows longer than one 64-row bin run on the packed-rows pipeline, which has no per-user row limit and
        # needs no host-side check: nothing to decide per batch)
        from . import fused

        self.long_windows = False
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                  # builds the inference plan and every cached buffer
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._body()
        self.graph_is_fused = True
        self.uses_plan = fused._plans.get(model) is not None       # the capture reads the fused inference plan
        self.struct_epoch = fused._struct_epoch[0]                 # a later Module._apply (.to / .cuda / .float) makes
                                                                   # the captured plan buffers stale
        self.stats.copy_(keep)                       # warm-up runs do not count

    def _body(self) -> None:
        b = self.static
        if self.prologue is not None:
            self.prologue()
        y = self.model.forward(profile=(b["p_x"], None, b["p_c"]), targets=[(b["o_x"], None, b["o_c"])])
        if type(self.loss_fn) is BinaryCrossEntropy:       # loss + HR@k + NDCG@k in one launch
            self._metrics(self.stats, y, b["y_true"], b["o_x"], self.k)
        else:
            from . import ops

            self.stats[3:4].add_(self.loss_fn.forward(y, b["y_true"], get_mask(b["o_x"])))
            ops.rank_metrics_(self.stats[:3], y, b["y_true"], self.k)
        if self.result is not None:
            self.result.copy_(self.stats, non_blocking=True)

    def refresh(self) -> None:
        """Rebuilds the inference plan (folded tables, packed weights) the captured kernels read, in place, from the
        model's current weights.  `replay()` calls it automatically when the weights changed since the last build
        (optimizer steps, FusedAdam / GraphedTrainStep included, load_state_dict); moving the model to another device
        or dtype invalidates the capture itself and raises."""
        from . import fused

        if not self.uses_plan:
            return
        before = fused._plans.get(self.model)
        bufs = None if before is None else (before.plan.data_ptr(), None if before.rows is None else before.rows.data_ptr())
        emb = self.model.embeds
        n_ctx = self.static["p_c"].shape[-1]
        ent = fused.eval_plan(self.model, emb.attr_table, n_ctx)
        if bufs is not None and bufs[1] is not None:
            ent = fused.rows_plan(self.model, emb.attr_table, n_ctx)
        if bufs is not None and (ent.plan.data_ptr() != bufs[0] or
                                 (bufs[1] is not None and ent.rows.data_ptr() != bufs[1])):
            raise RuntimeError("GraphedEvalStep: the model's parameters were replaced (moved / cast) after capture; "
                               "build a new GraphedEvalStep")

    def replay(self) -> None:
        """Runs the step on the current contents of the static input buffers."""
        from . import fused

        if self.uses_plan and self.struct_epoch != fused._struct_epoch[0]:
            # (the plan may even look current again — rebuilt by an eager call into NEW buffers — while this graph
            # still reads the old ones)
            raise RuntimeError("GraphedEvalStep: a module was moved or cast after capture (Module.to / .cuda / .float); "
                               "build a new GraphedEvalStep")
        if self.uses_plan and not fused.plan_is_current(self.model):
            self.refresh()
        self.graph.replay()

    def __call__(self, batch: Dict[str, Tensor]) -> None:
        for key in self.KEYS:
            if batch[key] is not self.static[key]:
                self.static[key].copy_(batch[key], non_blocking=True)
        self.replay()
