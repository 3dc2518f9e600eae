"""Adam in one kernel launch: `FusedAdam` stands in for the `torch.optim.Adam(model.parameters(),
lr=..., weight_decay=..., betas=...)` of scripts/training.py:174 with the same constructor arguments, the same
update rule (csrc/optim.cuh) and the same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`), so
`src/train.py`'s `optim.zero_grad()` / `optim.step()` and lr schedulers work unchanged.  The step counter lives on
the device, which makes `step()` CUDA-graph capturable (GraphedTrainStep accepts it like a capturable Adam).
`lr` is passed by value: under GraphedTrainStep the captured launch keeps the learning rate of capture time (an lr
scheduler needs a re-capture; documented limitation)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        # capturable=True: the flag GraphedTrainStep checks (the step counter is a device tensor)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            N.require_device(*live)
            arr = (N.AdamTensor * len(live))()
            keep = []           # contiguous copies of strided gradients stay alive until the launch below is queued
            for i, p in enumerate(live):
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous float32")
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros(1, dtype=torch.float32, device=p.device)   # device counter per parameter
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif not (isinstance(st["step"], torch.Tensor) and st["step"].device == p.device
                          and st["step"].dtype == torch.float32):
                    # state loaded from a stock torch.optim.Adam checkpoint: python / CPU step counter
                    st["step"] = torch.full((1,), float(st["step"]), dtype=torch.float32, device=p.device)
                arr[i] = N.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                      st["step"].data_ptr(), p.numel())
            b1, b2 = group["betas"]
            N.call("carca_adam_step", arr, len(live), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                   float(group["weight_decay"]), N.stream())
            del keep
        # parameters were written through raw pointers: tensor versions did not move, so caches derived from the
        # weights (inference plans, folded item tables) are keyed on this epoch
        from . import fused

        fused.bump_weights_epoch()
        return loss
