"""get_mask / to — same helpers as the reference's src/utils.py:6-11."""
from typing import Tuple

import torch

from . import ops


def get_mask(input: torch.Tensor) -> torch.Tensor:
    """fp32 padding mask: 0.0 where the input is 0, else 1.0 (src/utils.py:6-7), computed by the
    carca_padding_mask / carca_padding_mask_f32 kernels.  Device tensors only: like every op of this
    package there is no CPU path (a host tensor raises RuntimeError)."""
    return ops.padding_mask(input)


def to(*tensors: torch.Tensor, device: str) -> Tuple[torch.Tensor, ...]:
    """Move a batch tuple to `device` (src/utils.py:10-11)."""
    return tuple(None if t is None else t.to(device) for t in tensors)
