"""get_mask / to — same helpers as the reference's src/utils.py:6-11."""
from typing import Tuple

import torch

from . import ops


def get_mask(input: torch.Tensor) -> torch.Tensor:
    """fp32 padding mask: 0.0 where the id is 0, else 1.0 (src/utils.py:6-7).

    Integer ids on a CUDA device go through the carca_padding_mask kernel; anything else
    (float inputs, the loss mask computed on host tensors in user code) keeps torch semantics.
    """
    if input.is_cuda and input.dtype in (torch.int32, torch.int64):
        return ops.padding_mask(input)
    return torch.where(input == 0.0, 0.0, 1.0)


def to(*tensors: torch.Tensor, device: str) -> Tuple[torch.Tensor, ...]:
    """Move a batch tuple to `device` (src/utils.py:10-11)."""
    return tuple(None if t is None else t.to(device) for t in tensors)
