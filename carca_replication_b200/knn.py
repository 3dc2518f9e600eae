"""Attribute nearest-neighbour baseline behind the Model API (src/knn.py:8-21)."""
from typing import List, Tuple

import torch
from torch import Tensor

from . import ops
from .abstract import Model


class KNN(Model):
    """y = <attributes of the last profile item, attributes of each candidate>, on the dense [B, N, A]
    attribute tensors of the reference API (no parameters, no sigmoid)."""

    def __init__(self):
        super().__init__()

    def forward(self, profile: Tuple[Tensor, Tensor, Tensor], targets: List[Tuple[Tensor, Tensor, Tensor]]) -> Tensor:
        p_x, p_a, p_c = profile
        ys = [ops.knn_scores(p_a, o_a) for o_x, o_a, o_c in targets]
        return ys[0] if len(ys) == 1 else torch.cat(ys, dim=-1)
