"""Plug-in interfaces of the hot path — the drop-in boundary on the Python side.

Same five abstract `nn.Module` roles, names and `forward` signatures as the reference's
`src/abstract.py:8-50`, so code typed against them (src/train.py:4,35,57;
scripts/training.py:13) accepts the B200 modules unchanged.
"""
from abc import ABC, abstractmethod
from typing import List, Tuple

import torch.nn as nn
from torch import Tensor

Triple = Tuple[Tensor, Tensor, Tensor]


class _Role(nn.Module, ABC):
    """Common base: an abstract torch module (keeps isinstance(nn.Module) and ABC checks)."""

    def __init__(self) -> None:
        super().__init__()


class Model(_Role):
    """(profile, targets) -> probabilities [B, sum(T)]; src/abstract.py:8-14."""

    @abstractmethod
    def forward(self, profile: Triple, targets: List[Triple]) -> Tensor:
        ...


class Embedding(_Role):
    """(x ids, a attributes, c context, mask, target flag) -> [B, N, d]; src/abstract.py:17-23."""

    @abstractmethod
    def forward(self, x: Tensor, a: Tensor, c: Tensor, mask: Tensor, target: bool) -> Tensor:
        ...


class Encoding(_Role):
    """Positional encoding x -> x (+ pos); src/abstract.py:26-32."""

    @abstractmethod
    def forward(self, x: Tensor) -> Tensor:
        ...


class Encoder(_Role):
    """(x, mask) -> x'; src/abstract.py:35-41."""

    @abstractmethod
    def forward(self, x: Tensor, mask: Tensor) -> Tensor:
        ...


class Decoder(_Role):
    """(o, o_mask, p, p_mask) -> scores [B, T]; src/abstract.py:44-50."""

    @abstractmethod
    def forward(self, o: Tensor, o_mask: Tensor, p: Tensor, p_mask: Tensor) -> Tensor:
        ...
