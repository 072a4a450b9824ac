"""Shared protocol for every model in the zoo.

The block-coordinate trainers only need three things from a model
(reference protocol: /root/reference/src/simple_models.py:28-39):

* ``train_order_block_ids()``  -> list of ``[low, high]`` inclusive ranges of
  parameter indices (positions in ``net.parameters()``), in training order;
* ``linear_layer_ids()``       -> parameter indices of the dense-layer weights
  (used by the elastic-net gate);
* ``linear_layer_parameters()``-> flat vector of the dense-layer parameters.

Here the tables are class attributes (``BLOCK_TABLE`` / ``LINEAR_IDS`` /
``LINEAR_MODULES``) and the methods are provided once by this mixin.  Because
each block is a *contiguous index range* in registration order, a block is also
a contiguous slice of the model's flat parameter arena
(:mod:`federated_pytorch_test_b200.utils.flat`), which is what makes zero-copy
aggregation possible.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn


class BlockPartitioned(nn.Module):
    BLOCK_TABLE: Sequence[Sequence[int]] = ()
    LINEAR_IDS: Sequence[int] = ()
    LINEAR_MODULES: Sequence[str] = ()

    def train_order_block_ids(self) -> List[List[int]]:
        return [list(b) for b in self.BLOCK_TABLE]

    def linear_layer_ids(self) -> List[int]:
        return list(self.LINEAR_IDS)

    def linear_layer_parameters(self) -> torch.Tensor:
        """All dense-layer parameters as one vector.

        The reference chains generators with ``or`` and therefore only ever
        returns fc1 (simple_models.py:33-35); the method is never
        called there.  We return what the docstring promises.
        """
        chunks = []
        for name in self.LINEAR_MODULES:
            for p in getattr(self, name).parameters():
                chunks.append(p.reshape(-1))
        if not chunks:
            ref = next(self.parameters())
            return ref.new_zeros(0)
        return torch.cat(chunks)

    # -- helpers used by tests and the flat arena -------------------------
    def block_numel(self, blockid: int) -> int:
        lo, hi = self.train_order_block_ids()[blockid]
        return sum(p.numel() for i, p in enumerate(self.parameters()) if lo <= i <= hi)
