"""Small CIFAR10 classifiers: ``Net``, ``Net1``, ``Net2``.

Behavioural spec: /root/reference/src/simple_models.py:9-128 (ELU everywhere,
same registration order so parameter indices / block tables carry over, same
attribute names so ``state_dict`` files interoperate).  Each network is built
from a declarative layer plan instead of hand-written forward code.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import BlockPartitioned
from ..ops import functional as FX


class _PlanNet(BlockPartitioned):
    """conv stack -> flatten -> dense stack, driven by two small tables.

    ``CONV_PLAN``: tuples ``(attr, in_ch, out_ch, kernel, padding, pool_after)``
    ``FC_PLAN``  : tuples ``(attr, in_features, out_features)``; ELU is applied
    after every dense layer but the last.
    """

    CONV_PLAN = ()
    FC_PLAN = ()

    def __init__(self):
        super().__init__()
        pools = 0
        for attr, cin, cout, k, pad, pool in self.CONV_PLAN:
            setattr(self, attr, nn.Conv2d(cin, cout, k, padding=pad))
            if pool:
                pools += 1
        self._register_pools(pools)
        for attr, fin, fout in self.FC_PLAN:
            setattr(self, attr, nn.Linear(fin, fout))

    def _register_pools(self, n: int) -> None:
        for i in range(n):
            setattr(self, "pool%d" % (i + 1), nn.MaxPool2d(2, 2))

    def features(self, x: torch.Tensor) -> torch.Tensor:
        for attr, _cin, _cout, _k, _pad, pool in self.CONV_PLAN:
            x = FX.conv_act_pool(x, getattr(self, attr), act=True, pool=bool(pool))     # conv + bias + ELU (+ pool) fused
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.features(x)
        x = x.reshape(x.shape[0], -1)
        last = len(self.FC_PLAN) - 1
        for i, (attr, _fin, _fout) in enumerate(self.FC_PLAN):
            x = FX.linear_act(x, getattr(self, attr), act=(i != last))                 # GEMM + bias + ELU epilogue
        return x


class Net(_PlanNet):
    """LeNet-style default model, 62 006 parameters in 10 tensors."""

    CONV_PLAN = (("conv1", 3, 6, 5, 0, True), ("conv2", 6, 16, 5, 0, True))
    FC_PLAN = (("fc1", 16 * 5 * 5, 120), ("fc2", 120, 84), ("fc3", 84, 10))
    BLOCK_TABLE = ((4, 5), (0, 1), (2, 3), (6, 7), (8, 9))
    LINEAR_IDS = (4, 6, 8)
    LINEAR_MODULES = ("fc1", "fc2", "fc3")

    def _register_pools(self, n: int) -> None:
        # single shared pooling module named ``pool`` (kept for attribute parity)
        self.pool = nn.MaxPool2d(2, 2)

    def __init__(self):
        # registration order must be conv1, conv2, fc1, fc2, fc3
        super().__init__()


class Net1(_PlanNet):
    """Four valid 3x3 convs + two pools, 890 410 parameters in 12 tensors."""

    CONV_PLAN = (
        ("conv1", 3, 32, 3, 0, False),
        ("conv2", 32, 32, 3, 0, True),
        ("conv3", 32, 64, 3, 0, False),
        ("conv4", 64, 64, 3, 0, True),
    )
    FC_PLAN = (("fc1", 64 * 5 * 5, 512), ("fc2", 512, 10))
    BLOCK_TABLE = ((4, 5), (10, 11), (2, 3), (6, 7), (0, 1), (8, 9))
    LINEAR_IDS = (8, 10)
    LINEAR_MODULES = ("fc1", "fc2")


class Net2(_PlanNet):
    """Four padded conv+pool stages and five dense layers, 2 513 418 parameters."""

    CONV_PLAN = (
        ("conv1", 3, 64, 3, 1, True),
        ("conv2", 64, 128, 3, 1, True),
        ("conv3", 128, 256, 3, 1, True),
        ("conv4", 256, 512, 3, 1, True),
    )
    FC_PLAN = (
        ("fc1", 512 * 2 * 2, 128),
        ("fc2", 128, 256),
        ("fc3", 256, 512),
        ("fc4", 512, 1024),
        ("fc5", 1024, 10),
    )
    BLOCK_TABLE = ((14, 15), (4, 5), (2, 3), (8, 9), (16, 17), (12, 13), (6, 7), (0, 1), (10, 11))
    LINEAR_IDS = (12, 14, 16)
    LINEAR_MODULES = ("fc1", "fc2", "fc3", "fc4", "fc5")
