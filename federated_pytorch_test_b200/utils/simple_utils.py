"""Parameter-subset plumbing with the reference's public names.

Behavioural spec: /root/reference/src/simple_utils.py:9-87.  The
nine functions keep their names, argument order and observable results.  When
the network lives in a :class:`~.flat.FlatArena` (the normal case in this
framework) pack/unpack degenerate to one strided copy of a contiguous slice —
the trainers themselves never call them, they operate on the slice in place.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .flat import arena_of

_XAVIER_TYPES = (nn.Linear, nn.Conv2d)


def init_weights(m: nn.Module) -> None:
    """Xavier-uniform weights and 0.01 biases for *exactly* ``nn.Linear`` / ``nn.Conv2d``.

    Subclasses and other layer types (ConvTranspose2d, BatchNorm) keep their
    PyTorch defaults, as in the reference (exact ``type(m)==`` test).
    The random numbers are drawn into a dense temporary so the values do not
    depend on the memory format of the destination (arena views may be
    channels-last), i.e. seed-for-seed identical to the reference.
    """
    if type(m) not in _XAVIER_TYPES:
        return
    with torch.no_grad():
        tmp = torch.empty(m.weight.shape, dtype=m.weight.dtype, device=m.weight.device)
        nn.init.xavier_uniform_(tmp)
        m.weight.copy_(tmp)
        if getattr(m, "bias", None) is not None:
            m.bias.fill_(0.01)


def _set_flags(net: nn.Module, predicate) -> None:
    for idx, p in enumerate(net.parameters()):
        p.requires_grad = bool(predicate(idx))
    arena = arena_of(net)
    if arena is not None:
        arena.attach_grads()


def unfreeze_one_layer(net: nn.Module, layer_id: int) -> None:
    """Train only tensors ``2*layer_id`` and ``2*layer_id+1`` (weight+bias pair)."""
    _set_flags(net, lambda i: i in (2 * layer_id, 2 * layer_id + 1))


def unfreeze_all_layers(net: nn.Module) -> None:
    _set_flags(net, lambda i: True)


def freeze_all_layers(net: nn.Module) -> None:
    _set_flags(net, lambda i: False)


def unfreeze_one_block(net: nn.Module, blockid: int) -> None:
    """Train only the tensors whose index lies in ``train_order_block_ids()[blockid]``."""
    lo, hi = net.train_order_block_ids()[blockid]
    _set_flags(net, lambda i: lo <= i <= hi)


def _trainable(net: nn.Module) -> List[torch.Tensor]:
    return [p for p in net.parameters() if p.requires_grad]


def get_trainable_values(net: nn.Module, mydevice: Optional[torch.device] = None) -> torch.Tensor:
    """Trainable parameters packed into a new fp32 vector (registration order)."""
    arena = arena_of(net)
    rng = arena.trainable_range() if arena is not None else None
    if rng is not None:
        out = arena.compact(*rng)
        return out.to(mydevice) if mydevice is not None else out
    plist = _trainable(net)
    n = sum(p.numel() for p in plist)
    dev = mydevice if mydevice is not None else (plist[0].device if plist else torch.device("cpu"))
    out = torch.zeros(n, dtype=torch.float32, device=dev)
    pos = 0
    with torch.no_grad():
        for p in plist:
            k = p.numel()
            out[pos: pos + k].copy_(p.detach().reshape(-1))
            pos += k
    return out


def put_trainable_values(net: nn.Module, X: torch.Tensor) -> None:
    """Scatter vector ``X`` back into the trainable parameters."""
    pos = 0
    with torch.no_grad():
        for p in _trainable(net):
            k = p.numel()
            p.copy_(X[pos: pos + k].view(p.shape))
            pos += k


def number_of_layers(net: nn.Module) -> int:
    """Number of parameter *tensors* (not divided by two, as in the reference)."""
    return sum(1 for _ in net.parameters())


def number_of_blocks(net: nn.Module) -> int:
    return len(net.train_order_block_ids())
