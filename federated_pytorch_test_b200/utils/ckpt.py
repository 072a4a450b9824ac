"""Checkpoints: the reference's per-worker files + a true resume record.

Legacy schema (/root/reference/src/federated_multi.py:226-233,
100-103): ``./s{ck}.model`` = ``{'model_state_dict', 'epoch',
'optimizer_state_dict', 'running_loss'}``, loading restores the model weights
only.  CPC writes ``encoder{ck}.model`` / ``contextgen{ck}.model`` /
``predictor{ck}.model`` with only ``model_state_dict`` and loads the un-suffixed
``./encoder.model`` … into every worker (federated_cpc.py:125-134,308-318).

Files written here are readable by the reference scripts and vice versa: tensors
are detached from the flat arena and made dense before saving.

The resume record (new) additionally stores the schedule position, the
consensus variables (z, y_k, rho, BB state), optimizer state and RNG states.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn


def dense_state_dict(net: nn.Module) -> Dict[str, torch.Tensor]:
    """``state_dict`` with every tensor cloned out of the arena into standalone contiguous storage."""
    return {k: v.detach().clone(memory_format=torch.contiguous_format) for k, v in net.state_dict().items()}


def load_into(net: nn.Module, state: Dict[str, torch.Tensor]) -> None:
    """``load_state_dict`` that keeps arena views intact (copies element-wise into the existing storage)."""
    own = net.state_dict()
    missing = [k for k in own if k not in state]
    unexpected = [k for k in state if k not in own]
    if missing or unexpected:
        raise KeyError("state_dict mismatch: missing=%s unexpected=%s" % (missing, unexpected))
    with torch.no_grad():
        for k, v in own.items():
            v.copy_(state[k].to(v.device))
    if any(v.is_cuda for v in own.values()):
        from ..ops import cuda_ops

        cuda_ops.clear_caches()   # derived tensors (e.g. flipped dgrad filters) of frozen layers are now stale


def worker_path(ckpt_dir: str, ck: int, stem: str = "s") -> str:
    return os.path.join(ckpt_dir, "%s%d.model" % (stem, ck))


def save_worker(ckpt_dir: str, ck: int, net: nn.Module, epoch: int, optimizer, running_loss: float, stem: str = "s") -> str:
    path = worker_path(ckpt_dir, ck, stem)
    torch.save({
        "model_state_dict": dense_state_dict(net),
        "epoch": epoch,
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
        "running_loss": running_loss,
    }, path)
    return path


def load_worker(ckpt_dir: str, ck: int, net: nn.Module, device, stem: str = "s") -> None:
    ckpt = torch.load(worker_path(ckpt_dir, ck, stem), map_location=device, weights_only=False)
    load_into(net, ckpt["model_state_dict"])
    net.train()


def save_model_only(path: str, net: nn.Module) -> str:
    torch.save({"model_state_dict": dense_state_dict(net)}, path)
    return path


def load_model_only(path: str, net: nn.Module, device) -> None:
    ckpt = torch.load(path, map_location=device, weights_only=False)
    load_into(net, ckpt["model_state_dict"])
    net.train()


# ----------------------------------------------------------------------------
def save_resume(path: str, engine, position: Dict) -> str:
    """True resume record: schedule position + consensus state + per-replica weights/optimizers + RNG."""
    strat_state = {}
    for k, v in engine.strategy.state().items():
        if torch.is_tensor(v):
            strat_state[k] = v.detach().cpu().clone()
        elif isinstance(v, list):
            strat_state[k] = [t.detach().cpu().clone() for t in v]
        else:
            strat_state[k] = v
    rec = {
        "position": dict(position),
        "strategy": engine.strategy.name,
        "strategy_state": strat_state,
        "local_workers": list(engine.topo.local_workers),
        "replicas": {
            rep.ck: {key: dense_state_dict(net) for key, net in rep.nets.items()} for rep in engine.replicas
        },
        "optimizers": {
            rep.ck: opt.state_dict() for rep, opt in zip(engine.replicas, engine.optimizers)
        } if engine.optimizers else {},
        "rng": {"torch": torch.get_rng_state(),
                "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else None},
    }
    if engine.topo.is_distributed:
        path = "%s.rank%d" % (path, engine.topo.rank)
    torch.save(rec, path)
    return path


def load_resume(path: str, engine) -> Dict:
    if engine.topo.is_distributed:
        path = "%s.rank%d" % (path, engine.topo.rank)
    rec = torch.load(path, map_location="cpu", weights_only=False)
    for rep in engine.replicas:
        for key, net in rep.nets.items():
            load_into(net, rec["replicas"][rep.ck][key])
    torch.set_rng_state(rec["rng"]["torch"])
    if rec["rng"]["cuda"] is not None and torch.cuda.is_available():
        try:
            torch.cuda.set_rng_state_all(rec["rng"]["cuda"])
        except Exception:
            pass
    return rec
