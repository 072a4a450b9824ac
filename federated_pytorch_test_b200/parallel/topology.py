"""Worker <-> process <-> GPU mapping.

The reference simulates its K workers as K modules in one process on one device,
visited sequentially (/root/reference/src/federated_multi.py:45-48,168).  Here a
*topology* says which of the K logical workers live in this OS process:

* ``Topology.single_process(K, device)`` — the reference's arrangement (all K
  replicas co-resident); used for CPU plumbing runs, parity tests and K > #GPUs;
* ``Topology.from_env(K)`` — one process per GPU under ``torchrun``; worker ``k``
  is owned by rank ``k % world_size`` (1:1 when K == world_size; several
  co-resident replicas per GPU when K > world_size).

``torch.distributed`` (NCCL on GPUs, Gloo on CPU) is used for bootstrap, for
exchanging symmetric-memory handles and for the *baseline* collectives; the
product collectives are the kernels in ``csrc/comm_kernels.cu``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class Topology:
    K: int
    world_size: int = 1
    rank: int = 0
    device: torch.device = field(default_factory=lambda: torch.device("cpu"))
    group: Optional[object] = None

    @property
    def local_workers(self) -> List[int]:
        return [k for k in range(self.K) if k % self.world_size == self.rank]

    def owner(self, ck: int) -> int:
        return ck % self.world_size

    @property
    def is_distributed(self) -> bool:
        return self.world_size > 1

    @property
    def is_root(self) -> bool:
        return self.rank == 0

    # ------------------------------------------------------------------
    @staticmethod
    def single_process(K: int, device=None) -> "Topology":
        if device is None:
            device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        return Topology(K=K, world_size=1, rank=0, device=torch.device(device))

    @staticmethod
    def from_env(K: Optional[int] = None, use_cuda: bool = True, backend: Optional[str] = None) -> "Topology":
        """Build from ``RANK/WORLD_SIZE/LOCAL_RANK`` (torchrun).  Falls back to single process."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            dev = torch.device("cuda", 0) if (use_cuda and torch.cuda.is_available()) else torch.device("cpu")
            return Topology.single_process(K if K is not None else 1, dev)
        rank = int(os.environ["RANK"])
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        cuda = use_cuda and torch.cuda.is_available()
        if cuda:
            torch.cuda.set_device(local_rank)
            device = torch.device("cuda", local_rank)
        else:
            device = torch.device("cpu")
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {}
            if cuda:
                kw["device_id"] = device
            dist.init_process_group(backend or ("nccl" if cuda else "gloo"), rank=rank, world_size=world, **kw)
        return Topology(K=K if K is not None else world, world_size=world, rank=rank, device=device, group=dist.group.WORLD)

    def barrier(self) -> None:
        if self.is_distributed:
            if self.device.type == "cuda":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()
