"""``FusedCollective`` — the block collectives as ONE sm_100a kernel each, reducing
straight out of peer memory over NVLink (``csrc/comm_kernels.cu``).  No NCCL on this path.

Reference loops being replaced (Python loops over a dict of K modules): FedAvg mean + write-back + dual residual
``src/federated_multi.py:204-214``, FedProx mean + primal/dual residuals ``src/fedprox_multi.py:205-234``, ADMM z/y
updates + residuals ``src/consensus_multi.py:226-299`` (Barzilai-Borwein dots ``:253-273`` go through
``flatops.multi_dot``).

Memory model: every replica's flat parameter arena (and, for
ADMM, a same-shaped arena for the duals ``y``) is allocated from a
:class:`SymmetricHeap`:

* ``world == 1``  — ordinary device memory (all K replicas are co-resident; the kernel
  just gets K local pointers — the reference's topology, but one launch instead of
  K+2 ATen kernels and K·#tensors copies);
* ``world > 1``   — ``torch.distributed._symmetric_memory`` (CUDA VMM allocations
  mapped into every process, bound to an NVSwitch multicast object when the fabric
  supports it); if that is unavailable, plain allocations exported with CUDA IPC.
  ``torch.distributed`` is used for the handle exchange only.

A block is a slice at the same offset of every arena, so the kernel needs nothing but
K base pointers + one offset.  Cross-rank synchronisation is a flag barrier in a
peer-mapped control pad (release/acquire at system scope, epoch counted in device
memory), i.e. an aggregation round is host-free: one cooperative launch.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import cuda_ops
from .collective import TorchCollective
from .topology import Topology

_MAX_LOCAL = 16


class SymmetricHeap:
    """Allocates fp32/int32 buffers addressable by every rank; remembers peer base pointers."""

    def __init__(self, topo: Topology):
        self.topo = topo
        self.allocs: List[Dict] = []     # {tensor, base, nbytes, peer_ptrs[world], mc_ptr}
        self.transport = "local"
        self._symm = None
        if topo.is_distributed:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self._symm = symm_mem
                self.transport = "symm_mem"
            except Exception:  # pragma: no cover
                self.transport = "ipc"

    # -- allocation ---------------------------------------------------------
    def alloc(self, numel: int, dtype=torch.float32) -> torch.Tensor:
        dev = self.topo.device
        if not self.topo.is_distributed:
            t = torch.zeros(numel, dtype=dtype, device=dev)
            self._record(t, [t.data_ptr()], 0)
            return t
        if self.transport == "symm_mem":
            try:
                t = self._symm.empty(numel, dtype=dtype, device=dev)
                hdl = self._symm.rendezvous(t, group=dist.group.WORLD)
                t.zero_()
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                self._record(t, ptrs, mc, handle=hdl)
                return t
            except Exception as exc:  # fall back once, loudly
                if self.allocs:
                    raise
                print("[fedb200] torch symmetric memory unavailable (%s); using CUDA IPC" % (exc,), flush=True)
                self.transport = "ipc"
        return self._alloc_ipc(numel, dtype)

    def _alloc_ipc(self, numel: int, dtype) -> torch.Tensor:
        e = cuda_ops.ext()
        t = torch.zeros(numel, dtype=dtype, device=self.topo.device)
        torch.cuda.synchronize()
        share = t.untyped_storage()._share_cuda_()
        handle, offset = share[1], share[3]
        gathered: List = [None] * self.topo.world_size
        dist.all_gather_object(gathered, (bytes(handle), int(offset)))
        ptrs = []
        for r, (h, off) in enumerate(gathered):
            if r == self.topo.rank:
                ptrs.append(t.data_ptr())
            else:
                ptrs.append(int(e.ipc_open_handle(h)) + off)
        self._record(t, ptrs, 0, keep=share)
        return t

    def _record(self, t: torch.Tensor, ptrs: List[int], mc: int, **keep) -> None:
        self.allocs.append(dict(tensor=t, base=t.data_ptr(), nbytes=t.numel() * t.element_size(), peer_ptrs=ptrs, mc_ptr=mc, **keep))

    # -- lookup -----------------------------------------------------------------
    def locate(self, t: torch.Tensor) -> Tuple[Dict, int]:
        p = t.data_ptr()
        for a in self.allocs:
            if a["base"] <= p < a["base"] + a["nbytes"]:
                return a, p - a["base"]
        raise KeyError("tensor does not live in the symmetric heap")


class FusedCollective(TorchCollective):
    name = "fused"
    fused = True

    def __init__(self, topo: Topology):
        super().__init__(topo)
        if topo.device.type != "cuda":
            raise RuntimeError("FusedCollective needs a CUDA device")
        if topo.is_distributed and topo.K % topo.world_size != 0:
            raise RuntimeError("fused collectives need K to be a multiple of the number of ranks")
        self.ext = cuda_ops.ext()
        self.heap = SymmetricHeap(topo)
        dev = topo.device
        self.out = torch.zeros(4 + _MAX_LOCAL, dtype=torch.float32, device=dev)
        self.sync = torch.zeros(4, dtype=torch.int32, device=dev)
        self.ctrl = self.heap.alloc(4 * 16, dtype=torch.int32)
        self.ctrl_ptrs = list(self.heap.allocs[-1]["peer_ptrs"])
        self._aux: Dict[Tuple[int, str], torch.Tensor] = {}
        self.use_multimem = True
        self.last_nonfinite = 0.0

    # -- arena hooks ------------------------------------------------------------
    def arena_allocator(self) -> Callable:
        def alloc(numel: int, device) -> torch.Tensor:
            return self.heap.alloc(numel)
        return alloc

    def register_arena(self, arena) -> None:
        self.heap.locate(arena.data)  # raises if the arena was not allocated from the heap

    def zeros_like_block(self, x: torch.Tensor, tag: str) -> torch.Tensor:
        """A zeroed buffer that mirrors block slice ``x`` (same offset in a same-sized symmetric arena)."""
        a, off = self.heap.locate(x)
        key = (a["base"], tag)
        buf = self._aux.get(key)
        if buf is None:
            buf = self.heap.alloc(a["nbytes"] // 4)
            self._aux[key] = buf
        sl = buf[off // 4: off // 4 + x.numel()]
        sl.zero_()
        return sl

    # -- pointer tables -------------------------------------------------------------
    def _tables(self, slices: List[torch.Tensor]):
        """Pointers of ALL K workers' slices (worker ck = rank + j*world for local replica j), local indices, multicast."""
        W, rank, K = self.topo.world_size, self.topo.rank, self.topo.K
        ptrs = [0] * K
        mc = 0
        for j, t in enumerate(slices):
            a, off = self.heap.locate(t)
            for r in range(W):
                ptrs[r + j * W] = a["peer_ptrs"][r] + off
            if len(slices) == 1 and a["mc_ptr"] and self.use_multimem and W > 1:
                mc = a["mc_ptr"] + off
        local_idx = [rank + j * W for j in range(len(slices))]
        return ptrs, local_idx, mc

    def _launch(self, mode: int, xs, ys, z, inv_scale: float, rho: float) -> torch.Tensor:
        n = xs[0].numel()
        if any(t.numel() != n for t in xs) or z.numel() != n:
            raise ValueError("block slices must have equal length")
        xp, local_idx, mcx = self._tables(xs)
        yp, mcy = [], 0
        if ys is not None:
            yp, _, mcy = self._tables(ys)
        if mode == 2 and not (mcx and mcy):
            mcx = mcy = 0
        self.ext.block_reduce(mode, xp, yp, local_idx, z, n, inv_scale, rho, self.out, self.ctrl_ptrs, self.sync,
                              self.topo.world_size, self.topo.rank, mcx, mcy)
        self.launches += 1
        return self.out

    # -- operators ----------------------------------------------------------------------
    @torch.no_grad()
    def fedavg_(self, xs, z, write_back: bool = True):
        out = self._launch(0 if write_back else 1, xs, None, z, 1.0 / self.topo.K, 0.0)
        return out[0].clone()

    @torch.no_grad()
    def fedprox_(self, xs, z, rho: float):
        out = self._launch(1, xs, None, z, 1.0 / self.topo.K, rho)
        res = out[:2].clone()
        return res[0], res[1]

    @torch.no_grad()
    def admm_(self, xs, ys, z, rho: float):
        out = self._launch(2, xs, ys, z, 1.0 / (self.topo.K * rho), rho)
        res = out[:2].clone()
        return res[0], res[1]
