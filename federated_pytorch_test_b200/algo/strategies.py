"""Aggregation strategies over one parameter block: FedAvg, FedProx, consensus
ADMM (with optional Barzilai-Borwein / spectral adaptive rho).

Mathematical spec: SURVEY §2.4; sources /root/reference/src/federated_multi.py:
203-217, fedprox_multi.py:183-234, consensus_multi.py:152-299.  In the reference
these are copy-pasted inline loops over Python dicts; here each is a small
object driven by the block-coordinate engine:

    strat.begin_block(ci, N, xs)      # xs: flat slices of the local replicas
    strat.penalty(i)                  # what the local optimizer must add to the loss
    strat.aggregate(nadmm)            # the collective + bookkeeping -> metrics

All vector work is delegated to a collective (``parallel.collective``), i.e. to
one fused NVLink kernel per aggregation on B200.

Preserved reference behaviour: ``z`` (and ``y``) start at 0 for
every block visit (Q6); FedProx/ADMM never write ``z`` back (Q7); ``rho`` is an
``[L,3]`` table of which only column 0 is used, shared by all workers and
updated sequentially over workers inside the BB step (Q8); the BB state
``yhat0`` is seeded with the parameter values (Q9).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass
class Penalty:
    """Terms the local objective gets on top of the data loss, on the block vector ``x``:
    ``y.(x-z) + rho/2 ||x-z||^2``."""

    z: Optional[torch.Tensor] = None
    y: Optional[torch.Tensor] = None
    rho: float = 0.0


class Strategy:
    name = "base"
    write_back = False

    def __init__(self, collective, topo):
        self.coll = collective
        self.topo = topo
        self.xs: List[torch.Tensor] = []
        self.z: Optional[torch.Tensor] = None
        self.N = 0
        self.ci = 0

    def begin_block(self, ci: int, N: int, xs: List[torch.Tensor]) -> None:
        self.ci, self.N, self.xs = ci, int(N), xs
        self.z = torch.zeros_like(xs[0])  # Q6: restart from the origin on every block visit

    def penalty(self, i: int) -> Penalty:
        return Penalty()

    def aggregate(self, nadmm: int) -> Dict[str, float]:
        raise NotImplementedError

    def rho_mean(self) -> float:
        return float("nan")

    # extra vectors a true-resume checkpoint must carry
    def state(self) -> Dict[str, object]:
        return {"z": self.z}


class NoConsensus(Strategy):
    """Stand-alone training: nothing is exchanged (no_consensus_multi.py)."""

    name = "none"

    def aggregate(self, nadmm: int) -> Dict[str, float]:
        return {}


class FedAvg(Strategy):
    name = "fedavg"
    write_back = True

    def aggregate(self, nadmm: int) -> Dict[str, float]:
        dual_sq = self.coll.fedavg_(self.xs, self.z, write_back=True)
        return {"dual": math.sqrt(max(float(dual_sq), 0.0)) / self.N}


class FedProx(Strategy):
    name = "fedprox"

    def __init__(self, collective, topo, num_blocks: int, rho0: float = 1.0):
        super().__init__(collective, topo)
        self.rho = torch.ones(num_blocks, 3) * rho0

    def penalty(self, i: int) -> Penalty:
        return Penalty(z=self.z, y=None, rho=float(self.rho[self.ci, 0]))

    def rho_mean(self) -> float:
        return float(self.rho.mean())

    def aggregate(self, nadmm: int) -> Dict[str, float]:
        rho = float(self.rho[self.ci, 0])
        dual_sq, primal = self.coll.fedprox_(self.xs, self.z, rho)
        return {"dual": math.sqrt(max(float(dual_sq), 0.0)) / self.N, "primal": float(primal) / self.N}


@dataclass
class BBConfig:
    enabled: bool = False
    period_T: int = 2
    alphacorrmin: float = 0.2
    epsilon: float = 1e-3
    rhomax: float = 0.1
    seed_yhat0_with_x: bool = True   # Q9 (reference behaviour); False seeds with zeros


class ADMM(Strategy):
    name = "admm"

    def __init__(self, collective, topo, num_blocks: int, rho0: float = 0.1, bb: Optional[BBConfig] = None, log=print):
        super().__init__(collective, topo)
        self.rho = torch.ones(num_blocks, 3) * rho0
        self.bb = bb or BBConfig()
        self.ys: List[torch.Tensor] = []
        self.yhat0: List[torch.Tensor] = []
        self.x0: List[torch.Tensor] = []
        self.log = log

    def begin_block(self, ci: int, N: int, xs: List[torch.Tensor]) -> None:
        super().begin_block(ci, N, xs)
        self.ys = [self.coll.zeros_like_block(x, "y") for x in xs]
        if self.bb.enabled:
            self.yhat0 = [x.clone() if self.bb.seed_yhat0_with_x else torch.zeros_like(x) for x in xs]
            self.x0 = [torch.zeros_like(x) for x in xs]

    def penalty(self, i: int) -> Penalty:
        return Penalty(z=self.z, y=self.ys[i], rho=float(self.rho[self.ci, 0]))

    def rho_mean(self) -> float:
        return float(self.rho.mean())

    def state(self) -> Dict[str, object]:
        return {"z": self.z, "y": self.ys, "rho": self.rho, "yhat0": self.yhat0, "x0": self.x0}

    # -- adaptive rho ---------------------------------------------------------
    def _bb_update(self, nadmm: int) -> None:
        """Spectral penalty selection, replayed identically on every rank.

        The reference loops over workers, each one reading and possibly
        overwriting the shared ``rho[ci,0]`` (consensus_multi.py:248-278).  With
        ``a=y-yhat0, b=x-z, c=x-x0`` the quantities it needs are
        ``d11 = a.a + 2 rho a.b + rho^2 b.b``, ``d12 = a.c + rho b.c``, ``d22 = c.c``,
        so six local dots per worker + one tiny gather reproduce the sequential
        rule without serialising the GPUs.
        """
        cfg = self.bb
        rows = self.coll.bb_dots(self.xs, self.ys, self.yhat0, self.x0, self.z).double().cpu()
        rho = float(self.rho[self.ci, 0])
        rho_at_turn = []
        for ck in range(self.topo.K):
            aa, ab, bb_, ac, bc, cc = (float(v) for v in rows[ck])
            rho_at_turn.append(rho)
            d11 = aa + 2.0 * rho * ab + rho * rho * bb_
            d12 = ac + rho * bc
            d22 = cc
            self.log("admm %d deltas=(%e,%e,%e)" % (nadmm, d11, d12, d22))
            rhonew = rho
            if abs(d12) > cfg.epsilon and d11 > cfg.epsilon and d22 > cfg.epsilon:
                alpha = d12 / math.sqrt(d11 * d22)
                alphaSD = d11 / d22
                alphaMG = d12 / d22
                alphahat = alphaMG if 2.0 * alphaMG > alphaSD else alphaSD - 0.5 * alphaMG
                if alpha >= cfg.alphacorrmin and alphahat < cfg.rhomax:
                    rhonew = alphahat
                self.log("admm %d alphas=(%e,%e,%e)" % (nadmm, alpha, alphaSD, alphaMG))
            rho = rhonew
        self.rho[self.ci, 0] = rho
        # carry forward: yhat0_k <- y_k + rho_k (x_k - z) with the rho in force at worker k's turn
        for i, ck in enumerate(self.topo.local_workers):
            r = rho_at_turn[ck]
            torch.add(self.ys[i], self.xs[i] - self.z, alpha=r, out=self.yhat0[i])
            self.x0[i].copy_(self.xs[i])

    def aggregate(self, nadmm: int) -> Dict[str, float]:
        if self.bb.enabled:
            if nadmm == 0:
                for i in range(len(self.xs)):
                    self.x0[i].copy_(self.xs[i])
            elif nadmm % self.bb.period_T == 0:
                self._bb_update(nadmm)
        rho = float(self.rho[self.ci, 0])
        dual_sq, primal = self.coll.admm_(self.xs, self.ys, self.z, rho)
        return {"dual": math.sqrt(max(float(dual_sq), 0.0)) / self.N, "primal": float(primal) / self.N}
