"""Block-coordinate federated training engine.

One engine replaces the seven copy-pasted driver loops of the reference
(/root/reference/src/federated_multi.py:143-220 and siblings):

    for nloop:  for block visit:  unfreeze block in every replica; z=0; new optimizer state
        for nadmm:  for epoch:
            for local replica:  for minibatch:  opt.step(closure); diagnostics loss
            strategy.aggregate()          <- ONE fused collective on the block slice
            (optional) evaluate all replicas on the test set

What differs from the reference by construction:

* replicas live where the :class:`~..parallel.topology.Topology` puts them (one
  per GPU under torchrun); only *local* replicas are stepped, all ranks run the
  same schedule and meet in the collective;
* a block is a zero-copy slice of each replica's flat arena — no pack/unpack, no
  ``torch.cat`` inside closures; penalty gradients are closed-form inside the
  optimizer kernel;
* the per-minibatch step of the Adam drivers can be captured once per
  (replica, block) as a CUDA graph and replayed (``graphs=True``), which removes
  the ~400 eager launches and the per-step host sync of the reference;
* the diagnostics loss is accumulated on the device; it is read back once per
  round (or per minibatch only when ``be_verbose``).

A *task* (``api/*.py``) supplies models, data, loss, schedule and evaluation.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ..ops import flatops
from ..optim.block_adam import BlockAdam
from ..optim.lbfgsnew import LBFGSNew
from ..parallel.topology import Topology
from ..utils.flat import FlatArena
from ..utils.metrics import MetricsLog, PhaseTimers
from .strategies import Penalty, Strategy


@dataclass
class Visit:
    """One entry of the block schedule."""

    model: str                 # key of the sub-model being trained ('net', 'encoder', ...)
    lo: int                    # first trainable parameter index (inclusive)
    hi: int                    # last trainable parameter index (inclusive)
    ci: int                    # block index: selects rho[ci], the elastic-net gate, log labels
    label: Tuple[int, int]     # what the legacy log lines print as block=[a,b]
    optimizer: str = "adam"    # 'adam' | 'lbfgs'
    opt_kwargs: Dict = field(default_factory=dict)
    lambda1: float = 0.0       # elastic net on the block vector (already gated by the task)
    lambda2: float = 0.0
    tag: Dict = field(default_factory=dict)


class Replica:
    """One logical worker: its models, arenas and data."""

    def __init__(self, ck: int, nets: Dict[str, nn.Module], device: torch.device, allocator=None,
                 channels_last_weights: bool = False):
        self.ck = ck
        self.nets = nets
        self.device = device
        self.arenas: Dict[str, FlatArena] = {}
        for key, net in nets.items():
            net.to(device)
            self.arenas[key] = FlatArena(net, device=device, allocator=allocator,
                                         channels_last_weights=channels_last_weights)
        self.running_loss = 0.0
        self.extra: Dict = {}

    def set_trainable(self, visit: Visit) -> None:
        for key, net in self.nets.items():
            active = key == visit.model
            for idx, p in enumerate(net.parameters()):
                p.requires_grad = active and visit.lo <= idx <= visit.hi
            self.arenas[key].attach_grads()

    def block(self, visit: Visit) -> torch.Tensor:
        return self.arenas[visit.model].block(visit.lo, visit.hi)

    def block_grad(self, visit: Visit) -> torch.Tensor:
        return self.arenas[visit.model].block_grad(visit.lo, visit.hi)


class Task:
    """What a driver must provide.  See ``api/`` for the seven concrete tasks."""

    def build_replica(self, ck: int, device: torch.device, allocator) -> Replica:
        raise NotImplementedError

    def visits(self, nloop: int) -> Iterable[Visit]:
        raise NotImplementedError

    def batches(self, rep: Replica, visit: Visit, epoch: int) -> Iterator:
        raise NotImplementedError

    def loss(self, rep: Replica, batch) -> torch.Tensor:
        raise NotImplementedError

    def batch_size_of(self, batch) -> int:
        return int(batch[0].shape[0])

    def evaluate(self, reps: List[Replica], engine: "Engine") -> Optional[List[float]]:
        return None

    def on_epoch_start(self, epoch: int, engine: "Engine") -> None:
        """Hook at the top of every epoch (``no_consensus`` prints ``Epoch %d``)."""

    def after_minibatch(self, rep: Replica, visit: Visit, batch, i: int, epoch: int, nloop: int, N: int,
                        loss1: torch.Tensor, engine: "Engine") -> None:
        """Per-minibatch logging hook (verbose drivers)."""

    def aggregate_log(self, visit: Visit, metrics: Dict[str, float], ctx: Dict, engine: "Engine") -> None:
        """Print the legacy per-round line."""


@dataclass
class EngineConfig:
    Nloop: int = 1
    Nadmm: int = 1
    Nepoch: int = 1
    check_results: bool = False
    be_verbose: bool = False
    diagnostics: str = "post"        # 'post' = reference (extra forward after the step, Q17) | 'pre' = reuse closure loss
    graphs: bool = False             # CUDA-graph the Adam minibatch step
    max_minibatches: Optional[int] = None   # cap per round (benchmarks / smoke tests)
    aggregate_in_epoch_loop: bool = True    # reference: aggregation sits inside the epoch loop
    reset_optimizer_each_epoch: bool = False  # no_consensus_multi.py:129-132 recreates Adam every epoch (Q18)
    nan_guard: str = "raise"         # non-finite aggregation residual: 'raise' | 'warn' | 'off' 


class Engine:
    def __init__(self, task: Task, topo: Topology, strategy: Strategy, collective, cfg: EngineConfig,
                 log: Callable[[str], None] = print, metrics: Optional[MetricsLog] = None):
        self.task, self.topo, self.strategy, self.coll, self.cfg = task, topo, strategy, collective, cfg
        self._log = log
        self.metrics = metrics or MetricsLog(None)
        self.timers = PhaseTimers(topo.device)
        alloc = collective.arena_allocator()
        self.replicas: List[Replica] = [task.build_replica(ck, topo.device, alloc) for ck in topo.local_workers]
        for rep in self.replicas:
            for arena in rep.arenas.values():
                collective.register_arena(arena)
        self.optimizers: List = []
        self.images_seen = 0
        self.steps_done = 0
        self.last_epoch = 0
        self._graphs: Dict = {}
        self._adam_cache: Dict = {}
        self.stop_requested = False
        self.step_hook: Optional[Callable[["Engine"], None]] = None
        self.last_loss1: Optional[torch.Tensor] = None
        self.graph_replays = 0
        self.graph_kernel_launches = 0

    # ------------------------------------------------------------------
    def log(self, msg: str, root_only: bool = False) -> None:
        if root_only and not self.topo.is_root:
            return
        self._log(msg)

    # ------------------------------------------------------------------
    def _make_optimizer(self, rep: Replica, visit: Visit):
        arena = rep.arenas[visit.model]
        if visit.optimizer == "adam":
            key = (rep.ck, visit.model, visit.lo, visit.hi)
            opt = self._adam_cache.get(key)
            lr = visit.opt_kwargs.get("lr", 1e-3)
            if opt is None:
                opt = BlockAdam(arena, visit.lo, visit.hi, lr=lr)
                self._adam_cache[key] = opt
            else:
                opt.reset(lr=lr)  # same as a freshly constructed Adam (Q18), but buffers/graphs persist
            return opt
        if visit.optimizer == "lbfgs":
            return LBFGSNew(arena.params[visit.lo: visit.hi + 1], **visit.opt_kwargs)
        raise ValueError("unknown optimizer %r" % visit.optimizer)

    # ------------------------------------------------------------------
    def _train_step(self, rep: Replica, opt, visit: Visit, batch, pen: Penalty) -> torch.Tensor:
        """One ``opt.step(closure)`` + diagnostics; returns the diagnostics loss (0-dim, on device)."""
        task, cfg = self.task, self.cfg
        pre_loss = [None]
        if isinstance(opt, BlockAdam):
            opt.set_penalty(pen.z, pen.y, pen.rho, visit.lambda1, visit.lambda2, pen.rho_dev)

            def closure():
                opt.zero_grad()
                loss = task.loss(rep, batch)
                loss.backward()
                pre_loss[0] = loss.detach()
                return loss

            opt.step(closure)
        else:
            x = rep.block(visit)
            g = rep.block_grad(visit)
            has_pen = pen.z is not None or visit.lambda1 != 0.0 or visit.lambda2 != 0.0

            def closure():
                if torch.is_grad_enabled():
                    opt.zero_grad()
                loss = task.loss(rep, batch)
                if loss.requires_grad:
                    loss.backward()
                    if has_pen:
                        flatops.add_penalty_grad_(g, x, pen.z, pen.y, pen.rho, visit.lambda1, visit.lambda2)
                total = loss.detach()
                if pre_loss[0] is None:
                    pre_loss[0] = total
                if has_pen:
                    total = total + flatops.penalty_value(x, pen.z, pen.y, pen.rho, visit.lambda1, visit.lambda2)
                return total

            opt.step(closure)
        if cfg.diagnostics == "post":
            with torch.no_grad():
                return task.loss(rep, batch).detach()
        return pre_loss[0]

    # ------------------------------------------------------------------
    def run(self) -> Dict:
        cfg, task, strat = self.cfg, self.task, self.strategy
        t0 = time.time()
        for nloop in range(cfg.Nloop):
            for visit in task.visits(nloop):
                if self.stop_requested:
                    break
                self._run_visit(nloop, visit)
            if self.stop_requested:
                break
        self.log("Finished Training", root_only=True)
        return {"images_seen": self.images_seen, "steps": self.steps_done, "wall_s": time.time() - t0}

    def _run_visit(self, nloop: int, visit: Visit) -> None:
        cfg, task, strat = self.cfg, self.task, self.strategy
        for rep in self.replicas:
            rep.set_trainable(visit)
        xs = [rep.block(visit) for rep in self.replicas]
        arena0 = self.replicas[0].arenas[visit.model]
        N = arena0.count(visit.lo, visit.hi)
        strat.begin_block(visit.ci, N, xs)
        self.optimizers = [self._make_optimizer(rep, visit) for rep in self.replicas]
        for nadmm in range(cfg.Nadmm):
            for epoch in range(cfg.Nepoch):
                self.last_epoch = epoch
                if cfg.reset_optimizer_each_epoch and epoch > 0:
                    self.optimizers = [self._make_optimizer(rep, visit) for rep in self.replicas]
                task.on_epoch_start(epoch, self)
                for i_rep, rep in enumerate(self.replicas):
                    self._run_shard(rep, self.optimizers[i_rep], visit, strat.penalty(i_rep), nloop, epoch, N)
                    if self.stop_requested:
                        return
                if cfg.aggregate_in_epoch_loop:
                    self._aggregate(visit, nloop, nadmm, epoch, N)
            if not cfg.aggregate_in_epoch_loop:
                self._aggregate(visit, nloop, nadmm, cfg.Nepoch - 1, N)

    def _run_shard(self, rep: Replica, opt, visit: Visit, pen: Penalty, nloop: int, epoch: int, N: int) -> None:
        cfg, task = self.cfg, self.task
        running = None
        for i, batch in enumerate(task.batches(rep, visit, epoch)):
            if cfg.max_minibatches is not None and i >= cfg.max_minibatches:
                break
            with self.timers.phase("step"):
                if cfg.graphs and isinstance(opt, BlockAdam) and rep.device.type == "cuda":
                    loss1 = self._graphed_step(rep, opt, visit, batch, pen)
                else:
                    loss1 = self._train_step(rep, opt, visit, batch, pen)
            running = loss1 if running is None else running + loss1
            self.last_loss1 = loss1
            self.images_seen += task.batch_size_of(batch)
            self.steps_done += 1
            task.after_minibatch(rep, visit, batch, i, epoch, nloop, N, loss1, self)
            if self.step_hook is not None:
                self.step_hook(self)
                if self.stop_requested:
                    break
        rep.running_loss = float(running) if running is not None else 0.0

    def _aggregate(self, visit: Visit, nloop: int, nadmm: int, epoch: int, N: int) -> None:
        with self.timers.phase("aggregate"):
            metrics = self.strategy.aggregate(nadmm)
        ctx = {"nloop": nloop, "nadmm": nadmm, "epoch": epoch, "N": N, "rho_mean": self.strategy.rho_mean()}
        self._check_finite(visit, metrics, ctx)
        if metrics:
            self.task.aggregate_log(visit, metrics, ctx, self)
            self.metrics.write(dict(kind="round", block=visit.ci, label=list(visit.label), model=visit.model, **ctx, **metrics))
        if self.cfg.check_results:
            with self.timers.phase("eval"):
                acc = self.task.evaluate(self.replicas, self)
            if acc is not None:
                self.metrics.write(dict(kind="eval", block=visit.ci, **ctx, accuracy=acc))

    def _check_finite(self, visit: Visit, metrics: Optional[Dict[str, float]], ctx: Dict) -> None:
        """Failure detection: the residuals are norms over the REDUCED vector of every worker, so one
        NaN/Inf anywhere in any replica's block shows up here, one round after it appeared, on every rank."""
        if not metrics or self.cfg.nan_guard == "off":
            return
        bad = [k for k in ("dual", "primal") if k in metrics and not math.isfinite(float(metrics[k]))]
        if not bad:
            return
        msg = ("non-finite %s residual after aggregating block %s (ids %s) at loop %d, round %d, epoch %d on rank %d: "
               "a replica diverged (NaN/Inf in its parameters, duals or consensus vector)"
               % ("/".join(bad), visit.ci, list(visit.label), ctx["nloop"], ctx["nadmm"], ctx["epoch"], self.topo.rank))
        self.metrics.write(dict(kind="fault", block=visit.ci, **ctx, nonfinite=bad))
        if self.cfg.nan_guard == "raise":
            raise FloatingPointError(msg)
        self.log("WARNING: " + msg)

    # ------------------------------------------------------------------
    # CUDA-graphed Adam step
    # ------------------------------------------------------------------
    def _graphed_step(self, rep: Replica, opt: BlockAdam, visit: Visit, batch, pen: Penalty) -> torch.Tensor:
        from .graphs import GraphedAdamStep

        key = (rep.ck, visit.model, visit.lo, visit.hi, tuple(tuple(t.shape) for t in batch if torch.is_tensor(t)))
        gs = self._graphs.get(key)
        if gs is None:
            gs = GraphedAdamStep(self, rep, opt, visit, batch, pen)
            self._graphs[key] = gs
        return gs.run(batch, pen)
