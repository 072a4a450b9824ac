"""Observability: CUDA-event phase timers (device time, max over ranks), JSONL
metrics and the reference's print formats.

The reference has ``import time`` (``src/federated_multi.py:6``) and never calls it; its only outputs are the
``print`` lines whose exact formats are reproduced in :mod:`..utils.legacy_log`.
"""
from __future__ import annotations

import contextlib
import json
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class PhaseTimers:
    """Accumulates device time per named phase with CUDA events (wall clock on CPU).

    Events are recorded on the current stream and resolved lazily in
    :meth:`summary`, so timing adds no host synchronisation to the hot loop.
    """

    def __init__(self, device: torch.device, enabled: bool = True, keep: int = 4096):
        self.device = torch.device(device)
        self.enabled = enabled
        self.cuda = self.device.type == "cuda"
        self._pending: Dict[str, List] = {}
        self._acc_ms: Dict[str, float] = {}
        self._count: Dict[str, int] = {}
        self._keep = keep

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        if self.cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            try:
                yield
            finally:
                b.record()
                lst = self._pending.setdefault(name, [])
                lst.append((a, b))
                if len(lst) > self._keep:
                    self._resolve(name)
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self._acc_ms[name] = self._acc_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
                self._count[name] = self._count.get(name, 0) + 1

    def _resolve(self, name: str) -> None:
        lst = self._pending.get(name, [])
        if lst:
            lst[-1][1].synchronize()
        for a, b in lst:
            self._acc_ms[name] = self._acc_ms.get(name, 0.0) + a.elapsed_time(b)
            self._count[name] = self._count.get(name, 0) + 1
        self._pending[name] = []

    def summary(self, reduce_max: bool = False, names=None) -> Dict[str, Dict[str, float]]:
        """Accumulated device ms / counts per phase.  ``names`` restricts which phases have their pending CUDA events
        resolved now (resolving ~50 step events costs the host ~0.5 ms: the engine does that once per block visit, not in
        the gap between an aggregation and the next round's first launch)."""
        for name in list(self._pending):
            if names is None or name in names:
                self._resolve(name)
        out = {k: {"ms": v, "count": self._count.get(k, 0)} for k, v in self._acc_ms.items()}
        if reduce_max and dist.is_available() and dist.is_initialized():
            keys = sorted(out)
            t = torch.tensor([out[k]["ms"] for k in keys], dtype=torch.float64, device=self.device if self.cuda else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for k, v in zip(keys, t.tolist()):
                out[k]["ms"] = v
        return out


class MetricsLog:
    """Append-only JSONL sink (one object per aggregation round / evaluation)."""

    def __init__(self, path: Optional[str]):
        self.path = path
        self.rows: List[dict] = []
        self._fh = open(path, "a") if path else None

    def write(self, row: dict) -> None:
        row = dict(row)
        row.setdefault("t", time.time())
        self.rows.append(row)
        if self._fh:
            self._fh.write(json.dumps(row, default=float) + "\n")
            self._fh.flush()

    def close(self) -> None:
        if self._fh:
            self._fh.close()
            self._fh = None


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range when CUDA is present (shows up in ncu/nsys timelines), no-op otherwise."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
