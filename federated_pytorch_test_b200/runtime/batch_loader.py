"""Python face of the native batch assembler (``csrc/batch_loader.cpp``).

``BatchAssembler(images, labels, batch)`` owns a ring of pinned staging slots.
``next_batch_to(device)`` waits for the next staged batch, enqueues the H2D copy
of exactly that batch's bytes on a dedicated copy stream, makes the compute
stream wait on it, and hands the slot back to the native threads once the copy
has been consumed.  This is the ingress of the end-to-end path that ``bench.py``
times (``h2d_bytes_per_step``).

Reference equivalent: ``torch.utils.data.DataLoader`` over a ``SubsetRandomSampler``-less shard with CPU
``transforms`` and a synchronous ``.to(mydevice)`` per minibatch (``src/federated_multi.py:60-71``, ``:172-175``).

If the C++ extension cannot be built (no compiler) a pure-PyTorch gather with
the same interface is used; it is functionally identical, just not overlapped.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .. import _ext


class BatchAssembler:
    def __init__(self, images: torch.Tensor, labels: torch.Tensor, batch_size: int, slots: int = 3, threads: int = 2):
        assert not images.is_cuda, "host-resident dataset expected"
        self.images = images.contiguous()
        self.labels = labels.contiguous().to(torch.int64)
        self.batch = int(batch_size)
        self.row_shape = tuple(images.shape[1:])
        pin = torch.cuda.is_available()
        self.slot_images = [torch.empty((self.batch,) + self.row_shape, dtype=torch.uint8, pin_memory=pin) for _ in range(slots)]
        self.slot_labels = [torch.empty(self.batch, dtype=torch.int64, pin_memory=pin) for _ in range(slots)]
        ext = _ext.load("fedb200_runtime", required=False)
        self._native = None
        if ext is not None:
            self._native = ext.BatchAssembler(self.images, self.labels, self.batch, self.slot_images, self.slot_labels, threads)
        self._order: Optional[torch.Tensor] = None
        self._cursor = 0
        self._copy_stream = torch.cuda.Stream() if pin else None
        self._pending: List[Tuple[int, Optional[torch.cuda.Event]]] = []
        self.bytes_copied = 0

    @property
    def native(self) -> bool:
        return self._native is not None

    def start_epoch(self, order: torch.Tensor) -> None:
        self._drain()
        self._order = order.to(torch.int64).contiguous()
        self._cursor = 0
        if self._native is not None:
            self._native.start_epoch(self._order)

    # ------------------------------------------------------------------
    def _stage(self) -> Tuple[int, int]:
        if self._native is not None:
            return self._native.acquire()
        if self._order is None or self._cursor >= self._order.numel():
            return -1, 0
        idx = self._order[self._cursor: self._cursor + self.batch]
        slot = (self._cursor // self.batch) % len(self.slot_images)
        n = idx.numel()
        torch.index_select(self.images, 0, idx, out=self.slot_images[slot][:n])
        self.slot_labels[slot][:n].copy_(self.labels.index_select(0, idx))
        self._cursor += n
        return slot, n

    def _drain(self) -> None:
        for slot, ev in self._pending:
            if ev is not None:
                ev.synchronize()
            if self._native is not None:
                self._native.release(slot)
        self._pending = []

    def next_batch(self) -> Optional[Tuple[torch.Tensor, torch.Tensor, int]]:
        """Host view of the next staged batch: ``(images_u8, labels, slot)`` or None at epoch end."""
        slot, n = self._stage()
        if slot < 0:
            return None
        return self.slot_images[slot][:n], self.slot_labels[slot][:n], slot

    def release(self, slot: int) -> None:
        if self._native is not None:
            self._native.release(slot)

    def next_batch_to(self, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
        got = self.next_batch()
        if got is None:
            raise StopIteration
        u8, lab, slot = got
        if device.type != "cuda":
            out = u8.clone(), lab.clone()
            self.release(slot)
            return out
        # retire slots whose copies have completed (keeps at most slots-1 copies in flight)
        while len(self._pending) >= len(self.slot_images) - 1:
            s, ev = self._pending.pop(0)
            ev.synchronize()
            self.release(s)
        cur = torch.cuda.current_stream(device)
        with torch.cuda.stream(self._copy_stream):
            d_u8 = u8.to(device, non_blocking=True)
            d_lab = lab.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        cur.wait_stream(self._copy_stream)
        d_u8.record_stream(cur)
        d_lab.record_stream(cur)
        self._pending.append((slot, ev))
        self.bytes_copied += u8.numel() * u8.element_size() + lab.numel() * lab.element_size()
        return d_u8, d_lab

    def close(self) -> None:
        self._drain()
        self._native = None
