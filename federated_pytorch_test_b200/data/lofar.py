"""LOFAR visibility patches for the CPC driver — synthetic source, same tensor contract.

Contract of the reference's ``get_data_minibatch`` (/root/reference/src/
federated_cpc.py:52-108): from an HDF5 group ``measurement/saps/<SAP>`` holding
``visibilities`` int8 ``[nbase, ntime, nfreq, 4, 2]`` and
``visibility_scale_factors`` float32 ``[nbase, nfreq, 4]``, pick ``batch_size``
random baselines, build ``x[B, 8, ntime, nfreq]`` (channel ``2*pol+re/im`` =
int8 value x per-(baseline,freq,pol) scale), cut 32x32 patches at stride 16 and
return ``(patchx, patchy, y[B*patchx*patchy, 8, 32, 32])`` clamped to +-1e6.

Differences by design:

* the source is an in-memory :class:`LofarSource` (synthetic by default; an HDF5
  file is read once if ``h5py`` happens to be installed) resident on the
  device — the reference re-opens the file and loops over baselines and
  polarisations in Python for every minibatch;
* the whole assembly is a handful of batched tensor ops (gather, broadcast
  multiply, ``unfold``);
* patch ordering: the reference writes patch-major rows but later views them as
  batch-major.  ``layout='batch_major'`` (default) is the
  self-consistent order; ``layout='reference'`` reproduces the original.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch


@dataclass
class LofarSource:
    visibilities: torch.Tensor   # int8 [nbase, ntime, nfreq, 4, 2]
    scale: torch.Tensor          # float32 [nbase, nfreq, 4]

    @staticmethod
    def synthetic(nbase: int = 64, ntime: int = 64, nfreq: int = 64, seed: int = 0,
                  device: Optional[torch.device] = None) -> "LofarSource":
        """Smooth fringe patterns + noise, quantised to int8 with per-(baseline,freq,pol) scales."""
        g = torch.Generator().manual_seed(seed)
        t = torch.linspace(0, 1, ntime).view(1, ntime, 1, 1)
        f = torch.linspace(0, 1, nfreq).view(1, 1, nfreq, 1)
        rate = 2.0 + 10.0 * torch.rand(nbase, 1, 1, 4, generator=g)
        delay = 2.0 + 10.0 * torch.rand(nbase, 1, 1, 4, generator=g)
        phase = 2 * np.pi * (rate * t + delay * f)
        amp = 0.5 + torch.rand(nbase, 1, 1, 4, generator=g)
        re = amp * torch.cos(phase) + 0.2 * torch.randn(nbase, ntime, nfreq, 4, generator=g)
        im = amp * torch.sin(phase) + 0.2 * torch.randn(nbase, ntime, nfreq, 4, generator=g)
        vis = torch.stack((re, im), dim=-1)
        peak = vis.abs().amax(dim=(1, 4)).clamp_min(1e-6)            # [nbase, nfreq, 4]
        scale = (peak / 127.0).to(torch.float32)
        q = torch.round(vis / scale.view(nbase, 1, nfreq, 4, 1)).clamp_(-127, 127).to(torch.int8)
        src = LofarSource(q, scale)
        return src.to(device) if device is not None else src

    @staticmethod
    def from_h5(filename: str, SAP: str = "0") -> "LofarSource":
        import h5py  # optional dependency

        with h5py.File(filename, "r") as f:
            grp = f["measurement"]["saps"][SAP]
            vis = torch.from_numpy(np.asarray(grp["visibilities"]))
            sc = torch.from_numpy(np.asarray(grp["visibility_scale_factors"]))
        return LofarSource(vis.to(torch.int8), sc.to(torch.float32))

    def to(self, device) -> "LofarSource":
        return LofarSource(self.visibilities.to(device), self.scale.to(device))

    @property
    def nbase(self) -> int:
        return self.visibilities.shape[0]


def get_data_minibatch(source: LofarSource, batch_size: int = 2, patch_size: int = 32,
                       generator: Optional[torch.Generator] = None,
                       layout: str = "batch_major") -> Tuple[int, int, torch.Tensor]:
    """Random-baseline minibatch of overlapping patches; see module docstring."""
    vis, sc = source.visibilities, source.scale
    nbase, ntime, nfreq = vis.shape[0], vis.shape[1], vis.shape[2]
    pick = torch.randint(0, nbase, (batch_size,), generator=generator).to(vis.device)
    v = vis.index_select(0, pick).to(torch.float32)                      # [B,T,F,4,2]
    v = v * sc.index_select(0, pick).view(batch_size, 1, nfreq, 4, 1)    # scale per (b,f,pol)
    x = v.permute(0, 3, 4, 1, 2).reshape(batch_size, 8, ntime, nfreq)    # channel = 2*pol + {re,im}
    stride = patch_size // 2
    y = x.unfold(2, patch_size, stride).unfold(3, patch_size, stride)    # [B,8,px,py,ps,ps]
    px, py = y.shape[2], y.shape[3]
    if layout == "reference":      # rows ordered (patch, baseline)
        y = y.permute(2, 3, 0, 1, 4, 5)
    elif layout == "batch_major":  # rows ordered (baseline, patch) — matches the later view(batch, px, py, -1)
        y = y.permute(0, 2, 3, 1, 4, 5)
    else:
        raise ValueError("layout must be 'batch_major' or 'reference'")
    y = y.reshape(batch_size * px * py, 8, patch_size, patch_size).contiguous()
    return px, py, y.clamp_(-1e6, 1e6)
