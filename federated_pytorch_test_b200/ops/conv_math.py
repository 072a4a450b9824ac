"""Stride-2 convolution data gradient as ONE stride-1 convolution + a pixel shuffle (dgrad).

For ``y = conv(x, W)`` with stride 2 (3x3 / pad 1 or 1x1 / pad 0, reference sites ``src/simple_models.py:137-147``)

    dx[n, 2i+ph, 2j+pw, ci] = sum_{dr,ds in {0,1}} sum_co dy[n, i+dr, j+ds, co] * Wp[(ph,pw,ci), dr, ds, co]

i.e. every output *phase* (ph, pw) is a small stride-1 convolution of ``dy`` anchored at the top-left pixel, and all
four phases together are one 2x2 convolution with ``4*Ci`` output channels whose filter ``Wp`` holds the original taps
(zeros where a phase has no tap):   row phase 0 <- filter row 1 at dr=0;  row phase 1 <- filter row 2 at dr=0 and
filter row 0 at dr=1 (same for columns).  The result ``[N, Ho, Wo, 2, 2, Ci]`` is interleaved into ``[N, 2Ho, 2Wo, Ci]``.

Why: the strided data gradients were the slowest library kernels left in the step (cuDNN ``strided_dgrad``: 104 us for
layer2.0.conv1 alone, profiles/r1_run19_step_kernels.txt) while the same FLOPs run in ~10 us on our implicit-GEMM
kernel; 7/16 of the packed filter is zero padding, which is cheaper than four launches.

Everything here is plain PyTorch (runs on CPU for the tests); the convolution itself is injected.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.nn.functional as F

# (phase, d) -> original filter index along that axis for k=3, pad=1, stride=2; (0, 1) has no tap
_TAP3 = {(0, 0): 1, (1, 0): 2, (1, 1): 0}


def _tap_index_table(k: int) -> torch.Tensor:
    """[2,2,2,2] (ph, pw, dr, ds) -> flat tap index r*k+s into the KRSC filter, k*k = "zero tap"."""
    zero = k * k
    idx = torch.full((2, 2, 2, 2), zero, dtype=torch.long)
    for ph in (0, 1):
        for pw in (0, 1):
            for dr in (0, 1):
                for ds in (0, 1):
                    if k == 3:
                        r, s = _TAP3.get((ph, dr)), _TAP3.get((pw, ds))
                    else:  # k == 1, pad 0: only phase (0,0) at (dr,ds) = (0,0)
                        r = 0 if (ph == 0 and dr == 0) else None
                        s = 0 if (pw == 0 and ds == 0) else None
                    if r is not None and s is not None:
                        idx[ph, pw, dr, ds] = r * k + s
    return idx


def pack_dgrad_s2_weight(w_krsc: torch.Tensor) -> torch.Tensor:
    """``[Co, k, k, Ci]`` (k in {1, 3}) -> ``[4*Ci, 2, 2, Co]`` filter of the equivalent stride-1 convolution over dy."""
    Co, kh, kw, Ci = w_krsc.shape
    assert kh == kw and kh in (1, 3), "stride-2 phase decomposition implemented for 1x1/pad0 and 3x3/pad1"
    # Pure device ops with Python-side indices: no index tensor is copied host -> device, so this is legal inside a
    # CUDA-graph capture (the active block's filters are re-packed every step).
    flat = w_krsc.reshape(Co, kh * kw, Ci)
    zero = w_krsc.new_zeros(Co, Ci)
    idx = _tap_index_table(kh).reshape(-1).tolist()
    sel = torch.stack([flat[:, i] if i < kh * kw else zero for i in idx], dim=1).reshape(Co, 2, 2, 2, 2, Ci)   # co ph pw dr ds ci
    return sel.permute(1, 2, 5, 3, 4, 0).reshape(4 * Ci, 2, 2, Co).contiguous()


def dgrad_s2(dy_nhwc: torch.Tensor, w_packed: torch.Tensor, conv2x2: Callable[[torch.Tensor, torch.Tensor], torch.Tensor]
             ) -> torch.Tensor:
    """dy ``[N, Ho, Wo, Co]``, packed filter -> dx ``[N, 2Ho, 2Wo, Ci]``.  ``conv2x2(x, w)`` must return the 2x2,
    stride-1 convolution anchored top-left with zero padding on the bottom/right, output size = input size."""
    N, Ho, Wo, _ = dy_nhwc.shape
    Ci = w_packed.shape[0] // 4
    y = conv2x2(dy_nhwc, w_packed)                                   # [N, Ho, Wo, 4*Ci]  (ph, pw, ci)
    return y.view(N, Ho, Wo, 2, 2, Ci).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * Ho, 2 * Wo, Ci)


def conv2x2_oracle(x_nhwc: torch.Tensor, w_krsc: torch.Tensor) -> torch.Tensor:
    """ATen oracle of the injected convolution (CPU tests, fall-back)."""
    x = F.pad(x_nhwc.permute(0, 3, 1, 2), (0, 1, 0, 1))
    return F.conv2d(x, w_krsc.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()
