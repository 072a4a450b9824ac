"""GPU checks of code paths that were written after this round's GPU budget was spent and are therefore OPT-IN
(environment switches, default off).  They are skipped unless ``FEDB200_EXPERIMENTAL=1`` so that the default suite only
contains paths that have been confirmed on a B200; run them first thing next round:

    FEDB200_EXPERIMENTAL=1 FEDB200_CONV_ACT=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import math
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FEDB200_EXPERIMENTAL", "0") != "1", reason="experimental paths are opt-in")]

if torch.cuda.is_available():
    from federated_pytorch_test_b200.ops import cuda_ops
DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


# VAE encoder (simple_models.py:249-255) and CPC encoder incl. the dilated first layer (:441-451)
@pytest.mark.parametrize("B,H,Ci,Co,k,s,p,d", [
    (16, 32, 3, 12, 4, 2, 1, 1), (16, 16, 12, 24, 4, 2, 1, 1), (16, 8, 24, 48, 4, 2, 1, 1), (16, 4, 48, 96, 4, 2, 1, 1),
    (32, 32, 8, 8, 4, 2, 1, 1), (32, 32, 8, 8, 4, 2, 3, 2), (32, 32, 8, 8, 4, 2, 6, 4), (32, 16, 40, 64, 4, 2, 1, 1),
    (32, 8, 64, 128, 4, 2, 1, 1), (32, 4, 128, 256, 4, 2, 1, 1)])
@pytest.mark.parametrize("act", [True, False])
def test_conv_bias_act_forward_backward(B, H, Ci, Co, k, s, p, d, act):
    torch.manual_seed(B + H + Ci + Co + d)
    conv = nn.Conv2d(Ci, Co, k, stride=s, padding=p, dilation=d).to(DEV)
    x = torch.randn(B, Ci, H, H, device=DEV, requires_grad=True)
    assert cuda_ops.conv_act_supported(x, conv), "set FEDB200_CONV_ACT=1"
    y = cuda_ops.conv_act(x, conv, act)
    ref = conv(x)
    ref = F.elu(ref) if act else ref
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 3e-3
    g = torch.randn_like(ref)
    gx, gw, gb = torch.autograd.grad(y, (x, conv.weight, conv.bias), g)
    rx, rw, rb = torch.autograd.grad(ref, (x, conv.weight, conv.bias), g)
    assert rel_err(gx, rx) < 5e-3 and rel_err(gw, rw) < 5e-3 and rel_err(gb, rb) < 5e-3
