"""Why do Net2's conv gradients differ by a few per cent from an fp64 oracle?  Hypothesis: TF32 operand rounding (1e-3) flips
2x2 max-pool winners, and every flip moves a gradient term.  Test: an fp64 oracle whose convolutions round their operands
to TF32 exactly like the tensor core (forward AND both backward products) must agree with the fast path to ~1e-5 + rare flips.
Prints per-parameter L2 errors against (a) plain fp64, (b) fp64 with truncated TF32 operands, (c) with round-to-nearest TF32
operands, and the number of max-pool winners that differ per stage."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import losses
from federated_pytorch_test_b200.ops import functional as FX
from federated_pytorch_test_b200.utils.tf32_oracle import tf32_conv_oracle

DEV = torch.device("cuda", 0)
print("cudnn.allow_tf32", torch.backends.cudnn.allow_tf32, "matmul precision", torch.get_float32_matmul_precision(),
      "cudnn", torch.backends.cudnn.version(), flush=True)


def l2(u, v):
    return float((u.double() - v.double()).norm() / v.double().norm().clamp_min(1e-30))


for seed in (1000, 1001):
    torch.manual_seed(seed)
    x = torch.randn(32, 3, 32, 32, device=DEV)
    y = torch.randint(0, 10, (32,), device=DEV)
    torch.manual_seed(0)
    a = models.Net2().to(DEV)
    FX.set_fast_path(True)
    la = losses.cross_entropy(a(x), y)
    la.backward()
    for mode in ("fp64", "trunc", "rna"):
        torch.manual_seed(0)
        o = models.Net2().to(DEV)
        o.load_state_dict(a.state_dict())
        o = o.double()
        if mode != "fp64":
            tf32_conv_oracle(o, mode)
        FX.set_fast_path(False)
        lo = losses.cross_entropy(o(x.double()), y)
        lo.backward()
        FX.set_fast_path(True)
        print("seed %d oracle %-5s loss ours %.7f oracle %.7f" % (seed, mode, float(la), float(lo)),
              {n: "%.1e" % l2(pa.grad, po.grad) for (n, pa), (_, po) in zip(a.named_parameters(), o.named_parameters())}, flush=True)
    # winner flips per stage: ours (tcgen05 TF32) vs plain fp64, stage inputs taken from the fp64 chain
    torch.manual_seed(0)
    o = models.Net2().to(DEV)
    o.load_state_dict(a.state_dict())
    o = o.double()
    h32, h64 = x, x.double()
    for name in ("conv1", "conv2", "conv3", "conv4"):
        FX.set_fast_path(True)
        z32 = FX.conv_act_pool(h32, getattr(a, name), True, False)
        FX.set_fast_path(False)
        z64 = F.elu(getattr(o, name)(h64))
        _, i32 = F.max_pool2d(z32.contiguous(), 2, 2, return_indices=True)
        p64, i64 = F.max_pool2d(z64, 2, 2, return_indices=True)
        print("  %s: value err %.1e, winners differing %d of %d" % (name, l2(z32, z64), int((i32 != i64).sum()), i64.numel()), flush=True)
        h64 = p64
        h32 = p64.float()
    FX.set_fast_path(True)
