"""Which inputs make Net's conv weight gradients deviate?  Sweeps seeds; for a failing one compares, layer by layer, our fused
conv+ELU+pool (values, winner indices, dz) with the ATen composition evaluated in fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import cuda_ops, losses
from federated_pytorch_test_b200.ops import functional as FX

DEV = torch.device("cuda", 0)
def rel(u, v): return float((u.double() - v.double()).abs().max() / v.double().abs().max().clamp_min(1e-30))

bad_seed = None
for seed in range(12):
    torch.manual_seed(0)
    a, b = models.Net().to(DEV), models.Net().to(DEV)
    b.load_state_dict(a.state_dict()); b = b.double()
    torch.manual_seed(1000 + seed)
    x = torch.randn(32, 3, 32, 32, device=DEV); y = torch.randint(0, 10, (32,), device=DEV)
    FX.set_fast_path(True)
    la = losses.cross_entropy(a(x), y); la.backward()
    FX.set_fast_path(False)
    lb = losses.cross_entropy(b(x.double()), y); lb.backward()
    errs = {n: rel(pa.grad, pb.grad) for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters())}
    print("seed", seed, {k: "%.1e" % v for k, v in errs.items() if "conv" in k}, flush=True)
    if bad_seed is None and errs["conv2.weight"] > 1e-3:
        bad_seed = seed
        keep = (a, b, x, y)
if bad_seed is not None:
    a, b, x, y = keep
    FX.set_fast_path(True)
    e = cuda_ops.ext()
    # layer 1
    y1, i1 = e.smallconv_fwd(x, a.conv1.weight.contiguous(), a.conv1.bias, 0, True, True)
    z1 = F.conv2d(x.double(), b.conv1.weight, b.conv1.bias)
    r1 = F.elu(z1)
    p1, pi1 = F.max_pool2d(r1, 2, 2, return_indices=True)
    print("layer1 pooled value err", rel(y1, p1))
    # winner index comparison: ours q in {0..3} -> flat index in the 28x28 map
    Hp = 14
    hh = torch.arange(Hp, device=DEV).view(1, 1, Hp, 1); ww = torch.arange(Hp, device=DEV).view(1, 1, 1, Hp)
    ours_flat = (2 * hh + (i1.long() >> 1)) * 28 + 2 * ww + (i1.long() & 1)
    mism = (ours_flat != pi1)
    print("layer1 winner mismatches: %d of %d" % (int(mism.sum()), mism.numel()))
    if int(mism.sum()):
        idx = mism.nonzero()[:5]
        for t in idx:
            n_, c_, h_, w_ = [int(v) for v in t]
            win = r1[n_, c_, 2 * h_:2 * h_ + 2, 2 * w_:2 * w_ + 2]
            print("  window", (n_, c_, h_, w_), win.flatten().tolist(), "ours q", int(i1[n_, c_, h_, w_]), "aten", int(pi1[n_, c_, h_, w_]))
