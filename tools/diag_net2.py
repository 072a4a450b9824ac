"""Net2 on the fast path vs an fp64 oracle: per-parameter gradient errors, with the tcgen05 weight gradient on / off
(FEDB200_WGRAD) — localises which backward kernel of the conv+bias+ELU+maxpool chain is off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import cuda_ops, losses
from federated_pytorch_test_b200.ops import functional as FX

DEV = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
def rel(u, v): return float((u.double() - v.double()).abs().max() / v.double().abs().max())
for B in (32, 128):
    torch.manual_seed(0)
    a, b = models.Net2().to(DEV), models.Net2().to(DEV)
    b.load_state_dict(a.state_dict()); b = b.double()
    x = torch.randn(B, 3, 32, 32, device=DEV); y = torch.randint(0, 10, (B,), device=DEV)
    FX.set_fast_path(False)
    lb = losses.cross_entropy(b(x.double()), y); lb.backward()
    for wg in (True, False):
        cuda_ops.WGRAD = wg
        for p in a.parameters():
            p.grad = None
        FX.set_fast_path(True)
        la = losses.cross_entropy(a(x), y); la.backward()
        print("B=%d WGRAD=%s loss %.6f vs %.6f" % (B, wg, float(la), float(lb)), {n: "%.1e" % rel(pa.grad, pb.grad) for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters())}, flush=True)
    cuda_ops.WGRAD = True
# layer-level: conv4-like weight gradient with ELU-shaped (mostly positive) inputs
e = cuda_ops.ext()
for (Bn, H, Ci, Co) in ((32, 4, 256, 512), (32, 8, 128, 256), (32, 16, 64, 128), (32, 32, 4, 64)):
    g = torch.Generator(device=DEV).manual_seed(Ci)
    xn = torch.nn.functional.elu(torch.randn(Bn, H, H, Ci, device=DEV, generator=g)) + 0.5
    dy = torch.randn(Bn, H, H, Co, device=DEV, generator=g) * 1e-3
    dw = cuda_ops.conv_wgrad(xn, dy, 3, 3, Ci, 1, 1, 1)
    ref = torch.ops.aten.convolution_backward(dy.permute(0, 3, 1, 2).double(), xn.permute(0, 3, 1, 2).double(), torch.zeros(Co, Ci, 3, 3, device=DEV, dtype=torch.float64),
                                              None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False])[1]
    print("wgrad B=%d H=%d %d->%d relerr %.2e" % (Bn, H, Ci, Co, rel(dw, ref)), flush=True)
