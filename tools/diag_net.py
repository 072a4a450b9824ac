"""Per-parameter gradient error of Net on the fast path against an fp64 oracle (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import losses
from federated_pytorch_test_b200.ops import functional as FX

DEV = torch.device("cuda", 0)
torch.manual_seed(0)
a, b = models.Net().to(DEV), models.Net().to(DEV)
b.load_state_dict(a.state_dict()); b = b.double()
x = torch.randn(32, 3, 32, 32, device=DEV); y = torch.randint(0, 10, (32,), device=DEV)
FX.set_fast_path(True)
xa = x.clone().requires_grad_()
feat = a.features(xa)
la = losses.cross_entropy(a(xa), y); la.backward()
FX.set_fast_path(False)
xb = x.double().requires_grad_()
lb = losses.cross_entropy(b(xb), y); lb.backward()
print("loss", float(la), float(lb))
def rel(u, v): return float((u.double() - v).abs().max() / v.abs().max())
print("dx", rel(xa.grad, xb.grad))
for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
    print(n, rel(pa.grad, pb.grad))
# stage by stage forward
FX.set_fast_path(True)
fa = a.features(x)
FX.set_fast_path(False)
fb = b.features(x.double())
print("features", rel(fa, fb))
