"""Error budget of the whole-network gradient check: fast path (tf32 tcgen05) and ATen (cuDNN, TF32 on / off) against an
fp64 oracle, ResNet18, blocks 0 / 4 / 9 active.  Used to set the tolerances of tests/test_gpu_kernels.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import cuda_ops
from federated_pytorch_test_b200.ops import functional as FX
from federated_pytorch_test_b200.utils import FlatArena, unfreeze_one_block

DEV = torch.device("cuda", 0)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


for B in (16, 64):
    torch.manual_seed(0)
    a, b, c = models.ResNet18().to(DEV), models.ResNet18().to(DEV), models.ResNet18().to(DEV)
    b.load_state_dict(a.state_dict()); c.load_state_dict(a.state_dict())
    c = c.double()
    FlatArena(a, channels_last_weights=True)
    x = torch.randn(B, 3, 32, 32, device=DEV).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (B,), device=DEV)
    for blk in (0, 4, 9):
        for m in (a, b, c):
            unfreeze_one_block(m, blk)
        a._flat_arena.zero_grads()
        for m in (b, c):
            for p in m.parameters():
                p.grad = None
        FX.set_fast_path(True)
        la = cuda_ops.cross_entropy(a(x), y); la.backward()
        FX.set_fast_path(False)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            for p in b.parameters():
                p.grad = None
            lb = F.cross_entropy(b(x), y); lb.backward()
            if tf32:
                gb_tf32 = [None if p.grad is None else p.grad.clone() for p in b.parameters()]
            else:
                gb_fp32 = [None if p.grad is None else p.grad.clone() for p in b.parameters()]
        lc = F.cross_entropy(c(x.double()), y); lc.backward()
        FX.set_fast_path(True)
        ga = [None if p.grad is None else p.grad.contiguous() for p in a.parameters()]
        gc_ = [p.grad for p in c.parameters()]
        e_fast = max(rel(u, v) for u, v in zip(ga, gc_) if v is not None)
        e_tf32 = max(rel(u, v) for u, v in zip(gb_tf32, gc_) if v is not None)
        e_fp32 = max(rel(u, v) for u, v in zip(gb_fp32, gc_) if v is not None)
        e_fast_vs_tf32 = max(rel(u, v) for u, v in zip(ga, gb_tf32) if v is not None)
        print("B=%d block %d: loss fast %.6f fp64 %.6f | grad rel err vs fp64: fast %.2e  cudnn-tf32 %.2e  cudnn-fp32 %.2e | fast vs cudnn-tf32 %.2e"
              % (B, blk, float(la), float(lc), e_fast, e_tf32, e_fp32, e_fast_vs_tf32), flush=True)
