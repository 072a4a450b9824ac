"""One launch of the weight-stationary conv (layer1 shape) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops
dev = torch.device("cuda", 0)
x = torch.randn(128, 32, 32, 64, device=dev)
w = torch.randn(64, 3, 3, 64, device=dev) / 24.0
st = torch.zeros(128, device=dev)
for _ in range(3):
    y = cuda_ops.conv2d_nhwc(x, w, st, 1, 1)
torch.cuda.synchronize()
