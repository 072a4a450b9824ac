"""Two launches each of the kernel modes added late in round 2, for `ncu -k regex:...` captures and compute-sanitizer:
shuffle-store stride-2 data gradient, multi-dilation + tap-packed CPC stem, tap-packed RGB / 12-channel convolutions,
pipelined fp32 GEMM (32 x 32 tiles), transposed-conv weight packing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from federated_pytorch_test_b200.ops import cuda_ops
from federated_pytorch_test_b200.ops import functional as FX

dev = torch.device("cuda", 0)
e = cuda_ops.ext()
torch.manual_seed(0)
# (a) layer2.0.conv1 data gradient: dy [128, 16, 16, 128] -> dx [128, 32, 32, 64] through the 5-D store
dy = torch.randn(128, 16, 16, 128, device=dev)
w = torch.randn(128, 3, 3, 64, device=dev) * 0.05
for _ in range(2):
    dx = cuda_ops._s2_dgrad(e, dy, w, True)
# (b) CPC stem: five dilated 4x4 / stride-2 convolutions of 1152 x 8 x 32 x 32 as one launch
convs = [nn.Conv2d(8, 8, 4, stride=2, dilation=d, padding=(3 * d) // 2).to(dev) for d in (1, 2, 4, 8, 16)]
x = torch.randn(1152, 8, 32, 32, device=dev)
for _ in range(2):
    y = FX.dilated_stem(x, convs)
    y.sum().backward()
# (c) VAE encoder convs (3 -> 12 -> 24 channels, 4x4 stride 2): tap packing with 8- and 16-channel sub-tiles
c1, c2 = nn.Conv2d(3, 12, 4, stride=2, padding=1).to(dev), nn.Conv2d(12, 24, 4, stride=2, padding=1).to(dev)
xi = torch.rand(128, 3, 32, 32, device=dev)
for _ in range(2):
    h = FX.conv_act(FX.conv_act(xi, c1), c2)
    h.sum().backward()
# (d) dense layers of the VAE / VAE-CL
for (M, K, N) in ((128, 384, 16), (1280, 394, 128)):
    lin = nn.Linear(K, N).to(dev)
    a = torch.randn(M, K, device=dev, requires_grad=True)
    for _ in range(2):
        cuda_ops.linear_act(a, lin, True).sum().backward()
# (e) transposed-conv weight packing
wt = torch.randn(96, 48, 4, 4, device=dev)
for _ in range(2):
    e.convT_pack(wt)
torch.cuda.synchronize()
print("ok")
