"""A few launches each of the non-convolution kernels, for `ncu -k regex:...` captures (block_reduce K=8 co-resident,
bb_update, lbfgs_two_loop, gemm_f32, info_nce, act_bwd_bias, adam_prox)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_pytorch_test_b200.algo.strategies import BBConfig
from federated_pytorch_test_b200.ops import cuda_ops, flatops, losses
from federated_pytorch_test_b200.parallel import Topology
from federated_pytorch_test_b200.parallel.fused import FusedCollective

dev = torch.device("cuda", 0)
K = 8
topo = Topology.single_process(K, dev)
coll = FusedCollective(topo)
for n in (1856, 1180672, 4720640):
    st = -(-n // 32) * 32
    arena = coll.heap.alloc(K * st)
    xs = [arena[k * st: k * st + n] for k in range(K)]
    for x in xs:
        x.normal_()
    ys = [coll.zeros_like_block(x, "y") for x in xs]
    z = torch.zeros(n, device=dev)
    rho = torch.full((1,), 0.1, device=dev)
    for _ in range(2):
        coll._launch(0, xs, None, z, 0.0)
        coll._launch(2, xs, ys, z, 0.1, rho)
    yh = [torch.randn(n, device=dev) * 0.01 for _ in range(K)]
    x0 = [x + 0.01 * torch.randn(n, device=dev) for x in xs]
    coll._bb_launch(xs, ys, yh, x0, z, rho, BBConfig(enabled=True), False)
torch.cuda.synchronize()
# L-BFGS two-loop (history 10 over the layer3.1 block), Adam, dense layers, InfoNCE
n = 1180672
hist = flatops.PairHistory(10, torch.zeros(n, device=dev))
for i in range(10):
    hist.push(torch.randn(n, device=dev), torch.randn(n, device=dev))
g = torch.randn(n, device=dev)
for _ in range(2):
    hist.two_loop(g, 1.0)
x, gr, m, v = (torch.randn(n, device=dev) for _ in range(4))
v.abs_()
for t in range(1, 3):
    cuda_ops.adam_prox_step(x, gr, m, v, t, 1e-3, 0.9, 0.999, 1e-8)
lin = torch.nn.Linear(394, 128).to(dev)
a = torch.randn(1280, 394, device=dev, requires_grad=True)
for _ in range(2):
    cuda_ops.linear_act(a, lin, True).sum().backward()
zz, zh = torch.randn(128, 32, 3, 3, device=dev, requires_grad=True), torch.randn(128, 32, 3, 3, device=dev, requires_grad=True)
for _ in range(2):
    losses.info_nce(zz, zh).backward()
torch.cuda.synchronize()
print("ok")
