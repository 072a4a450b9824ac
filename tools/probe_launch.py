"""Launch-floor probe: device time per kernel (CUDA-graph replay of 20 launches) of kernels that do NOTHING, as a function
of their dynamic shared memory size, alone and alternating with a streaming kernel that wants the L1 configuration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_conv import timed  # noqa: E402
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402


def main():
    e = cuda_ops.ext()
    dev = torch.device("cuda", 0)
    buf = torch.zeros(1 << 20, device=dev)
    for grid in (148, 296):
        for kb in (0, 16, 48, 100, 164, 200, 227):
            t0 = timed(lambda: e.probe_launch(0, grid, kb * 1024))
            t1 = timed(lambda: e.probe_launch(1, grid, kb * 1024))
            print("empty kernel   grid=%3d smem=%3d KB : %6.2f us (touching smem %6.2f us)" % (grid, kb, t0, t1), flush=True)
    ts = timed(lambda: buf.add_(1.0))
    print("aten add_ (4 MB) alone                 : %6.2f us" % ts, flush=True)
    for kb in (0, 48, 200, 227):
        def pair():
            e.probe_launch(1, 148, kb * 1024)
            buf.add_(1.0)
        t = timed(pair)
        print("empty(smem=%3d KB) + add_ pair          : %6.2f us per pair (%.2f over add_ alone)" % (kb, t, t - ts), flush=True)
    for grid, kb in ((148, 1), (148, 200), (296, 200), (1024, 200)):
        t = timed(lambda: e.probe_launch(2, grid, kb * 1024))
        print("GEMM prologue skeleton grid=%4d smem=%3d KB : %6.2f us" % (grid, kb, t), flush=True)


if __name__ == "__main__":
    main()
