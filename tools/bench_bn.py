"""Micro-benchmark of the BatchNorm/ELU elementwise kernels on the ResNet18 activation shapes (batch 128).
Reports device time per call (CUDA-graph replay) and the effective bandwidth = (tensors read + written) / time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_conv import timed  # noqa: E402
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402

SHAPES = [("layer1", 128 * 32 * 32, 64), ("layer2", 128 * 16 * 16, 128), ("layer3", 128 * 8 * 8, 256), ("layer4", 128 * 4 * 4, 512)]


def main():
    dev = torch.device("cuda", 0)
    e = cuda_ops.ext()
    for name, M, C in SHAPES:
        y = torch.randn(M, C, device=dev)
        res = torch.randn(M, C, device=dev)
        dout = torch.randn(M, C, device=dev)
        gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        mb = M * C * 4 / 1e6
        stats = torch.zeros(2 * C + 1, device=dev)

        def fill_stats():
            stats[:C] = y.sum(0)
            stats[C:2 * C] = (y * y).sum(0)
        for label, r, tensors in (("fwd", None, 2), ("fwd+res", res, 3)):
            fill_stats()
            # self_clean=False: the statistics stay valid across the timed replays
            us = timed(lambda: e.bn_elu_fwd(y, stats, gamma, beta, r, rm, rv, 1e-5, 0.1, True, False))
            print("%-7s %-14s %7.1f us  %6.2f TB/s" % (name, label, us, tensors * mb / us), flush=True)
        out, mean, invstd = e.bn_elu_fwd(y, stats, gamma, beta, res, rm, rv, 1e-5, 0.1, True, False)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        for label, o, want_res, tensors in (("bwd recompute", None, False, 5), ("bwd", out, False, 7), ("bwd+dres", out, True, 8)):
            sb = torch.zeros(2 * C + 1, device=dev)      # self-cleaning per-layer scratch (no memset per call)
            us = timed(lambda: e.bn_elu_bwd(dout, o, y, mean, invstd, gamma, beta, dg, db, want_res, True, sb))
            print("%-7s %-14s %7.1f us  %6.2f TB/s" % (name, label, us, tensors * mb / us), flush=True)
        st2 = torch.zeros(2 * C, device=dev)
        us = timed(lambda: e.col_stats(y, st2))
        print("%-7s %-14s %7.1f us  %6.2f TB/s" % (name, "col_stats", us, mb / us), flush=True)
        us = timed(lambda: y.add_(res))
        print("%-7s %-14s %7.1f us  %6.2f TB/s" % (name, "aten add_", us, 3 * mb / us), flush=True)
        us = timed(lambda: torch.empty_like(y).copy_(res))
        print("%-7s %-14s %7.1f us  %6.2f TB/s" % (name, "aten copy", us, 2 * mb / us), flush=True)


if __name__ == "__main__":
    main()
