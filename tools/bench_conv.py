"""Micro-benchmark of the tcgen05 conv kernel variants on the ResNet18 shapes (batch 128).
Usage on the GPU box:  python tools/bench_conv.py [shape-filter]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402

SHAPES = [  # name, B, H, Cin, Cout, k, stride, pad
    ("stem", 128, 32, 4, 64, 3, 1, 1), ("layer1", 128, 32, 64, 64, 3, 1, 1), ("l2.0.c1", 128, 32, 64, 128, 3, 2, 1),
    ("layer2", 128, 16, 128, 128, 3, 1, 1), ("l3.0.c1", 128, 16, 128, 256, 3, 2, 1), ("layer3", 128, 8, 256, 256, 3, 1, 1),
    ("l4.0.c1", 128, 8, 256, 512, 3, 2, 1), ("layer4", 128, 4, 512, 512, 3, 1, 1), ("l2.sc", 128, 32, 64, 128, 1, 2, 0),
]
BASE = dict(FEDB200_HALO="1", FEDB200_CLUSTER="1", FEDB200_BLOCK_N="0", FEDB200_2CTA="0", FEDB200_SPLITK="0", FEDB200_WS="1",
            FEDB200_KPS="2", FEDB200_DBG="0", FEDB200_PERSIST="1", FEDB200_MT="1", FEDB200_MT2_KPS="1")
VARIANTS = [
    ("default", {}),
    ("generic persist", dict(FEDB200_WS="0", FEDB200_HALO="0")),
    ("generic mt1", dict(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_MT="1")),
    ("generic mt2", dict(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_MT="2")),
    ("generic mt2 2x96K", dict(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_MT="2", FEDB200_MT2_KPS="2")),
    ("generic 6x32K", dict(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_KPS="3")),
    ("generic bn128", dict(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_BLOCK_N="128")),
]


def timed(fn, iters=20):
    """Device time per call: the calls are captured into ONE CUDA graph and replayed, so neither the Python wrapper nor
    the tensor-map encode (host side, ~20 us per call) is part of the number."""
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    first = (time.perf_counter() - t0) * 1e6
    if first > 2e5:          # pathological: do not loop
        return first
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters


def main():
    dev = torch.device("cuda", 0)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    torch.backends.cudnn.allow_tf32 = True
    for name, B, H, Ci, Co, k, s, p in SHAPES:
        if flt and flt not in name:
            continue
        x = torch.randn(B, H, H, Ci, device=dev)
        w = torch.randn(Co, k, k, Ci, device=dev) / (k * k * Ci) ** 0.5
        Ho = (H + 2 * p - k) // s + 1
        gf = 2.0 * B * Ho * Ho * Co * k * k * Ci / 1e9
        xc, wc = x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2)   # channels_last views for cuDNN
        ref = torch.nn.functional.conv2d(xc.double(), wc.double(), None, s, p).permute(0, 2, 3, 1).float()
        us = timed(lambda: torch.nn.functional.conv2d(xc, wc, None, s, p))
        print("%-8s %-16s %8.1f us  %7.1f TFLOP/s" % (name, "cudnn tf32 NHWC", us, gf / us * 1e-3), flush=True)
        for label, var in VARIANTS:
            env = dict(BASE)
            env.update(var)
            os.environ.update(env)
            try:
                y = cuda_ops.conv2d_nhwc(x, w, None, s, p)
                torch.cuda.synchronize()
            except Exception as exc:
                print("%-8s %-16s FAILED %s" % (name, label, str(exc)[:120]), flush=True)
                continue
            err = float((y - ref).abs().max() / ref.abs().max())
            us = timed(lambda: cuda_ops.conv2d_nhwc(x, w, None, s, p))
            print("%-8s %-16s %8.1f us  %7.1f TFLOP/s  relerr=%.1e" % (name, label, us, gf / us * 1e-3, err), flush=True)
    # the layer1 conv as a plain 2-D GEMM with the same FLOPs and operand bytes (isolates the 4-D TMA path)
    os.environ.update(BASE)
    M, N, K = 131072, 64, 576
    a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    us = timed(lambda: cuda_ops.linear_tf32(a, b))
    print("gemm 131072x64x576 tcgen05   %8.1f us  %7.1f TFLOP/s" % (us, 2.0 * M * N * K / 1e9 / us * 1e-3), flush=True)
    torch.backends.cuda.matmul.allow_tf32 = True
    us = timed(lambda: a @ b.t())
    print("gemm 131072x64x576 cublas tf32 %6.1f us  %7.1f TFLOP/s" % (us, 2.0 * M * N * K / 1e9 / us * 1e-3), flush=True)
    M, N, K = 8192, 256, 2304
    a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    us = timed(lambda: cuda_ops.linear_tf32(a, b))
    print("gemm 8192x256x2304 tcgen05   %8.1f us  %7.1f TFLOP/s" % (us, 2.0 * M * N * K / 1e9 / us * 1e-3), flush=True)
    us = timed(lambda: a @ b.t())
    print("gemm 8192x256x2304 cublas tf32 %6.1f us  %7.1f TFLOP/s" % (us, 2.0 * M * N * K / 1e9 / us * 1e-3), flush=True)


if __name__ == "__main__":
    main()
