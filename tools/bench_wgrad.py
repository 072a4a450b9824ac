"""Weight-gradient micro-benchmark on the ResNet18 sites (batch 128): tcgen05 wgrad kernel (csrc/wgrad_tcgen05.cuh) vs
cuDNN (aten.convolution_backward, TF32, channels_last), device time from a replayed CUDA graph.
Usage on the GPU box:  python tools/bench_wgrad.py [site-filter]      (FEDB200_WGRAD_GPC / FEDB200_WGRAD_SPLITS to sweep)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402
from tools.bench_conv import timed  # noqa: E402

SITES = [  # name, B, H, Cin, Cout, k, stride, pad
    ("stem", 128, 32, 4, 64, 3, 1, 1), ("layer1", 128, 32, 64, 64, 3, 1, 1), ("l2.0.c1", 128, 32, 64, 128, 3, 2, 1),
    ("layer2", 128, 16, 128, 128, 3, 1, 1), ("l2.sc", 128, 32, 64, 128, 1, 2, 0), ("l3.0.c1", 128, 16, 128, 256, 3, 2, 1),
    ("layer3", 128, 8, 256, 256, 3, 1, 1), ("l3.sc", 128, 16, 128, 256, 1, 2, 0), ("l4.0.c1", 128, 8, 256, 512, 3, 2, 1),
    ("layer4", 128, 4, 512, 512, 3, 1, 1), ("l4.sc", 128, 8, 256, 512, 1, 2, 0),
]
PEAK_TF32 = 855.0     # TFLOP/s: half of the measured bf16 cuBLAS burst (MEASURED_PEAKS.json: 1710.4)


def main():
    dev = torch.device("cuda", 0)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    torch.backends.cudnn.allow_tf32 = True
    print("%-8s %10s %10s %9s %9s %8s" % ("site", "ours us", "cudnn us", "ours TF/s", "% of 855", "relerr"))
    for name, B, H, Ci, Co, k, s, p in SITES:
        if flt and flt not in name:
            continue
        Ho = (H + 2 * p - k) // s + 1
        x = torch.randn(B, H, H, Ci, device=dev)
        dy = torch.randn(B, Ho, Ho, Co, device=dev)
        w = torch.zeros(Co, Ci, k, k, device=dev).contiguous(memory_format=torch.channels_last)
        xc, dyc = x.permute(0, 3, 1, 2), dy.permute(0, 3, 1, 2)
        gf = 2.0 * B * Ho * Ho * Co * k * k * Ci / 1e9

        def cudnn():
            return torch.ops.aten.convolution_backward(dyc, xc, w, None, [s, s], [p, p], [1, 1], False, [0, 0], 1, [False, True, False])[1]

        ref = torch.ops.aten.convolution_backward(dyc.double(), xc.double(), w.double(), None, [s, s], [p, p], [1, 1], False, [0, 0], 1,
                                                  [False, True, False])[1]
        buf = torch.zeros(Co, k, k, Ci, device=dev)

        def ours():
            cuda_ops.ext().conv_wgrad(x, dy, buf, s, p, 1)       # accumulates: what the training step does (no fill)

        buf.zero_()
        ours()
        err = float((buf.permute(0, 3, 1, 2).double() - ref).abs().max() / ref.abs().max())
        t_c = timed(cudnn)
        t_o = timed(ours)
        print("%-8s %10.1f %10.1f %9.1f %8.1f%% %8.1e" % (name, t_o, t_c, gf / t_o * 1e3, 100 * gf / t_o * 1e3 / PEAK_TF32, err), flush=True)


if __name__ == "__main__":
    main()
