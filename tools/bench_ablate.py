"""Ablation timing of the tcgen05 conv kernels (FEDB200_DBG bits: 1 no TMA loads, 2 no MMAs, 4 no global stores,
8 WS only: atom-aligned tap reads).  Outputs are garbage by construction; only the times matter.  They tell which of
the three pipelines (TMA ingest, tensor core, epilogue) bounds each shape.  Usage: python tools/bench_ablate.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_conv import SHAPES, timed  # noqa: E402
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402

MODES = [("full", 0), ("noTMA", 1), ("noMMA", 2), ("noST", 4), ("noTMA+noST", 5), ("noMMA+noST", 6), ("noTMA+noMMA", 3),
         ("empty", 7), ("aligned", 8)]
PATHS = [("default", dict(FEDB200_WS="1", FEDB200_HALO="1", FEDB200_SPLITK="0", FEDB200_BLOCK_N="0"))]


def main():
    dev = torch.device("cuda", 0)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    print("%-8s %-14s " % ("shape", "path") + " ".join("%11s" % m for m, _ in MODES))
    for name, B, H, Ci, Co, k, s, p in SHAPES:
        if flt and flt not in name:
            continue
        x = torch.randn(B, H, H, Ci, device=dev)
        w = torch.randn(Co, k, k, Ci, device=dev) / (k * k * Ci) ** 0.5
        for pname, env in PATHS:
            os.environ.update(env)
            row = []
            for _, bits in MODES:
                os.environ["FEDB200_DBG"] = str(bits)
                row.append(timed(lambda: cuda_ops.conv2d_nhwc(x, w, None, s, p)))
            print("%-8s %-14s " % (name, pname) + " ".join("%9.1fus" % t for t in row), flush=True)
    os.environ["FEDB200_DBG"] = "0"


if __name__ == "__main__":
    main()
