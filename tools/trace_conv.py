"""Per-role clock64 timeline of CTA 0 of the persistent conv kernel (cycles since kernel entry).
stamps: 0 entry | 1 prologue done | 2,3 producer issued all loads of tile 0,1 | 4 first operands landed |
5,6 MMA warp committed tile 0,1 | 7/8 epilogue got / finished tile 0 | 9/10 same for tile 1 | 11 all roles done"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_conv import SHAPES  # noqa: E402
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402

NAMES = ["entry", "prologue", "prod t0", "prod t1", "1st data", "mma t0", "mma t1", "epi0 start", "epi0 end", "epi1 start",
         "epi1 end", "exit"]


def main():
    dev = torch.device("cuda", 0)
    e = cuda_ops.ext()
    os.environ.update(FEDB200_WS="0", FEDB200_HALO="0", FEDB200_SPLITK="1")
    for dbg in ("0", "7"):
        os.environ["FEDB200_DBG"] = dbg
        for name, B, H, Ci, Co, k, s, p in SHAPES:
            if name in ("stem", "layer1"):
                continue
            x = torch.randn(B, H, H, Ci, device=dev)
            w = torch.randn(Co, k, k, Ci, device=dev)
            buf = torch.zeros(16, dtype=torch.int64, device=dev)
            for _ in range(3):
                cuda_ops.conv2d_nhwc(x, w, None, s, p)
            e.set_conv_trace(buf)
            cuda_ops.conv2d_nhwc(x, w, None, s, p)
            torch.cuda.synchronize()
            e.set_conv_trace(None)
            t = buf.tolist()
            rel = ["%s=%d" % (NAMES[i], t[i] - t[0]) for i in range(1, 12) if t[i] > 0]
            print("dbg=%s %-8s " % (dbg, name) + " ".join(rel), flush=True)


if __name__ == "__main__":
    main()
