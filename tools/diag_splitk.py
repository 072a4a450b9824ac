import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops
import torch.nn.functional as F
dev = torch.device("cuda", 0)
os.environ.update(FEDB200_WS="0", FEDB200_HALO="0")
for (H, Ci, Co) in ((16, 128, 128), (8, 256, 256)):
    for B in (4, 8, 32, 128):
        g = torch.Generator(device=dev).manual_seed(B)
        x = torch.randn(B, H, H, Ci, device=dev, generator=g)
        w = torch.randn(Co, 3, 3, Ci, device=dev, generator=g) / math.sqrt(9 * Ci)
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), None, 1, 1).permute(0, 2, 3, 1).float()
        for sk in ("1", "2", "4", "0"):
            os.environ["FEDB200_SPLITK"] = sk
            for with_stats in (False, True):
                stats = torch.zeros(2 * Co, device=dev) if with_stats else None
                y = cuda_ops.conv2d_nhwc(x, w, stats, 1, 1)
                torch.cuda.synchronize()
                err = float((y - ref).abs().max() / ref.abs().max())
                ratio = float((y * ref).sum() / (ref * ref).sum())
                print("H=%d C=%d B=%d splitk=%s stats=%d relerr=%.2e proj=%.3f" % (H, Ci, B, sk, with_stats, err, ratio), flush=True)
