import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops
dev = torch.device("cuda", 0)
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
for kps, dbg in (("1", "0"), ("2", "0")):
    os.environ["FEDB200_KPS"] = kps
    os.environ["FEDB200_DBG"] = dbg
    for (M, N, K) in [(128, 64, 1024), (128, 256, 1024)]:
        a = torch.randn(M, K, device=dev)
        out = []
        for kb in range(K // 32):
            b = torch.zeros(N, K, device=dev)
            b[:, kb * 32:(kb + 1) * 32] = torch.randn(N, 32, device=dev)
            y = cuda_ops.linear_tf32(a, b)
            ref = (a.double() @ b.double().t()).float()
            e = rel(y, ref)
            out.append("%d:%s" % (kb, "ok" if e < 3e-3 else ("ZERO" if float(y.abs().max()) == 0 else "%.1e" % e)))
        print("kps", kps, "dbg", dbg, M, N, K, " ".join(out), flush=True)
