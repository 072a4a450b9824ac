"""One launch each of: cuBLAS tf32 GEMM, our single-CTA tcgen05 GEMM, our CTA-pair GEMM (same problem) — for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.ops import cuda_ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda", 0)
M, N, K = 8192, 256, 2304
a, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
for _ in range(2):
    (a @ b.t())
    os.environ.update(FEDB200_2CTA="0", FEDB200_CLUSTER="1", FEDB200_BLOCK_N="0")
    cuda_ops.linear_tf32(a, b)
    os.environ.update(FEDB200_2CTA="1", FEDB200_BLOCK_N="128")
    cuda_ops.linear_tf32(a, b)
torch.cuda.synchronize()
