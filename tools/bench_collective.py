"""Aggregation micro-benchmark (BASELINE.md §4 'nccl-allreduce' + fused kernels): for each ResNet18 block size,
device time of one FedAvg aggregation (reduce + 1/K + write-back + dual residual)
  * fused P2P kernel (ld.relaxed.sys from peer memory), * fused NVLS kernel (multimem.ld_reduce),
  * NCCL baseline (all_reduce + div + norm + copy: what `--impl nccl` does),
timed with CUDA events, max over ranks.  Launch:  torchrun --nproc-per-node N tools/bench_collective.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from federated_pytorch_test_b200.parallel import Topology, TorchCollective  # noqa: E402
from federated_pytorch_test_b200.parallel.fused import FusedCollective  # noqa: E402

SIZES = [1856, 73984, 73984, 230144, 295424, 919040, 1180672, 3673088, 4720640, 5130]


def timed(fn, dev, iters=20, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) * 1e3 / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    topo = Topology.from_env()
    dev, W = topo.device, topo.world_size
    fused, base = FusedCollective(topo), TorchCollective(topo)
    rows = []
    for i, N in enumerate(SIZES):
        arena = fused.heap.alloc(-(-N // 32) * 32)
        x = arena[:N]
        x.normal_()
        z = torch.zeros(N, device=dev)
        res = {"block": i, "N": N, "bytes": 4 * N}
        for label, mc in (("fused_p2p_us", False), ("fused_nvls_us", True)):
            fused.use_multimem = mc
            res[label] = timed(lambda: fused.fedavg_([x], z, True), dev)
        res["fused_admm_us"] = None
        y = fused.zeros_like_block(x, "y")
        fused.use_multimem = True
        res["fused_admm_us"] = timed(lambda: fused.admm_([x], [y], z, 0.1), dev)
        xr, zr = x.clone(), z.clone()
        res["nccl_us"] = timed(lambda: base.fedavg_([xr], zr, True), dev)
        raw = x.clone()
        res["nccl_allreduce_only_us"] = timed(lambda: dist.all_reduce(raw), dev)
        # bus bandwidth convention of nccl-tests: 2 (W-1)/W * bytes / time
        for k in ("fused_p2p_us", "fused_nvls_us", "nccl_us"):
            res[k.replace("_us", "_busGBs")] = 2 * (W - 1) / W * 4 * N / (res[k] * 1e-6) / 1e9
        rows.append(res)
        if topo.is_root:
            print(json.dumps(res), flush=True)
    if topo.is_root:
        print(json.dumps({"world": W, "transport": fused.heap.transport, "multicast": bool(fused.heap.allocs[-1]["mc_ptr"])}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
