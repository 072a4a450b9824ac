"""Multi-GPU: the fused NVLink collectives across processes (one rank per GPU) against NCCL all-reduce.
Needs >= 2 GPUs (``gpurun --gpus 2``); skipped otherwise."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip(">= 2 CUDA devices required", allow_module_level=True)


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from federated_pytorch_test_b200.parallel import Topology, TorchCollective
    from federated_pytorch_test_b200.parallel.fused import FusedCollective

    topo = Topology.from_env(world)
    dev = topo.device
    fused, base = FusedCollective(topo), TorchCollective(topo)
    report = {"transport": fused.heap.transport, "multicast": bool(fused.heap.allocs[-1]["mc_ptr"]), "cases": []}
    for use_mc in (True, False):
        fused.use_multimem = use_mc
        for N in (456, 5130, 73984, 1180672):
            g = torch.Generator(device=dev).manual_seed(1000 * rank + N)
            arena = fused.heap.alloc(-(-N // 32) * 32)
            x = arena[:N]
            x.copy_(torch.randn(N, device=dev, generator=g))
            xr = x.clone()
            z, zr = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
            d1 = float(fused.fedavg_([x], z, True))
            d2 = float(base.fedavg_([xr], zr, True))
            ok = abs(d1 - d2) <= 1e-4 * abs(d2) + 1e-6 and torch.allclose(x, xr, rtol=1e-5, atol=1e-6) and torch.allclose(z, zr, rtol=1e-5, atol=1e-6)
            x.add_(torch.randn(N, device=dev, generator=g))
            xr.copy_(x)
            y, yr = fused.zeros_like_block(x, "y"), torch.zeros(N, device=dev)
            for _ in range(2):
                a = fused.admm_([x], [y], z, 0.1)
                b = base.admm_([xr], [yr], zr, 0.1)
                ok = ok and abs(float(a[0]) - float(b[0])) <= 1e-3 * abs(float(b[0])) + 1e-6
                ok = ok and abs(float(a[1]) - float(b[1])) <= 1e-3 * abs(float(b[1])) + 1e-6
                ok = ok and torch.allclose(y, yr, rtol=1e-4, atol=1e-5) and torch.allclose(z, zr, rtol=1e-4, atol=1e-5)
            a = fused.fedprox_([x], z, 1.0)
            b = base.fedprox_([xr], zr, 1.0)
            ok = ok and abs(float(a[1]) - float(b[1])) <= 1e-3 * abs(float(b[1])) + 1e-6
            report["cases"].append((use_mc, N, bool(ok)))
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(report, os.path.join(out_dir, "report.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_fused_collectives_across_ranks(tmp_path, world):
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rep = torch.load(str(tmp_path / "report.pt"), weights_only=False)
    print(rep["transport"], "multicast:", rep["multicast"])
    bad = [c for c in rep["cases"] if not c[2]]
    assert not bad, bad
