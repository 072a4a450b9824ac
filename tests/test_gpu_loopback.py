"""Cross-rank code of the fused collectives on ONE GPU: W virtual ranks in one process (parallel/loopback.py).

Covers what a single-GPU box otherwise never executes: per-CTA flag barriers through peer control pads, one-shot and
two-shot (slice ownership + P2P broadcast stores) aggregation, payload exchange of the residual scalars, the
Barzilai-Borwein row gather + deterministic replay, and CUDA-graph capture of a whole aggregation round.
Oracle: plain PyTorch on the gathered tensors ("Collective").
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from federated_pytorch_test_b200.algo.strategies import BBConfig  # noqa: E402
from federated_pytorch_test_b200.parallel import Topology, TorchCollective  # noqa: E402
from federated_pytorch_test_b200.parallel.fused import FusedCollective  # noqa: E402
from federated_pytorch_test_b200.parallel.loopback import LoopbackWorld  # noqa: E402

DEV = torch.device("cuda", 0)


def _slices(world, N, seed):
    stride = -(-N // 32) * 32
    xs = [t[:N] for t in world.alloc(stride)]
    g = torch.Generator(device=DEV).manual_seed(seed)
    for x in xs:
        x.copy_(torch.randn(N, device=DEV, generator=g))
    return xs


@pytest.mark.parametrize("W", [2, 4])
@pytest.mark.parametrize("N", [850, 5130, 73984, 295424])
@pytest.mark.parametrize("two_shot", ["0", "1"])
def test_loopback_fedavg_fedprox_admm(W, N, two_shot):
    world = LoopbackWorld(W, DEV, max_blocks=8, timeout_s=10.0)
    for c in world.colls:
        c.two_shot_mode = two_shot
    xs = _slices(world, N, 7 * N + W)
    zs = [c.zeros_like_block(x, "z") for c, x in zip(world.colls, xs)]
    g = torch.Generator(device=DEV).manual_seed(N)
    z0 = torch.randn(N, device=DEV, generator=g)
    for z in zs:
        z.copy_(z0)
    # ---- FedAvg ----
    mean = torch.stack(xs).mean(0)
    dual_ref = float(torch.dot(z0 - mean, z0 - mean))
    world.run(lambda r, c: c._launch(0, [xs[r]], None, zs[r], 0.0))
    for r, c in enumerate(world.colls):
        v = c.read_record()
        assert v[0] == pytest.approx(dual_ref, rel=1e-4)
        assert bool(v[6]) == (two_shot == "1")
        torch.testing.assert_close(xs[r], mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(zs[r], mean, rtol=1e-5, atol=1e-6)
    # ---- FedProx (no write-back, primal residual summed over ranks) ----
    for x in xs:
        x.add_(torch.randn(N, device=DEV, generator=g))
    zold = zs[0].clone()
    mean = torch.stack(xs).mean(0)
    rho = 1.5
    primal_ref = sum(float(torch.norm(rho * (x - mean))) for x in xs)
    keep = [x.clone() for x in xs]
    world.run(lambda r, c: c._launch(1, [xs[r]], None, zs[r], rho))
    for r, c in enumerate(world.colls):
        v = c.read_record()
        assert v[0] == pytest.approx(float(torch.dot(zold - mean, zold - mean)), rel=1e-4)
        assert v[1] == pytest.approx(primal_ref, rel=1e-4)
        torch.testing.assert_close(zs[r], mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(xs[r], keep[r], rtol=0, atol=0)
    # ---- ADMM, two rounds, rho from device memory ----
    ys = [c.zeros_like_block(x, "y") for c, x in zip(world.colls, xs)]
    yr = [torch.zeros(N, device=DEV) for _ in range(W)]
    rho_dev = torch.full((1,), 0.1, device=DEV)
    zr = zs[0].clone()
    for _ in range(2):
        znew = sum(y + 0.1 * x for x, y in zip(xs, yr)) / (W * 0.1)
        dual_ref = float(torch.dot(zr - znew, zr - znew))
        primal_ref = 0.0
        for x, y in zip(xs, yr):
            yd = 0.1 * (x - znew)
            primal_ref += float(torch.norm(yd))
            y.add_(yd)
        zr = znew
        world.run(lambda r, c: c._launch(2, [xs[r]], [ys[r]], zs[r], 123.0, rho_dev))     # host rho ignored
        for r, c in enumerate(world.colls):
            v = c.read_record()
            assert v[0] == pytest.approx(dual_ref, rel=1e-3, abs=1e-6)
            assert v[1] == pytest.approx(primal_ref, rel=1e-4)
            assert v[4] == pytest.approx(0.1)
            torch.testing.assert_close(zs[r], zr, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(ys[r], yr[r], rtol=1e-4, atol=1e-5)


def test_loopback_co_resident_replicas_per_rank():
    """K = 4 workers on 2 virtual ranks (two replicas each): one-shot path with 4 contributions."""
    W, K, N = 2, 4, 10164
    world = LoopbackWorld(W, DEV, max_blocks=8, timeout_s=10.0, K=K)
    a, b = _slices(world, N, 1), _slices(world, N, 2)          # replica j=0 and j=1 of every rank
    zs = [torch.zeros(N, device=DEV) for _ in range(W)]
    mean = (torch.stack(a).sum(0) + torch.stack(b).sum(0)) / K
    world.run(lambda r, c: c._launch(0, [a[r], b[r]], None, zs[r], 0.0))
    for r, c in enumerate(world.colls):
        c.read_record()
        for t in (a[r], b[r], zs[r]):
            torch.testing.assert_close(t, mean, rtol=1e-5, atol=1e-6)


def test_loopback_nonfinite_is_counted():
    W, N = 2, 4096
    world = LoopbackWorld(W, DEV, max_blocks=4, timeout_s=10.0)
    xs = _slices(world, N, 3)
    xs[1][17] = float("nan")
    zs = [torch.zeros(N, device=DEV) for _ in range(W)]
    world.run(lambda r, c: c._launch(0, [xs[r]], None, zs[r], 0.0))
    for c in world.colls:
        v = c.read_record()
        assert v[2] >= 1.0 and not math.isfinite(v[0])


def test_barrier_timeout_is_reported_not_fatal():
    """Only rank 0 of 2 launches: its barrier times out, the status names the missing rank, the context survives."""
    from federated_pytorch_test_b200.parallel.fused import CollectiveTimeout

    world = LoopbackWorld(2, DEV, max_blocks=2, timeout_s=0.2)
    xs = _slices(world, 4096, 5)
    z = torch.zeros(4096, device=DEV)
    world.colls[0]._launch(0, [xs[0]], None, z, 0.0)
    with pytest.raises(CollectiveTimeout, match="waiting for rank 1"):
        world.colls[0].read_record()
    assert float(torch.ones(4, device=DEV).sum()) == 4.0        # CUDA context still alive


@pytest.mark.parametrize("W", [1, 2])
def test_bb_update_kernel_matches_reference_rule(W):
    """Six dots + gather + sequential accept/reject + carry-forward in one kernel vs the ATen/Python oracle
    (/root/reference/src/consensus_multi.py:242-278)."""
    K, N = 4, 48120
    cfg = BBConfig(enabled=True)
    g = torch.Generator(device=DEV).manual_seed(11)
    if W == 1:
        topo = Topology.single_process(K, DEV)
        colls = [FusedCollective(topo)]
        locals_ = [list(range(K))]
    else:
        world = LoopbackWorld(W, DEV, max_blocks=8, timeout_s=10.0, K=K)
        colls = world.colls
        locals_ = [c.topo.local_workers for c in colls]
    full = {}
    for name, scale in (("x", 1.0), ("y", 0.05), ("yh", 0.05), ("x0", 1.0)):
        full[name] = [scale * torch.randn(N, device=DEV, generator=g) for _ in range(K)]
    for k in range(K):
        full["x0"][k] = full["x"][k] + 0.05 * torch.randn(N, device=DEV, generator=g)
    z = torch.randn(N, device=DEV, generator=g) * 0.1 + torch.stack(full["x"]).mean(0)
    # oracle: the whole K-worker problem in one process through TorchCollective
    otopo = Topology.single_process(K, DEV)
    oracle = TorchCollective(otopo)
    o = {k: [t.clone() for t in v] for k, v in full.items()}
    rows_ref = oracle.bb_update_(o["x"], o["y"], o["yh"], o["x0"], z, 0.1, None, cfg)
    per_rank = []
    for c, loc in zip(colls, locals_):
        per_rank.append(dict(x=[full["x"][k].clone() for k in loc], y=[full["y"][k].clone() for k in loc],
                             yh=[full["yh"][k].clone() for k in loc], x0=[full["x0"][k].clone() for k in loc],
                             rho=torch.full((1,), 0.1, device=DEV)))
    if W == 1:
        d = per_rank[0]
        colls[0]._bb_launch(d["x"], d["y"], d["yh"], d["x0"], z, d["rho"], cfg, False)
    else:
        world.run(lambda r, c: c._bb_launch(per_rank[r]["x"], per_rank[r]["y"], per_rank[r]["yh"], per_rank[r]["x0"], z,
                                            per_rank[r]["rho"], cfg, False))
    torch.cuda.synchronize()
    for c, loc, d in zip(colls, locals_, per_rank):
        rows = c.bb_log[: 8 * K].view(K, 8).tolist()
        for k in range(K):
            for q in range(8):
                assert rows[k][q] == pytest.approx(rows_ref[k][q], rel=2e-3, abs=1e-5), (k, q)
        assert float(d["rho"]) == pytest.approx(rows_ref[K - 1][7], rel=2e-3)
        for i, k in enumerate(loc):
            torch.testing.assert_close(d["yh"][i], o["yh"][k], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(d["x0"][i], o["x"][k], rtol=0, atol=0)


def test_aggregation_round_is_graph_capturable():
    """No memset, clone or host read inside a round: a FedAvg + an ADMM aggregation (rho in device memory) captured
    into ONE CUDA graph and replayed."""
    K, N = 4, 73984
    topo = Topology.single_process(K, DEV)
    coll = FusedCollective(topo)
    stride = -(-N // 32) * 32
    arena = coll.heap.alloc(K * stride)
    xs = [arena[k * stride: k * stride + N] for k in range(K)]
    g = torch.Generator(device=DEV).manual_seed(3)
    for x in xs:
        x.copy_(torch.randn(N, device=DEV, generator=g))
    ys = [coll.zeros_like_block(x, "y") for x in xs]
    z = coll.zeros_like_block(xs[0], "z")
    rho_dev = torch.full((1,), 0.1, device=DEV)
    coll._launch(2, xs, ys, z, 0.0, rho_dev)                  # warm-up (lazy init) outside the capture
    torch.cuda.synchronize()
    for y in ys:
        y.zero_()
    z.zero_()
    x_keep = [x.clone() for x in xs]
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=st):
        coll._launch(2, xs, ys, z, 0.0, rho_dev)
    e0 = int(coll.sync[0])
    for rep in range(3):
        rho_dev.fill_(0.1 * (rep + 1))                        # a new penalty needs no re-capture
        graph.replay()
        v = coll.read_record()
        assert v[4] == pytest.approx(0.1 * (rep + 1))
    assert int(coll.sync[0]) == e0 + 3
    # oracle of the three rounds
    yr, zr = [torch.zeros(N, device=DEV) for _ in range(K)], torch.zeros(N, device=DEV)
    for rep in range(3):
        rho = 0.1 * (rep + 1)
        zr = sum(y + rho * x for x, y in zip(x_keep, yr)) / (K * rho)
        for x, y in zip(x_keep, yr):
            y.add_(rho * (x - zr))
    torch.testing.assert_close(z, zr, rtol=1e-4, atol=1e-5)
    for a, b in zip(ys, yr):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
