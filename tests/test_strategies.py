"""Aggregation math against literal transcriptions of the reference formulas,
and the closed-form penalty gradient against autograd."""
import math

import pytest
import torch

from federated_pytorch_test_b200.algo.strategies import ADMM, BBConfig, FedAvg, FedProx
from federated_pytorch_test_b200.ops import flatops
from federated_pytorch_test_b200.parallel import Topology, TorchCollective


def _setup(K=4, N=257, seed=0):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(N, generator=g) for _ in range(K)]
    topo = Topology.single_process(K, "cpu")
    return xs, topo, TorchCollective(topo)


def test_fedavg():
    xs, topo, coll = _setup()
    ref = [x.clone() for x in xs]
    s = FedAvg(coll, topo)
    s.begin_block(0, 257, xs)
    m = s.aggregate(0)
    znew = sum(ref) / 4
    assert m["dual"] == pytest.approx(float(torch.norm(torch.zeros(257) - znew)) / 257, rel=1e-6)
    for x in xs:
        torch.testing.assert_close(x, znew)
    m2 = s.aggregate(1)
    assert m2["dual"] == pytest.approx(0.0, abs=1e-9)


def test_fedprox():
    xs, topo, coll = _setup()
    ref = [x.clone() for x in xs]
    s = FedProx(coll, topo, num_blocks=3, rho0=1.5)
    s.begin_block(1, 257, xs)
    assert s.penalty(0).rho == 1.5 and float(s.penalty(0).z.abs().sum()) == 0.0  # Q6: z starts at 0
    m = s.aggregate(0)
    z = sum(ref) / 4
    primal = sum(float(torch.norm(1.5 * (x - z))) for x in ref) / 257
    assert m["primal"] == pytest.approx(primal, rel=1e-5)
    for x, r in zip(xs, ref):
        assert torch.equal(x, r)  # Q7: no write-back


def _reference_admm_round(xs, ys, z, rho, K, N):
    znew = torch.zeros_like(z)
    for k in range(K):
        znew = znew + ys[k] + rho * xs[k]
    znew = znew / (K * rho)
    dual = float(torch.norm(z - znew)) / N
    primal = 0.0
    for k in range(K):
        yd = rho * (xs[k] - znew)
        primal += float(torch.norm(yd))
        ys[k].add_(yd)
    return znew, primal / N, dual


def test_admm_rounds_match_transcription():
    xs, topo, coll = _setup(K=3, N=100)
    s = ADMM(coll, topo, num_blocks=2, rho0=0.1)
    s.begin_block(0, 100, xs)
    ys_ref = [torch.zeros(100) for _ in range(3)]
    z_ref = torch.zeros(100)
    g = torch.Generator().manual_seed(5)
    for r in range(3):
        for x in xs:
            x.add_(0.1 * torch.randn(100, generator=g))
        z_ref, p, d = _reference_admm_round(xs, ys_ref, z_ref, 0.1, 3, 100)
        m = s.aggregate(r)
        assert m["primal"] == pytest.approx(p, rel=1e-5) and m["dual"] == pytest.approx(d, rel=1e-5)
        torch.testing.assert_close(s.z, z_ref)
        for a, b in zip(s.ys, ys_ref):
            torch.testing.assert_close(a, b)


def _reference_bb(xs, ys, yhat0, x0, z, rho, cfg):
    """Sequential shared-rho rule exactly as in consensus_multi.py:248-278."""
    for k in range(len(xs)):
        yhat1 = ys[k] + rho * (xs[k] - z)
        dy, dx = yhat1 - yhat0[k], xs[k] - x0[k]
        d11, d12, d22 = float(dy.dot(dy)), float(dy.dot(dx)), float(dx.dot(dx))
        rhonew = rho
        if abs(d12) > cfg.epsilon and d11 > cfg.epsilon and d22 > cfg.epsilon:
            alpha = d12 / math.sqrt(d11 * d22)
            aSD, aMG = d11 / d22, d12 / d22
            ahat = aMG if 2 * aMG > aSD else aSD - 0.5 * aMG
            if alpha >= cfg.alphacorrmin and ahat < cfg.rhomax:
                rhonew = ahat
        rho = rhonew
        yhat0[k], x0[k] = yhat1, xs[k].clone()
    return rho


def test_bb_replay_equals_sequential_rule():
    torch.manual_seed(3)
    K, N = 4, 300
    cfg = BBConfig(enabled=True, rhomax=10.0, alphacorrmin=-1.0, epsilon=1e-9)  # permissive: force updates
    xs, topo, coll = _setup(K, N, seed=2)
    s = ADMM(coll, topo, 1, rho0=0.1, bb=cfg, log=lambda m: None)
    s.begin_block(0, N, xs)
    yhat0 = [x.clone() for x in xs]
    x0 = [torch.zeros(N) for _ in xs]
    ys_ref, z_ref, rho_ref = [torch.zeros(N) for _ in xs], torch.zeros(N), 0.1
    g = torch.Generator().manual_seed(9)
    for r in range(5):
        for x in xs:
            x.add_(0.05 * torch.randn(N, generator=g))
        if r == 0:
            x0 = [x.clone() for x in xs]
        elif r % 2 == 0:
            rho_ref = _reference_bb(xs, ys_ref, yhat0, x0, z_ref, rho_ref, cfg)
        z_ref, p, d = _reference_admm_round(xs, ys_ref, z_ref, rho_ref, K, N)
        m = s.aggregate(r)
        assert float(s.rho[0, 0]) == pytest.approx(rho_ref, rel=1e-4)
        assert m["primal"] == pytest.approx(p, rel=1e-3)
        torch.testing.assert_close(s.z, z_ref, rtol=1e-3, atol=1e-5)
    assert float(s.rho[0, 0]) != pytest.approx(0.1)  # the rule actually fired


def test_penalty_grad_and_value_match_autograd():
    torch.manual_seed(0)
    N = 500
    x = torch.randn(N, requires_grad=True)
    z, y, g = torch.randn(N), torch.randn(N), torch.randn(N)
    rho, l1, l2 = 0.3, 1e-2, 2e-2
    xd = x - z
    val = torch.dot(y, xd) + 0.5 * rho * torch.norm(xd, 2) ** 2 + l1 * torch.norm(x, 1) + l2 * torch.norm(x, 2) ** 2
    (gx,) = torch.autograd.grad(val, x)
    torch.testing.assert_close(flatops.penalty_grad(x.detach(), g, z, y, rho, l1, l2), g + gx)
    torch.testing.assert_close(flatops.penalty_value(x.detach(), z, y, rho, l1, l2), val.detach())


def test_adam_prox_step_matches_torch_adam():
    torch.manual_seed(0)
    N = 300
    x = torch.randn(N)
    p = torch.nn.Parameter(x.clone())
    opt = torch.optim.Adam([p], lr=1e-3)
    m, v = torch.zeros(N), torch.zeros(N)
    z, y = torch.randn(N), torch.randn(N)
    for t in range(1, 6):
        g = torch.randn(N)
        opt.zero_grad()
        loss = (p * g).sum() + torch.dot(y, p - z) + 0.25 * torch.norm(p - z) ** 2 + 1e-3 * torch.norm(p, 1) + 1e-3 * torch.norm(p, 2) ** 2
        loss.backward()
        opt.step()
        flatops.adam_prox_step(x, g, m, v, t, 1e-3, 0.9, 0.999, 1e-8, z, y, 0.5, 1e-3, 1e-3)
        torch.testing.assert_close(x, p.detach(), rtol=1e-5, atol=1e-6)
