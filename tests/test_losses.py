"""Vectorised losses vs loop oracles."""
import torch
import torch.nn as nn

from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import losses


def test_info_nce_matches_loop():
    torch.manual_seed(0)
    z, zh = torch.randn(4, 6, 3, 2, requires_grad=True), torch.randn(4, 6, 3, 2, requires_grad=True)
    a, b = losses.info_nce(z, zh), losses.info_nce_reference(z, zh)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    ga = torch.autograd.grad(a, (z, zh))
    gb = torch.autograd.grad(b, (z, zh))
    for u, v in zip(ga, gb):
        torch.testing.assert_close(u, v, rtol=1e-4, atol=1e-5)


def test_vae_loss():
    torch.manual_seed(0)
    r, x, mu, lv = torch.rand(3, 3, 8, 8), torch.rand(3, 3, 8, 8), torch.randn(3, 10), torch.randn(3, 10)
    ref = nn.MSELoss(reduction="sum")(r, x) - 0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp())
    torch.testing.assert_close(losses.vae_loss(r, x, mu, lv), ref)


def test_vae_cl_loss_matches_loop():
    torch.manual_seed(0)
    net = models.AutoEncoderCNNCL(K=3, L=4)
    x = torch.rand(5, 3, 32, 32)
    out = net(x)
    a = losses.vae_cl_loss(*out, x)
    b = losses.vae_cl_loss_reference(*out, x)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-3)
    c1, c2, c21, c3 = losses.vae_cl_costs(*out, x)
    assert c1.shape == (3,) and c21.shape == (3,)
