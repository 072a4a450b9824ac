"""Variational auto-encoders: ``AutoEncoderCNN`` and the variational-clustering
``AutoEncoderCNNCL`` (https://arxiv.org/abs/2005.04613).

Behavioural spec: /root/reference/src/simple_models.py:243-432.  Attribute names
and registration order are preserved (24 resp. 42 parameter tensors; block
tables of 12 two-tensor blocks resp. encoder/decoder/latent).

Differences by design:

* the noise for the reparametrisation trick is drawn on the *input's* device
  with ``torch.randn_like`` (the reference keys on ``torch.cuda.is_available()``
  — SURVEY Q14, a bug when a CUDA box runs a CPU model);
* ``AutoEncoderCNNCL.forward`` evaluates the shared convolutional encoder once
  and runs the K one-hot cluster passes as ONE batch of ``K*B`` rows through
  the dense layers and the decoder (M=1280 GEMMs instead of ten
  M=128 ones, and 1 instead of K+1 conv-encoder evaluations).  The results are
  returned in the reference's format (dicts indexed by cluster).  A
  ``batched_clusters=False`` switch restores the literal per-cluster loop.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import BlockPartitioned
from ..ops import functional as FX


def _down(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 4, stride=2, padding=1)


def _up(cin: int, cout: int) -> nn.ConvTranspose2d:
    return nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1)


class _ConvEncoderMixin:
    """3x32x32 -> 384 features through four stride-2 4x4 convs (12/24/48/96 ch)."""

    def _build_conv_encoder(self) -> None:
        chans = (3, 12, 24, 48, 96)
        for i in range(4):
            setattr(self, "conv%d" % (i + 1), _down(chans[i], chans[i + 1]))

    def conv_features(self, x: torch.Tensor) -> torch.Tensor:
        for i in range(1, 5):
            x = FX.conv_act(x, getattr(self, "conv%d" % i))
        return torch.flatten(x, start_dim=1)


class AutoEncoderCNN(_ConvEncoderMixin, BlockPartitioned):
    BLOCK_TABLE = ((0, 1), (2, 3), (4, 5), (6, 7), (8, 9), (14, 15), (16, 17), (18, 19), (20, 21), (22, 23), (10, 11), (12, 13))

    def __init__(self):
        super().__init__()
        self.latent_dim = 10
        self._build_conv_encoder()
        self.fc1 = nn.Linear(384, 16)
        self.fc21 = nn.Linear(16, self.latent_dim)
        self.fc22 = nn.Linear(16, self.latent_dim)
        self.fc3 = nn.Linear(self.latent_dim, 384)
        chans = (96, 48, 24, 12, 3)
        for i in range(4):
            setattr(self, "tconv%d" % (i + 1), _up(chans[i], chans[i + 1]))

    def encode(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = FX.linear_act(self.conv_features(x), self.fc1)
        return FX.linear_act(h, self.fc21, act=False), FX.linear_act(h, self.fc22, act=False)

    def reparametrize(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        std = torch.exp(0.5 * logvar)
        return mu + torch.randn_like(std) * std

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        h = FX.linear_act(z, self.fc3, act=False).reshape(-1, 96, 2, 2)
        for i in range(1, 5):
            h = FX.conv_act(h, getattr(self, "tconv%d" % i))
        return torch.sigmoid(h)

    def forward(self, x: torch.Tensor):
        mu, logvar = self.encode(x)
        return self.decode(self.reparametrize(mu, logvar)), mu, logvar


class AutoEncoderCNNCL(_ConvEncoderMixin, BlockPartitioned):
    # encoder, decoder, latent space (simple_models.py:430-432)
    BLOCK_TABLE = ((0, 7), (32, 41), (8, 31))

    def __init__(self, K: int = 10, L: int = 32, batched_clusters: bool = True):
        super().__init__()
        self.K = K
        self.L = L
        self.repr_flag = True
        self.batched_clusters = batched_clusters
        self._build_conv_encoder()
        # q(k|x)
        self.fc11 = nn.Linear(384, 128)
        self.fc12 = nn.Linear(128, 64)
        self.fc13 = nn.Linear(64, K)
        # q(z|x,k)
        self.fc21 = nn.Linear(384 + K, 128)
        self.fc22 = nn.Linear(128, 128)
        self.fc23 = nn.Linear(128, L)
        self.fc24 = nn.Linear(128, L)
        # p(z|k)
        self.fc14 = nn.Linear(K, 64)
        self.fc15 = nn.Linear(64, 64)
        self.fc16 = nn.Linear(64, L)
        self.fc17 = nn.Linear(64, L)
        # p(x|z)
        self.fc25 = nn.Linear(L, 384)
        self.tconv1 = _up(96, 48)
        self.tconv2 = _up(48, 24)
        self.tconv3 = _up(24, 12)
        self.tconv4 = _up(12, 3)
        self.tconv5 = _up(12, 3)

    # -- API kept from the reference ------------------------------------
    def enable_repr(self) -> None:
        self.repr_flag = True

    def disable_repr(self) -> None:
        # The reference sets the flag to True here as well; the
        # published behaviour is therefore "reparametrisation always on" and we
        # keep it.  Use ``force_disable_repr`` for the intended semantics.
        self.repr_flag = True

    def force_disable_repr(self) -> None:
        self.repr_flag = False

    # -- pieces -----------------------------------------------------------
    def _cluster_head(self, feats: torch.Tensor) -> torch.Tensor:
        h = FX.linear_act(feats, self.fc11)
        h = FX.linear_act(h, self.fc12)
        return F.softmax(FX.linear_act(h, self.fc13), dim=1)

    def encodeclus(self, x: torch.Tensor) -> torch.Tensor:
        return self._cluster_head(self.conv_features(x))

    def _latent_head(self, feats: torch.Tensor, ek: torch.Tensor):
        h = FX.linear_act(torch.cat((feats, ek), 1), self.fc21)
        h = FX.linear_act(h, self.fc22)
        return FX.linear_act(h, self.fc23), F.softplus(FX.linear_act(h, self.fc24))

    def encode(self, x: torch.Tensor, ek: torch.Tensor):
        return self._latent_head(self.conv_features(x), ek)

    def decode(self, ek: torch.Tensor, z: torch.Tensor):
        h = FX.linear_act(FX.linear_act(ek, self.fc14), self.fc15)
        mu_b, sig2_b = FX.linear_act(h, self.fc16, act=False), F.softplus(FX.linear_act(h, self.fc17, act=False))
        g = FX.linear_act(z, self.fc25).reshape(-1, 96, 2, 2)
        for name in ("tconv1", "tconv2", "tconv3"):
            g = FX.conv_act(g, getattr(self, name))
        mu_th = FX.conv_act(g, self.tconv4)
        sig2_th = F.softplus(FX.conv_act(g, self.tconv5))
        return mu_b, sig2_b, mu_th, sig2_th

    def reparametrize(self, mu: torch.Tensor, sig2: torch.Tensor) -> torch.Tensor:
        if not self.repr_flag:
            return mu
        std = sig2.sqrt()
        return mu + torch.randn_like(std) * std

    # -- forward ----------------------------------------------------------
    def forward(self, x: torch.Tensor):
        B = x.shape[0]
        feats = self.conv_features(x)
        ekhat = self._cluster_head(feats)
        if self.batched_clusters:
            eye = torch.eye(self.K, device=x.device, dtype=feats.dtype)
            ek = eye.repeat_interleave(B, dim=0)  # [K*B, K]; rows k*B..(k+1)*B-1 = e_k
            mu_xi, sig2_xi = self._latent_head(feats.repeat(self.K, 1), ek)
            z = self.reparametrize(mu_xi, sig2_xi)
            mu_b, sig2_b, mu_th, sig2_th = self.decode(ek, z)

            def split(t: torch.Tensor) -> Dict[int, torch.Tensor]:
                return {k: t[k * B:(k + 1) * B] for k in range(self.K)}

            return ekhat, split(mu_xi), split(sig2_xi), split(mu_b), split(sig2_b), split(mu_th), split(sig2_th)

        out = [dict() for _ in range(6)]
        for k in range(self.K):
            ek = torch.zeros(B, self.K, device=x.device, dtype=feats.dtype)
            ek[:, k] = 1
            m, s = self._latent_head(feats, ek)
            z = self.reparametrize(m, s)
            for d, v in zip(out, (m, s) + tuple(self.decode(ek, z))):
                d[k] = v
        return (ekhat,) + tuple(out)
