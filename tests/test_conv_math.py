"""CPU check of the stride-2 data-gradient decomposition (ops/conv_math.py) against autograd."""
import pytest
import torch
import torch.nn.functional as F

from federated_pytorch_test_b200.ops import conv_math


@pytest.mark.parametrize("k,pad", [(3, 1), (1, 0)])
@pytest.mark.parametrize("N,H,Ci,Co", [(2, 8, 4, 8), (3, 16, 8, 12), (1, 4, 16, 4)])
def test_stride2_dgrad_equals_autograd(k, pad, N, H, Ci, Co):
    g = torch.Generator().manual_seed(k * 100 + H + Ci)
    x = torch.randn(N, Ci, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Co, Ci, k, k, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, None, 2, pad)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    w_krsc = w.permute(0, 2, 3, 1).contiguous()
    wp = conv_math.pack_dgrad_s2_weight(w_krsc)
    assert wp.shape == (4 * Ci, 2, 2, Co)
    dx = conv_math.dgrad_s2(dy.permute(0, 2, 3, 1).contiguous(), wp, conv_math.conv2x2_oracle)
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), dx_ref, rtol=1e-10, atol=1e-10)


def test_packed_filter_has_nine_of_sixteen_taps():
    w = torch.ones(5, 3, 3, 7)
    wp = conv_math.pack_dgrad_s2_weight(w)
    assert int((wp != 0).sum()) == 9 * 5 * 7


@pytest.mark.parametrize("N,H,Ci,Co", [(2, 2, 96, 48), (3, 4, 8, 12), (1, 8, 12, 3), (2, 16, 4, 3)])
def test_conv_transpose_k4_s2_p1_as_one_conv_plus_pixel_shuffle(N, H, Ci, Co):
    g = torch.Generator().manual_seed(H + Ci + Co)
    x = torch.randn(N, Ci, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(Ci, Co, 4, 4, generator=g, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=1)
    wp = conv_math.pack_convT_s2_weight(w)
    assert wp.shape == (4 * Co, 3, 3, Ci)
    y = conv_math.convT_s2(x.permute(0, 2, 3, 1).contiguous(), wp, conv_math.conv3x3_oracle)
    torch.testing.assert_close(y.permute(0, 3, 1, 2), ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("N,H,Ci,Co,dils", [(2, 32, 8, 8, (1, 2, 4, 8, 16)), (3, 16, 4, 6, (1, 2, 4)), (1, 32, 4, 12, (1, 3))])
def test_multi_dilation_merged_filter_equals_concatenated_branches(N, H, Ci, Co, dils):
    g = torch.Generator().manual_seed(H + Ci + Co)
    x = torch.randn(N, Ci, H, H, generator=g, dtype=torch.float64)
    ws = [torch.randn(Co, Ci, 4, 4, generator=g, dtype=torch.float64) for _ in dils]
    pads = [(3 * d) // 2 for d in dils]
    ref = torch.cat([F.conv2d(x, w, None, 2, p, d) for w, p, d in zip(ws, pads, dils)], 1)
    wm = conv_math.pack_multidil_weight(ws)
    assert wm.shape == (len(dils) * Co, len(dils) * 4, 4, Ci)
    assert int((wm != 0).sum()) == len(dils) * Co * 16 * Ci                    # block diagonal: 1 / B of the merged filter
    y = conv_math.multidil_conv_oracle(x.permute(0, 2, 3, 1).contiguous(), wm, 4, 2, list(dils), pads)
    torch.testing.assert_close(y.permute(0, 3, 1, 2), ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("Ci,cw,k,stride,pad", [(8, 8, 4, 2, 1), (4, 8, 3, 1, 1), (12, 16, 4, 2, 1), (16, 16, 3, 1, 1), (4, 8, 1, 1, 0)])
def test_tap_packed_k_blocks_equal_the_convolution(Ci, cw, k, stride, pad):
    g = torch.Generator().manual_seed(Ci * 10 + k)
    x = torch.randn(2, Ci, 16, 16, generator=g, dtype=torch.float64)
    w = torch.randn(12, Ci, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, None, stride, pad)
    y = conv_math.tap_packed_gemm_oracle(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous(), stride, pad, cw)
    torch.testing.assert_close(y.permute(0, 3, 1, 2), ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("N,Ho,Wo,Ci", [(3, 16, 16, 64), (2, 8, 8, 128), (5, 4, 4, 32), (1, 2, 2, 64)])
def test_shuffle_store_addressing_equals_pixel_shuffle(N, Ho, Wo, Ci):
    g = torch.Generator().manual_seed(Ho + Ci)
    y = torch.randn(N, Ho, Wo, 4 * Ci, generator=g)
    ref = y.view(N, Ho, Wo, 2, 2, Ci).permute(0, 1, 3, 2, 4, 5).reshape(N, 2 * Ho, 2 * Wo, Ci)
    assert torch.equal(conv_math.shuffle_store_oracle(y), ref)
