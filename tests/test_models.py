"""Model zoo: shapes, parameter/block inventory and forward parity with the reference."""
import pytest
import torch

from federated_pytorch_test_b200 import models
from federated_pytorch_test_b200.ops import functional as FX

SPEC = [  # (factory, #tensors, #params)
    (models.Net, 10, 62006), (models.Net1, 12, 890410), (models.Net2, 18, 2513418),
    (models.ResNet18, 62, 11173962), (models.ResNet9, 38, 4903242),
    (models.AutoEncoderCNN, 24, 205679), (lambda: models.AutoEncoderCNNCL(10, 32), 42, 350744),
    (lambda: models.EncoderCNN(256), 16, 701928), (lambda: models.ContextgenCNN(256), 4, 98304),
    (lambda: models.PredictorCNN(256, 32), 2, 16384),
]


@pytest.mark.parametrize("factory,ntensors,nparams", SPEC)
def test_inventory(factory, ntensors, nparams):
    net = factory()
    ps = list(net.parameters())
    assert len(ps) == ntensors
    assert sum(p.numel() for p in ps) == nparams
    covered = []
    for lo, hi in net.train_order_block_ids():
        covered += list(range(lo, hi + 1))
    assert sorted(covered) == list(range(ntensors)), "blocks must cover every parameter index exactly once"


def test_resnet18_block_sizes():
    net = models.ResNet18()
    sizes = [net.block_numel(i) for i in range(10)]
    assert sizes == [1856, 73984, 73984, 230144, 295424, 919040, 1180672, 3673088, 4720640, 5130]
    assert [models.Net().block_numel(i) for i in range(5)] == [48120, 456, 2416, 10164, 850]


def test_protocol_methods():
    net = models.Net()
    assert net.linear_layer_ids() == [4, 6, 8]
    assert net.train_order_block_ids() == [[4, 5], [0, 1], [2, 3], [6, 7], [8, 9]]
    assert net.linear_layer_parameters().numel() == 48120 + 10164 + 850
    assert models.ResNet18().linear_layer_ids() == []


PAIRS = [("Net", lambda m: m.Net(), (4, 3, 32, 32)), ("Net1", lambda m: m.Net1(), (4, 3, 32, 32)),
         ("Net2", lambda m: m.Net2(), (4, 3, 32, 32)), ("ResNet18", lambda m: m.ResNet18(), (4, 3, 32, 32)),
         ("ResNet9", lambda m: m.ResNet9(), (4, 3, 32, 32)),
         ("ContextgenCNN", lambda m: m.ContextgenCNN(32), (2, 32, 3, 3))]


@pytest.mark.parametrize("name,make,shape", PAIRS)
def test_forward_matches_reference(ref_models, name, make, shape):
    """Same state_dict keys, and identical outputs when given the reference's weights."""
    FX.set_fast_path(False)
    torch.manual_seed(0)
    ref = make(ref_models)
    mine = make(models)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(*shape)
    torch.testing.assert_close(mine(x), ref(x), rtol=1e-5, atol=1e-5)


def test_encoder_predictor_match_reference(ref_models):
    torch.manual_seed(0)
    ref, mine = ref_models.EncoderCNN(64), models.EncoderCNN(64)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(6, 8, 32, 32)
    torch.testing.assert_close(mine(x), ref(x), rtol=1e-5, atol=1e-5)
    rp, mp = ref_models.PredictorCNN(64, 8), models.PredictorCNN(64, 8)
    mp.load_state_dict(rp.state_dict())
    a, b = torch.randn(2, 64, 3, 3), torch.randn(2, 64, 3, 3)
    for u, v in zip(mp(a, b), rp(a, b)):
        torch.testing.assert_close(u, v)


def test_vae_deterministic_parts_match_reference(ref_models):
    torch.manual_seed(0)
    ref, mine = ref_models.AutoEncoderCNN(), models.AutoEncoderCNN()
    mine.load_state_dict(ref.state_dict())
    x = torch.rand(3, 3, 32, 32)
    for u, v in zip(mine.encode(x), ref.encode(x)):
        torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-6)
    z = torch.randn(3, 10)
    torch.testing.assert_close(mine.decode(z), ref.decode(z), rtol=1e-5, atol=1e-6)


def test_vae_cl_batched_equals_loop(ref_models):
    torch.manual_seed(0)
    ref = ref_models.AutoEncoderCNNCL(K=4, L=8)
    a = models.AutoEncoderCNNCL(K=4, L=8, batched_clusters=True)
    b = models.AutoEncoderCNNCL(K=4, L=8, batched_clusters=False)
    a.load_state_dict(ref.state_dict())
    b.load_state_dict(ref.state_dict())
    a.force_disable_repr()
    b.force_disable_repr()
    x = torch.rand(5, 3, 32, 32)
    torch.testing.assert_close(a.encodeclus(x), ref.encodeclus(x), rtol=1e-5, atol=1e-6)
    oa, ob = a(x), b(x)
    torch.testing.assert_close(oa[0], ob[0])
    for da, db in zip(oa[1:], ob[1:]):
        for k in range(4):
            torch.testing.assert_close(da[k], db[k], rtol=1e-5, atol=1e-6)
    # deterministic heads against the reference, cluster by cluster
    ek = torch.zeros(5, 4)
    ek[:, 2] = 1
    for u, v in zip(a.encode(x, ek), ref.encode(x, ek)):
        torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-6)
    z = torch.randn(5, 8)
    for u, v in zip(a.decode(ek, z), ref.decode(ek, z)):
        torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-6)


def test_disable_repr_quirk_preserved():
    net = models.AutoEncoderCNNCL(2, 4)
    net.disable_repr()
    assert net.repr_flag is True
    net.force_disable_repr()
    assert net.repr_flag is False
