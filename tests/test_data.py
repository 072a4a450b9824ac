"""Data layer: reference shard arithmetic, normalisation, loaders, native batch assembler, LOFAR patches."""
import pytest
import torch

from federated_pytorch_test_b200.data import (CifarData, LofarSource, ShardLoader, get_data_minibatch,
                                              make_synthetic_cifar, normalize_batch, shard_ranges, worker_norm)


def test_shard_sizes_match_reference():
    assert [len(r) for r in shard_ranges(8)] == [6249] * 8       # SURVEY §2.7 (off-by-one preserved)
    assert [len(r) for r in shard_ranges(10)] == [4999] * 10
    assert [len(r) for r in shard_ranges(1)] == [49999]
    assert [len(r) for r in shard_ranges(8, drop_last_sample=False)] == [6250] * 8
    r = shard_ranges(3)
    assert r[0][0] == 0 and r[2][-1] <= 49999
    assert -(-6249 // 128) == 49 and -(-4999 // 128) == 40 and -(-49999 // 128) == 391


def test_worker_norm_and_normalize():
    mean, std = worker_norm(3, True)
    assert mean == (0.53, 0.47, 0.5) and std == mean
    assert worker_norm(3, False)[0] == (0.5, 0.5, 0.5)
    u8 = torch.randint(0, 256, (4, 32, 32, 3), dtype=torch.uint8)
    x = normalize_batch(u8, mean, std)
    ref = (u8.float().permute(0, 3, 1, 2) / 255 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    torch.testing.assert_close(x, ref)
    xc = normalize_batch(u8, mean, std, channels_last=True)
    assert xc.shape == (4, 3, 32, 32) and xc.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(xc.contiguous(), ref)


def test_synthetic_is_deterministic_and_learnable():
    a = make_synthetic_cifar(True, seed=7, size=512)
    b = make_synthetic_cifar(True, seed=7, size=512)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert a[0].dtype == torch.uint8 and a[0].shape == (512, 32, 32, 3)
    # nearest-class-mean on raw pixels must beat chance by a wide margin
    x, y = a[0].float().flatten(1), a[1]
    means = torch.stack([x[y == c].mean(0) for c in range(10)])
    te = make_synthetic_cifar(False, seed=7, size=512)
    pred = torch.cdist(te[0].float().flatten(1), means).argmin(1)
    assert (pred == te[1]).float().mean() > 0.5


def test_shard_loader_epoch_covers_shard_once():
    imgs, labs = make_synthetic_cifar(True, seed=1, size=1000)
    mean, std = worker_norm(0)
    ld = ShardLoader(imgs, labs, range(100, 400), 128, torch.device("cpu"), mean, std, seed=3)
    assert len(ld) == 3
    seen = []
    for x, y in ld:
        assert x.shape[1:] == (3, 32, 32) and x.dtype == torch.float32
        seen.append(y)
    assert sum(t.numel() for t in seen) == 300 and seen[-1].numel() == 300 - 256
    first = torch.cat([y for _, y in ld])
    assert first.numel() == 300


def test_native_batch_assembler_matches_gather():
    from federated_pytorch_test_b200.runtime.batch_loader import BatchAssembler
    imgs, labs = make_synthetic_cifar(True, seed=2, size=700)
    asm = BatchAssembler(imgs, labs, 64, slots=3, threads=2)
    for epoch in range(2):
        order = torch.randperm(700)[:650]
        asm.start_epoch(order)
        for b in range(-(-650 // 64)):
            u8, lab = asm.next_batch_to(torch.device("cpu"))
            idx = order[b * 64:(b + 1) * 64]
            assert torch.equal(u8, imgs[idx]) and torch.equal(lab, labs[idx])
        assert asm.next_batch() is None
    asm.close()


def test_lofar_minibatch_contract():
    src = LofarSource.synthetic(nbase=10, ntime=64, nfreq=48, seed=0)
    g = torch.Generator().manual_seed(0)
    px, py, y = get_data_minibatch(src, batch_size=3, patch_size=32, generator=g)
    assert (px, py) == (3, 2) and y.shape == (3 * 6, 8, 32, 32)
    g = torch.Generator().manual_seed(0)
    _, _, yr = get_data_minibatch(src, batch_size=3, patch_size=32, generator=g, layout="reference")
    # reference layout is patch-major: row p*B+b  <->  batch-major row b*P+p
    P = 6
    for b in range(3):
        for p in range(P):
            assert torch.equal(y[b * P + p], yr[p * 3 + b])
    # channel = 2*pol + (re, im), scaled by the per-(baseline,freq,pol) factor
    g = torch.Generator().manual_seed(0)
    pick = torch.randint(0, 10, (3,), generator=g)
    v = src.visibilities[pick[0], :32, :32, 1, 1].float() * src.scale[pick[0], :32, 1].view(1, 32)
    torch.testing.assert_close(y[0, 3], v)
