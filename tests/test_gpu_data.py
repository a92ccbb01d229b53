"""DevicePrefetcher: batches arrive intact and in order while the ring of device buffers is reused
(replaces the synchronous uploads of search.py:212-220 / train.py:117-118)."""
import numpy as np
import pytest
import torch

from nas_3d_unet_b200.data import DevicePrefetcher

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [1, 2])
def test_prefetcher_order_and_reuse(depth):
    n = 7
    host = [(torch.full((2, 4, 32, 32, 32), float(i)).pin_memory(),
             np.full((2, 3, 32, 32, 32), i, dtype=np.int8)) for i in range(n)]
    pf = DevicePrefetcher(iter(host), "cuda", depth=depth)
    seen, ptrs = [], set()
    big = torch.randn(2048, 2048, device="cuda")
    for x, y in pf:
        assert x.is_cuda and x.dtype == torch.float32 and y.dtype == torch.int8
        ptrs.add(x.data_ptr())
        for _ in range(4):          # keep the compute stream busy so the next copy overlaps
            big = torch.tanh(big @ big * 1e-3)
        seen.append((float(x.sum()), int(y.to(torch.int64).sum())))
    assert seen == [(i * 2.0 * 4 * 32 ** 3, i * 2 * 3 * 32 ** 3) for i in range(n)]
    assert len(ptrs) == depth + 1           # a fixed ring, not one allocation per batch
    assert pf.h2d_bytes == n * (2 * 4 * 32 ** 3 * 4 + 2 * 3 * 32 ** 3)


def test_prefetcher_ragged_last_batch():
    host = [torch.ones(4, 8).pin_memory(), torch.ones(4, 8).pin_memory() * 2, torch.ones(1, 8) * 3]
    out = [float(b.sum()) for b in DevicePrefetcher(iter(host), "cuda")]
    assert out == [32.0, 64.0, 24.0]


def test_stage_batch_permutation_and_labels_bit_exact():
    """device staging = augment.permute_data + Generator.get_multi_class_labels, per sample
    (generator.py:195-248), written straight into the layout the stem reads"""
    import random
    from oracle import nas3d_oracle as O
    from nas_3d_unet_b200.data import permutation_keys, stage_batch
    rng = np.random.default_rng(5)
    N, P = 6, 16
    x = rng.standard_normal((N, 4, P, P, P)).astype(np.float32)
    seg = rng.choice(np.array([0, 1, 2, 4, 3], dtype=np.int16), size=(N, 1, P, P, P))
    keys = random.Random(3).sample(permutation_keys(), N - 1) + [None]
    for inclusive in (True, False):
        xo, yo = stage_batch(torch.as_tensor(x).cuda(), torch.as_tensor(seg).cuda(), keys, inclusive)
        assert xo.shape == (N, 4, P, P, P) and xo.is_contiguous(memory_format=torch.channels_last_3d)
        assert yo.dtype == torch.int8 and yo.shape == (N, 3, P, P, P)
        ref_x = np.stack([x[n] if k is None else O.permute_data(x[n], k) for n, k in enumerate(keys)])
        ref_s = np.stack([seg[n] if k is None else O.permute_data(seg[n], k) for n, k in enumerate(keys)])
        np.testing.assert_array_equal(xo.cpu().numpy(), ref_x)
        np.testing.assert_array_equal(yo.cpu().numpy(), O.multi_class_labels(ref_s, inclusive))
        # the same segmentation sent as int8 (1 byte per voxel over PCIe): identical masks
        _, yo8 = stage_batch(torch.as_tensor(x).cuda(), torch.as_tensor(seg.astype(np.int8)).cuda(), keys, inclusive)
        assert torch.equal(yo8, yo)
    # non-cubic patch, flips only, odd channel count, no labels
    x2 = rng.standard_normal((2, 3, 6, 10, 12)).astype(np.float32)
    k2 = [((0, 0), 1, 0, 1, 0), ((0, 0), 0, 1, 0, 0)]
    xo, yo = stage_batch(torch.as_tensor(x2).cuda(), None, k2)
    assert yo is None
    np.testing.assert_array_equal(xo.cpu().numpy(), np.stack([O.permute_data(x2[n], k2[n]) for n in range(2)]))
    with pytest.raises(TypeError):
        stage_batch(torch.as_tensor(x2), None, k2)       # CPU tensors: no fallback


def test_staged_batch_feeds_the_net_and_dice():
    """a staged batch (channels-last x, int8 masks) gives the same loss as the planar fp32 batch"""
    from oracle import nas3d_oracle as O
    from helpers import make_searched
    from nas_3d_unet_b200.data import stage_batch
    from nas_3d_unet_b200.loss import WeightedDiceLoss
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 4, 32, 32, 32)).astype(np.float32)
    seg = rng.choice(np.array([0, 1, 2, 4], dtype=np.int16), size=(2, 1, 32, 32, 32))
    model = make_searched().cuda()
    xg = torch.as_tensor(x).cuda()
    xo, yo = stage_batch(xg, torch.as_tensor(seg).cuda(), None, True)
    lossf = WeightedDiceLoss().cuda()
    with torch.no_grad():
        a = lossf(model(xo), yo)
        b = lossf(model(xg), torch.as_tensor(O.multi_class_labels(seg, True)).float().cuda())
    assert abs(float(a) - float(b)) <= 1e-6
