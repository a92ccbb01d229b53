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
