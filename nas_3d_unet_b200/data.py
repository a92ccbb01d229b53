"""Host -> device staging for the training / search loops.

The reference uploads every batch synchronously right before using it
(`torch.as_tensor(x, device=self.device, dtype=torch.float)`, search.py:212-220, train.py:117-118).
`DevicePrefetcher` wraps any iterator of host batches (tuples of pinned CPU tensors or numpy
arrays) and yields device tensors while the NEXT batch is already in flight on a copy stream, so
the PCIe transfer overlaps the kernels of the current step.  The device side is a ring of
`depth + 1` preallocated buffer sets that are reused for the whole epoch (no allocator traffic and
no cudaMalloc in the loop); a buffer set is overwritten only after the step that read it has been
enqueued, which the copy stream waits on through an event.  A yielded batch is therefore valid
until the next batch is requested - exactly how the reference loops use it.  Usage with the
unchanged step loop:

    for x, y in DevicePrefetcher(train_generator.epoch(), device):
        optim.zero_grad(); loss = lossf(model(x), y); loss.backward(); optim.step()
"""
import itertools
import random

import numpy as np
import torch

from . import _lib


def permutation_keys():
    """the 48 cube symmetries as the reference names them (augment.py:78-93):
    ((rotate_y, rotate_z), flip_x, flip_y, flip_z, transpose)"""
    rots = list(itertools.combinations_with_replacement(range(2), 2))
    return sorted(itertools.product(rots, range(2), range(2), range(2), range(2)))


def random_permutation_key(rng=random):
    """augment.random_permutation_key (augment.py:96-101)"""
    return rng.choice(permutation_keys())


def index_map(key, shape):
    """(base, stride_d, stride_h, stride_w): output voxel (d,h,w) of augment.permute_data(data, key)
    (augment.py:105-132) is source voxel base + stride_d*d + stride_h*h + stride_w*w.  Derived by
    pushing a cube of linear indices through the same numpy view operations (host, 3 ints/axis)."""
    D, H, W = (int(s) for s in shape)
    if key is None:
        return (0, H * W, W, 1)
    (rot_y, rot_z), flip_x, flip_y, flip_z, transpose = key
    if (rot_y or rot_z or transpose) and not D == H == W:
        raise AssertionError("Not a cubic patch!")       # generator.py:213
    idx = np.arange(D * H * W, dtype=np.int64).reshape(1, D, H, W)
    # strided views only - nothing the size of the patch is copied
    if rot_y:
        idx = np.rot90(idx, rot_y, axes=(1, 3))
    if rot_z:
        idx = np.rot90(idx, rot_z, axes=(2, 3))
    if flip_x:
        idx = idx[:, ::-1]
    if flip_y:
        idx = idx[:, :, ::-1]
    if flip_z:
        idx = idx[:, :, :, ::-1]
    v = idx[0].T if transpose else idx[0]
    base = int(v[0, 0, 0])
    step = lambda ax: int(v[tuple(1 if a == ax else 0 for a in range(3))]) - base if v.shape[ax] > 1 else 0
    return (base, step(0), step(1), step(2))


def stage_batch(x, seg=None, keys=None, inclusive_label=True):
    """device-side batch assembly: x (N,C,D,H,W) float32 CUDA (planar, as uploaded), seg (N,1,D,H,W)
    int16 (as the h5 files store it) or int8 (the labels 0, 1, 2, 4 fit: 1 byte per voxel over PCIe)
    CUDA or None, keys = one permutation key per sample (or None = no permutation).
    Returns (x in channels_last_3d memory - what the stem reads without a layout pass,
             y (N,3,D,H,W) int8 region masks or None)  [generator.py:195-248]"""
    if not x.is_cuda or x.dtype != torch.float32:
        raise TypeError("stage_batch expects a CUDA float32 batch (no CPU path)")
    x = x.contiguous()
    N, C, D, H, W = x.shape
    if seg is not None:
        if not seg.is_cuda or seg.dtype not in (torch.int16, torch.int8) or seg.numel() != N * D * H * W:
            raise TypeError("stage_batch: seg must be CUDA int16 or int8 of shape (N,1,D,H,W)")
        seg = seg.contiguous()
    if keys is None:
        keys = [None] * N
    if len(keys) != N:
        raise ValueError("stage_batch: one permutation key per sample")
    maps = []
    for k in keys:
        maps += index_map(k, (D, H, W))
    ld = (C + 3) // 4 * 4
    xo = torch.empty((N, D, H, W, ld), device=x.device, dtype=torch.float32)
    yo = torch.empty((N, 3, D, H, W), device=x.device, dtype=torch.int8) if seg is not None else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        fn = lib.nas3d_stage_patches_seg8 if (seg is not None and seg.dtype == torch.int8) else lib.nas3d_stage_patches
        _lib.check(fn(
            x.data_ptr(), seg.data_ptr() if seg is not None else None, N, C, D, H, W,
            _lib.int_array(maps), 1 if inclusive_label else 0, xo.data_ptr(), ld,
            yo.data_ptr() if yo is not None else None, st), "stage_patches")
    return xo[..., :C].permute(0, 4, 1, 2, 3), yo


class DevicePrefetcher:
    def __init__(self, batches, device, dtype=torch.float32, depth=1):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.dtype = dtype
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, int(depth))
        self.queue = []
        self.h2d_bytes = 0
        self.ring = [(None, None)] * (self.depth + 1)   # (device buffers, "consumed" event)
        self.next_slot = 0
        self.live_slot = None
        for _ in range(self.depth):
            self._issue()

    def _to_host_tensor(self, a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(a)
        if a.dtype != self.dtype and a.is_floating_point():
            a = a.to(self.dtype)
        if not a.is_pinned():
            a = a.pin_memory()
        return a

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return
        if not isinstance(batch, (tuple, list)):
            batch = (batch,)
        host = [self._to_host_tensor(a) for a in batch]
        slot = self.next_slot
        self.next_slot = (slot + 1) % len(self.ring)
        bufs, consumed = self.ring[slot]
        if bufs is None or len(bufs) != len(host) or any(
                b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host]
        with torch.cuda.stream(self.copy_stream):
            if consumed is not None:
                self.copy_stream.wait_event(consumed)   # the step that read this slot is enqueued
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ring[slot] = (bufs, None)
        self.h2d_bytes += sum(h.numel() * h.element_size() for h in host)
        self.queue.append((bufs, ev, host, slot))

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.live_slot is not None:
            # everything that reads the previously yielded batch has been enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self.ring[self.live_slot] = (self.ring[self.live_slot][0], ev)
            self.live_slot = None
        if not self.queue:
            raise StopIteration
        dev, ev, _host, slot = self.queue.pop(0)
        cur.wait_event(ev)
        self.live_slot = slot
        self._issue()
        out = [t if t.dtype == self.dtype or not t.is_floating_point() else t.to(self.dtype) for t in dev]
        return tuple(out) if len(out) > 1 else out[0]
