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
import numpy as np
import torch


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
