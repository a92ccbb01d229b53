"""Host -> device staging for the training / search loops.

The reference uploads every batch synchronously right before using it
(`torch.as_tensor(x, device=self.device, dtype=torch.float)`, search.py:212-220, train.py:117-118).
`DevicePrefetcher` wraps any iterator of host batches (tuples of pinned CPU tensors or numpy
arrays) and yields device tensors while the NEXT batch is already in flight on a copy stream, so
the PCIe transfer overlaps the kernels of the current step.  Usage with the unchanged step loop:

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
        with torch.cuda.stream(self.copy_stream):
            dev = [h.to(self.device, non_blocking=True) for h in host]
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.h2d_bytes += sum(h.numel() * h.element_size() for h in host)
        self.queue.append((dev, ev, host))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        dev, ev, _host = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)      # the caching allocator must not recycle it under the step
        self._issue()
        out = [t if t.dtype == self.dtype or not t.is_floating_point() else t.to(self.dtype) for t in dev]
        return tuple(out) if len(out) > 1 else out[0]
