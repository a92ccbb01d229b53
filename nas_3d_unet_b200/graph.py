"""CUDA-graph capture of a whole training / search step.

The step loops of the reference (search.py:222-238, train.py:121-128) launch thousands of small
kernels per step (5 900 for a supernet search step); at batch 1 the host cannot issue them as fast
as a B200 retires them.  `GraphedStep` records one full step - forward, Dice loss, backward, the
NCCL gradient all-reduce if data parallelism is enabled, and the optimizer update - into a CUDA
graph once and replays it with a single launch per step.  Everything the step touches stays at
fixed addresses: parameters are updated in place (optim.FlatAdam, or a torch optimizer built with
`capturable=True`),
activations come from the graph's private memory pool, inputs are copied into static buffers.

    step = GraphedStep(lambda x, y: train_step(model, lossf, optim, x, y), (x0, y0))
    for x, y in batches:
        loss = step(x, y)          # device scalar; loss.item() when needed

Host-fed loops use `stream()`: with `buffers=2` the step is captured twice over two static input
sets, the pinned-host batch i+1 is copied STRAIGHT into
the idle set on a copy stream while graph i runs (no device-to-device staging copy), and the loss
of step i is read back one step late from a pinned slot, so the host never drains the GPU queue:

    for loss_value in step.stream(host_batches):      # python floats, in step order
        ...
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, example_inputs, warmup=3, optimizers=(), buffers=1, share_pool=False):
        """step_fn(*tensors) -> tensor or tuple of tensors; it must do the COMPLETE step
        (zero_grad ... optimizer.step).  `warmup` eager steps run first (they are real steps).
        `optimizers`: optim.FlatAdam instances whose lr etc. are re-read before every replay, so
        ReduceLROnPlateau keeps working on a captured step.  `buffers` = 2 captures a second graph
        over a second set of static inputs for stream(); each graph gets its own memory pool unless
        `share_pool` (sharing is only valid if the graphs are always replayed alternately).
        Measured limit on B200 / driver 580: two instantiated graphs of a supernet 128^3 search step
        (~12 000 kernel nodes each) segfault inside cudaGraphLaunch, with shared or separate pools;
        one such graph, or two searched-net graphs (~530 nodes each), are fine."""
        self.optimizers = [o for o in optimizers if hasattr(o, "sync")]
        try:   # warm-up runs on a side stream, which torch would flag for every AccumulateGrad node
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        self.static_in = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step_fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)
        self.sets = [(self.static_in, self.graph, self.static_out)]
        for _ in range(1, max(1, int(buffers))):
            ins = [t.detach().clone() for t in example_inputs]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=(self.graph.pool() if share_pool else None)):
                out = step_fn(*ins)
            self.sets.append((ins, g, out))
        self._copy_stream = None

    def load(self, *inputs):
        """copy a batch (host-pinned or device tensors) into the static input buffers"""
        for s, t in zip(self.static_in, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)

    def replay(self):
        for o in self.optimizers:
            o.sync()
        self.graph.replay()
        return self.static_out

    def __call__(self, *inputs):
        self.load(*inputs)
        return self.replay()

    def stream(self, host_batches, scalar=lambda out: out.reshape(-1)[-1], h2d_delay_s=None):
        """run every batch of `host_batches` (tuples of pinned host tensors shaped like the example
        inputs) through the captured step; yields `scalar(step output)` of each step as a python
        float, in order, each value one step late (after the NEXT step has been enqueued).
        Per step the timeline holds: one H2D of the batch into the idle static input set (copy
        stream, overlapping the running graph), one graph launch, one 4-byte D2H.

        With two input sets the prefetch is PACED: the copy of batch i+1 is issued h2d_delay_s after
        graph i started instead of together with it, so that the DMA writes land in the step's
        FMA-/latency-bound middle and end rather than in its HBM-bound first third (measured on B200,
        searched net 8 x 128^3: the copy engine's 285 MB slow the streaming kernels down
        disproportionately - e2e 457.9 patches/s unpaced, 478.0 with 9 ms; profiles/r12_h2d_delay.txt).
        h2d_delay_s = None (default) measures the graph and the copy with CUDA events and uses
        graph - copy - 1 ms, i.e. the copy ends just before the next graph needs it; 0 when the
        copy is the longer of the two (many GPUs behind one host).  0.0 = unpaced."""
        import time
        auto = h2d_delay_s is None or h2d_delay_s < 0
        delay = 0.0 if auto else float(h2d_delay_s)
        dev = self.static_in[0].device
        main = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._pinned = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in self.sets]
        cs = self._copy_stream
        nset = len(self.sets)
        consumed = [None] * nset          # main-stream event: graph k has read its inputs
        it = iter(host_batches)

        copy_events = [None]              # (start, end) of the most recent H2D (auto pacing)

        def issue(k, batch):
            ins = self.sets[k][0]
            with torch.cuda.stream(cs):
                if consumed[k] is not None:
                    cs.wait_event(consumed[k])
                c0 = None
                if auto:
                    c0 = torch.cuda.Event(enable_timing=True)
                    c0.record(cs)
                for s, h in zip(ins, batch):
                    s.copy_(h, non_blocking=True)
                ev = torch.cuda.Event(enable_timing=auto)
                ev.record(cs)
                if auto:
                    copy_events[0] = (c0, ev)
            return ev

        try:
            nxt = next(it)
        except StopIteration:
            return
        ready = issue(0, nxt)
        pending = None                    # (event, pinned slot) of the previous step's scalar
        prev_graph_events = None
        i = 0
        while ready is not None:
            k = i % nset
            cur_ready = ready
            ready = None
            paced = nset > 1 and (auto or delay > 0)
            if nset > 1 and not paced:    # prefetch batch i+1 into the other set while graph k runs
                try:
                    ready = issue((i + 1) % nset, next(it))
                except StopIteration:
                    pass
            main.wait_event(cur_ready)
            for o in self.optimizers:
                o.sync()
            g0 = None
            if auto:
                g0 = torch.cuda.Event(enable_timing=True)
                g0.record(main)
            self.sets[k][1].replay()
            ev = torch.cuda.Event(enable_timing=auto)
            ev.record(main)
            consumed[k] = ev
            graph_events = (g0, ev)
            slot = self._pinned[k]
            slot.copy_(scalar(self.sets[k][2]).detach().reshape(1), non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            if nset == 1:                 # single set: the next H2D has to wait for this replay
                try:
                    ready = issue(0, next(it))
                except StopIteration:
                    pass
            prev = None
            if pending is not None:
                pending[0].synchronize()          # step i-1 is done: graph i starts about now
                prev = float(pending[1])
                if auto and prev_graph_events is not None and copy_events[0] is not None:
                    # graph i-1 has finished; the copy of batch i was issued during it
                    c0, c1 = copy_events[0]
                    if c1.query():
                        g_ms = prev_graph_events[0].elapsed_time(prev_graph_events[1])
                        c_ms = c0.elapsed_time(c1)
                        delay = max(0.0, min(g_ms - c_ms - 1.0, 0.7 * g_ms)) * 1e-3
                    else:
                        delay = 0.0               # the graph is waiting for the copy: do not hold it back
            if auto:
                prev_graph_events = graph_events
            if paced:
                t0 = time.perf_counter()
                try:
                    nxt = next(it)
                    while time.perf_counter() - t0 < delay:
                        pass
                    ready = issue((i + 1) % nset, nxt)
                except StopIteration:
                    pass
            if prev is not None:
                yield prev
            # the slot is overwritten nset steps from now; its value is read (above) before that
            pending = (done, slot) if nset > 1 else None
            if nset == 1:
                done.synchronize()
                yield float(slot)
            i += 1
        if pending is not None:
            pending[0].synchronize()
            yield float(pending[1])
