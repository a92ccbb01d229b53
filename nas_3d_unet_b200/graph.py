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
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, example_inputs, warmup=3, optimizers=()):
        """step_fn(*tensors) -> tensor or tuple of tensors; it must do the COMPLETE step
        (zero_grad ... optimizer.step).  `warmup` eager steps run first (they are real steps).
        `optimizers`: optim.FlatAdam instances whose lr etc. are re-read before every replay, so
        ReduceLROnPlateau keeps working on a captured step."""
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
