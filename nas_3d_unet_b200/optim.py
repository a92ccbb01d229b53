"""Flat-buffer Adam.

Drop-in for the `torch.optim.Adam` instances of the reference drivers (search.py:103-104 builds
`Adam(model.alphas())` and `Adam(model.kernel.parameters())`; train.py:58 `Adam(model.parameters())`):
same constructor arguments, `param_groups`, `zero_grad`, `state_dict` / `load_state_dict` layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter) and `ReduceLROnPlateau` compatibility, but the update
of a whole param group is ONE launch of `nas3d_adam_flat_step` over three flat fp32 arenas instead of
a multi-tensor loop over 1 784 tensors.

On the first `step()` every parameter of a group is moved into a flat arena (`p.data` becomes a view
of it; values unchanged), so do `model.to(device)` before the first step, as the drivers do.
Gradients are NOT flattened: the kernel reaches them through a table of pointers passed as kernel
parameters (nothing is uploaded; a captured CUDA graph keeps them as node constants).

lr / betas / eps / weight_decay and the step counter live in a device array, so a step captured in a
CUDA graph (graph.GraphedStep) keeps counting and sees scheduler updates after `sync()`.
Fails loudly (RuntimeError) for CPU tensors - there is no CPU fallback.

Deviation from torch.optim.Adam, by design: the step counter is ONE value per param group (advanced
by every `step()`), exported as every parameter's `step` in `state_dict()`; torch counts per
parameter and skips parameters whose grad is None.  The reference loops give every parameter of a
group a gradient on every step (search.py:222-238, train.py:121-128), where the two are identical; a
torch Adam checkpoint with differing per-parameter steps loads with the group's maximum, and a
parameter that is frozen for some steps gets the group's (larger) bias correction afterwards.
"""
import ctypes as C

import torch

from ._lib import check, int_array, load, ptr_array


def chunk_table(numels, chunk):
    """[(tensor, flat offset, offset in tensor, length)], tensor offsets rounded up to 4 floats;
    returns (rows, total floats, per-tensor flat offsets)"""
    rows, offs, off = [], [], 0
    for t, n in enumerate(numels):
        offs.append(off)
        for c in range(0, n, chunk):
            rows.append((t, off + c, c, min(chunk, n - c)))
        off += (n + 3) // 4 * 4
    return rows, off, offs


class _Group:
    __slots__ = ("params", "offs", "arena", "exp_avg", "exp_avg_sq", "chunks", "first_chunk", "hyper",
                 "hyper_host")


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                 maximize=False):
        if amsgrad:
            raise NotImplementedError("FlatAdam: amsgrad is not used by the reference and not built")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FlatAdam: invalid hyper-parameters")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                        maximize=maximize)
        super().__init__(params, defaults)
        self._groups = [None] * len(self.param_groups)
        self._lib = None

    # ---- arenas -----------------------------------------------------------------------------
    def _build(self, gi):
        group = self.param_groups[gi]
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return None
        dev = ps[0].device
        for p in ps:
            if p.device.type != "cuda":
                raise RuntimeError("FlatAdam needs CUDA parameters (got %s); there is no CPU path"
                                   % p.device)
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("FlatAdam: one device and float32 per param group")
        if self._lib is None:
            self._lib = load()
        chunk = self._lib.nas3d_adam_chunk_floats()
        rows, total, offs = chunk_table([p.numel() for p in ps], chunk)
        if total >= 2 ** 31:
            raise RuntimeError("FlatAdam: param group too large for 32-bit offsets")
        g = _Group()
        g.params, g.offs = ps, offs
        g.arena = torch.zeros(max(total, 4), device=dev, dtype=torch.float32)
        g.exp_avg = torch.zeros_like(g.arena)
        g.exp_avg_sq = torch.zeros_like(g.arena)
        with torch.no_grad():
            for p, off in zip(ps, offs):
                view = g.arena[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        g.chunks = torch.tensor(rows, dtype=torch.int32).reshape(-1, 4).to(dev)
        first = [0] * (len(ps) + 1)
        for t, _, _, _ in rows:
            first[t + 1] += 1
        for t in range(len(ps)):
            first[t + 1] += first[t]
        g.first_chunk = int_array(first)
        g.hyper = torch.zeros(16, device=dev, dtype=torch.float32)
        g.hyper_host = None
        self._groups[gi] = g
        self._sync_group(gi)
        return g

    def _sync_group(self, gi):
        g, group = self._groups[gi], self.param_groups[gi]
        if g is None:
            return
        lr = group["lr"]
        lr = float(lr.item()) if torch.is_tensor(lr) else float(lr)
        key = (lr, float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
               float(group["weight_decay"]), 1.0 if group.get("maximize", False) else 0.0)
        if key == g.hyper_host:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("FlatAdam: hyper-parameters changed while capturing a CUDA graph")
        g.hyper[0:5].copy_(torch.tensor(key[:5], dtype=torch.float32))
        g.hyper[8:11].copy_(torch.tensor([key[5], 1.0 - key[1], 1.0 - key[2]], dtype=torch.float32))
        g.hyper_host = key

    def sync(self):
        """push host-side changes of lr/betas/eps/weight_decay (e.g. ReduceLROnPlateau.step) to the
        device copy; step() does this itself, a replayed CUDA graph needs it called before replay"""
        for gi in range(len(self.param_groups)):
            self._sync_group(gi)

    def _grad_table(self, g):
        ptrs = []
        for p in g.params:
            gr = p.grad
            if gr is None:
                ptrs.append(0)
                continue
            if gr.is_sparse:
                raise RuntimeError("FlatAdam does not support sparse gradients")
            if gr.dtype != torch.float32 or gr.device != p.device:
                raise RuntimeError("FlatAdam: gradients must be float32 on the parameter's device")
            if not gr.is_contiguous():
                gr = p.grad = gr.contiguous()
            ptrs.append(gr.data_ptr())
        return ptr_array(ptrs)

    # ---- torch.optim.Optimizer surface --------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi in range(len(self.param_groups)):
            g = self._groups[gi] or self._build(gi)
            if g is None:
                continue
            if not torch.cuda.is_current_stream_capturing():
                self._sync_group(gi)
            tab = self._grad_table(g)
            with torch.cuda.device(g.arena.device):
                st = torch.cuda.current_stream().cuda_stream
                check(self._lib.nas3d_adam_flat_step(
                    g.arena.data_ptr(), g.exp_avg.data_ptr(), g.exp_avg_sq.data_ptr(),
                    tab, len(g.params), g.first_chunk, g.chunks.data_ptr(), g.hyper.data_ptr(), st),
                    "adam_flat_step")
        return loss

    def _export_state(self):
        for g in self._groups:
            if g is None:
                continue
            step = g.hyper[5].detach().cpu().clone()
            for p, off in zip(g.params, g.offs):
                n = p.numel()
                self.state[p] = {"step": step.clone(),
                                 "exp_avg": g.exp_avg[off:off + n].view(p.shape),
                                 "exp_avg_sq": g.exp_avg_sq[off:off + n].view(p.shape)}

    def state_dict(self):
        """same layout as torch.optim.Adam.state_dict() (search.py:171-172 pickles it)"""
        self._export_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for gi in range(len(self.param_groups)):
                g = self._groups[gi] or self._build(gi)
                if g is None:
                    continue
                step = None
                for p, off in zip(g.params, g.offs):
                    st = self.state.get(p)
                    if not st:
                        continue
                    n = p.numel()
                    g.exp_avg[off:off + n].view(p.shape).copy_(st["exp_avg"])
                    g.exp_avg_sq[off:off + n].view(p.shape).copy_(st["exp_avg_sq"])
                    s = float(st["step"])
                    step = s if step is None else max(step, s)
                if step is not None:
                    g.hyper[5:6].copy_(torch.tensor([step], dtype=torch.float32))
                g.hyper_host = None
                self._sync_group(gi)
        self._export_state()


Adam = FlatAdam
