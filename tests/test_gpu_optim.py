"""FlatAdam against torch.optim.Adam run on the CPU on the same parameters and
gradients - the update the reference drivers perform (search.py:228,237; train.py:127)."""
import copy

import pytest
import torch

from nas_3d_unet_b200.optim import FlatAdam, chunk_table

pytestmark = pytest.mark.gpu

SHAPES = [(4, 4, 3, 3, 3), (5,), (12, 4, 1, 1, 1), (1,), (3, 7), (64, 64, 3, 3, 3), (9, 6)]


def _make(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in SHAPES]


def _grads(step, skip=None, seed=100):
    g = torch.Generator().manual_seed(seed + step)
    return [None if i == skip else torch.randn(s, generator=g) * (10.0 ** (i % 3 - 1))
            for i, s in enumerate(SHAPES)]


def _run_pair(kw, steps=5, skip=None, lr_change=None):
    ref_p = [torch.nn.Parameter(t.clone()) for t in _make()]
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in _make()]
    ref = torch.optim.Adam(ref_p, **kw)
    ours = FlatAdam(our_p, **kw)
    for s in range(steps):
        if lr_change and s == lr_change[0]:
            for o in (ref, ours):
                o.param_groups[0]["lr"] = lr_change[1]
        gs = _grads(s, skip)
        for p, q, g in zip(ref_p, our_p, gs):
            p.grad = None if g is None else g.clone()
            q.grad = None if g is None else g.clone().cuda()
        ref.step()
        ours.step()
    return ref_p, our_p, ref, ours


def _max_rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("kw", [dict(lr=1e-3), dict(lr=3e-2, weight_decay=1e-2),
                                dict(lr=1e-2, betas=(0.8, 0.99), eps=1e-6),
                                dict(lr=1e-2, maximize=True)])
def test_flat_adam_matches_torch(kw):
    ref_p, our_p, ref, ours = _run_pair(kw)
    for p, q in zip(ref_p, our_p):
        assert _max_rel(q.detach().cpu(), p.detach()) <= 2e-6          # fp32 rounding only
    sd = ours.state_dict()
    rsd = ref.state_dict()
    assert sd["param_groups"][0]["params"] == rsd["param_groups"][0]["params"]
    for k in rsd["state"]:
        assert float(sd["state"][k]["step"]) == float(rsd["state"][k]["step"]) == 5.0
        assert _max_rel(sd["state"][k]["exp_avg"].cpu(), rsd["state"][k]["exp_avg"]) <= 2e-6
        assert _max_rel(sd["state"][k]["exp_avg_sq"].cpu(), rsd["state"][k]["exp_avg_sq"]) <= 2e-6


def test_flat_adam_lr_schedule_and_flat_views():
    ref_p, our_p, ref, ours = _run_pair(dict(lr=1e-2), steps=6, lr_change=(3, 1e-3))
    for p, q in zip(ref_p, our_p):
        assert _max_rel(q.detach().cpu(), p.detach()) <= 2e-6
    g = ours._groups[0]
    rows, total, offs = chunk_table([p.numel() for p in our_p], 4096)
    assert g.arena.numel() == total
    for q, off in zip(our_p, offs):     # parameters are views of the arena, 16-byte aligned
        assert q.data_ptr() == g.arena.data_ptr() + 4 * off and q.data_ptr() % 16 == 0


def test_flat_adam_missing_and_misaligned_grads():
    kw = dict(lr=1e-2)
    ref_p = [torch.nn.Parameter(t.clone()) for t in _make()]
    our_p = [torch.nn.Parameter(t.clone().cuda()) for t in _make()]
    ours = FlatAdam(our_p, **kw)
    # torch keeps a per-tensor step count; a tensor that never gets a gradient is simply skipped by
    # both, which is the only "missing gradient" case of the reference (unused supernet branches)
    ref = torch.optim.Adam(ref_p, **kw)
    for s in range(3):
        gs = _grads(s, skip=2)
        for p, q, g in zip(ref_p, our_p, gs):
            if g is None:
                p.grad = q.grad = None
                continue
            p.grad = g.clone()
            buf = torch.zeros(g.numel() + 1, device="cuda")
            buf[1:] = g.flatten().cuda()
            q.grad = buf[1:].view(g.shape)          # 4-byte aligned only: scalar path
        ref.step()
        ours.step()
    for i, (p, q) in enumerate(zip(ref_p, our_p)):
        assert _max_rel(q.detach().cpu(), p.detach()) <= 2e-6
    assert torch.equal(our_p[2].detach().cpu(), _make()[2])


def test_flat_adam_state_dict_round_trip():
    _, our_p, _, ours = _run_pair(dict(lr=1e-2), steps=3)
    sd = copy.deepcopy(ours.state_dict())
    new_p = [torch.nn.Parameter(q.detach().clone()) for q in our_p]
    new = FlatAdam(new_p, lr=5e-1)
    new.load_state_dict(sd)
    assert new.param_groups[0]["lr"] == 1e-2
    gs = _grads(7)
    for p, q, g in zip(our_p, new_p, gs):
        p.grad = g.clone().cuda()
        q.grad = g.clone().cuda()
    ours.step()
    new.step()
    for p, q in zip(our_p, new_p):
        assert torch.equal(p.detach(), q.detach())
    assert float(new.state_dict()["state"][0]["step"]) == 4.0


def test_flat_adam_reduce_on_plateau_in_graph():
    """ReduceLROnPlateau (search.py:105-106,155-156) acts on a step captured in a CUDA graph"""
    from torch.optim.lr_scheduler import ReduceLROnPlateau
    from nas_3d_unet_b200.graph import GraphedStep
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(1000, device="cuda"))
    w_ref = torch.nn.Parameter(w.detach().cpu().clone())
    opt = FlatAdam([w], lr=1e-2)
    ref = torch.optim.Adam([w_ref], lr=1e-2)
    sch = ReduceLROnPlateau(opt, factor=0.5, patience=0)
    sch_ref = ReduceLROnPlateau(ref, factor=0.5, patience=0)

    def step(x):
        opt.zero_grad()
        loss = ((w - x) ** 2).sum()
        loss.backward()
        opt.step()
        return loss

    x = torch.ones(1000, device="cuda")
    g = GraphedStep(step, (x,), warmup=3, optimizers=[opt])
    for _ in range(3):
        ref.zero_grad()
        ((w_ref - 1.0) ** 2).sum().backward()
        ref.step()
    for i in range(4):
        g(x)
        ref.zero_grad()
        ((w_ref - 1.0) ** 2).sum().backward()
        ref.step()
        sch.step(1.0)        # no improvement -> lr halves every call after the first
        sch_ref.step(1.0)
    assert opt.param_groups[0]["lr"] == ref.param_groups[0]["lr"] < 1e-2
    assert _max_rel(w.detach().cpu(), w_ref.detach()) <= 1e-5
