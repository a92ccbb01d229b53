"""GPU: net-level parity AT THE BASELINE.json SHAPES (the goldens of test_gpu_parity.py are 32^3,
where the deepest level is 1^3 and none of the big-tensor kernels engage).

Oracle = oracle/nas3d_oracle.py run live on the host cores in fp32 and, for the headline config,
in fp64 as well.  SURVEY.md 8c metrics, fixed up front:
  logits     max_rel(ours, ref32) <= 1e-3            (max_rel = max|a-b| / max|b|)
  Dice       |ours - ref32| <= 1e-4
  gradients  (i)  flat concatenated vector: max_rel(ours, ref32) <= 1e-3
             (ii) per tensor: err(ours vs fp64) <= max(1e-3, 4 * err(ref32 vs fp64)), the error of
                  tensor k measured as max|a-b| / max(max|g64_k|, 1e-3 * max|g64|): a tensor whose
                  whole gradient is below 0.1 % of the largest gradient entry (the SE fc weights of
                  the deepest cells; a conv bias in front of a GroupNorm, analytically zero) is
                  held to that absolute scale - relative to its own tiny magnitude both fp32
                  implementations are noise (observed: 1.4e-3 ours / 1.9e-4 reference on
                  up_cells.0._ops.5.fc.0.weight in one run, below 1e-3 in the next; our sums are
                  reordered from run to run by atomics).
Each test also asserts, through the per-variant launch counters of the C-ABI, that the kernels
which only engage on big tensors really served the call."""
import numpy as np
import pytest
import torch

from oracle import nas3d_oracle as O
from helpers import make_searched, make_supernet

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
GRAD_TOL = 1e-3
DICE_TOL = 1e-4


def _delta(before, after):
    return {k: after.get(k, 0) - before.get(k, 0) for k in after}


def _dropout_mask(seed, n, p):
    """the mask our Dropout3d draws for the head input when the CUDA generator is seeded with `seed`
    (feature_dropout: empty(N,C,1,1,1).bernoulli_(1-p).div_(1-p))"""
    torch.manual_seed(seed)
    return torch.empty((n, 12, 1, 1, 1), device='cuda').bernoulli_(1 - p).div_(1 - p).cpu()


def _oracle_searched(model, x, y, mask, dtype):
    sd = O.leaf_state(model.state_dict(), dtype)
    pred = O.searched_net(sd, x.to(dtype), 4, 3, O.G0, drop_mask=None if mask is None else mask.to(dtype))
    loss = O.dice_loss(pred, y.to(dtype))
    loss.backward()
    grads = {k: sd[k].grad for k, _ in model.named_parameters()}
    return pred.detach(), loss.item(), grads


def _train_step_parity(n, p, seed, fp64_twin, expect_engaged):
    from nas_3d_unet_b200 import _lib
    from nas_3d_unet_b200.loss import WeightedDiceLoss
    model = make_searched()
    x, y = O.synthetic_batch(n, p, seed=seed, brain_like=True)
    mask = _dropout_mask(77, n, 0.5)
    ref_pred, ref_loss, ref_g = _oracle_searched(model, x, y, mask, torch.float32)
    model = model.cuda()
    model.train()
    c0 = _lib.launch_counts()
    torch.manual_seed(77)
    pred = model(x.cuda())
    loss = WeightedDiceLoss()(pred, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    used = _delta(c0, _lib.launch_counts())
    for name in expect_engaged:
        assert used.get(name, 0) > 0, ("kernel variant %s did not serve this shape" % name, used)
    names = [k for k, _ in model.named_parameters()]
    ours_g = {k: q.grad.detach().cpu() for k, q in model.named_parameters()}
    e_logit = O.max_rel(pred, ref_pred)
    e_loss = abs(loss.item() - ref_loss)
    e_flat = O.max_rel(torch.cat([ours_g[k].reshape(-1) for k in names]),
                       torch.cat([ref_g[k].reshape(-1) for k in names]))
    print("searched %dx%d^3 train: logits %.2e dice %.2e flat-grad %.2e" % (n, p, e_logit, e_loss, e_flat))
    if not fp64_twin:
        assert e_logit <= LOGIT_TOL and e_loss <= DICE_TOL and e_flat <= GRAD_TOL
        return
    pred64, loss64, g64 = _oracle_searched(model.cpu(), x, y, mask, torch.float64)
    flat64 = torch.cat([g64[k].reshape(-1) for k in names])
    e_flat_o64 = O.max_rel(torch.cat([ours_g[k].reshape(-1) for k in names]), flat64)
    e_flat_r64 = O.max_rel(torch.cat([ref_g[k].reshape(-1) for k in names]), flat64)
    print("  flat gradient vs fp64: ours %.2e, ref32 %.2e" % (e_flat_o64, e_flat_r64))
    gm = flat64.abs().max().item()
    rows = sorted(((ours_g[k].double() - g64[k]).abs().max().item() / gm,
                   (ref_g[k].double() - g64[k]).abs().max().item() / gm, k) for k in names)[-6:]
    for eo, er, k in rows:
        print("    %-50s abs err / max|g|: ours %.2e ref32 %.2e  (max|g_k| %.2e)"
              % (k, eo, er, g64[k].abs().max().item()))
    # rule (i): against ref32 when ref32 itself is accurate; the fp64 twin arbitrates otherwise
    assert e_logit <= LOGIT_TOL and e_loss <= DICE_TOL
    assert e_flat <= GRAD_TOL or e_flat_o64 <= max(GRAD_TOL, 4 * e_flat_r64), (e_flat, e_flat_o64, e_flat_r64)
    gmax = max(g64[k].abs().max().item() for k in names)
    worst = (0.0, None)
    n_small = 0
    for k in names:
        t = g64[k]
        scale = max(t.abs().max().item(), 1e-3 * gmax)
        n_small += t.abs().max().item() < 1e-3 * gmax
        e_ours = (ours_g[k].double() - t).abs().max().item() / scale
        e_ref = (ref_g[k].double() - t).abs().max().item() / scale
        assert e_ours <= max(1e-3, 4 * e_ref), (k, e_ours, e_ref)
        if e_ours > worst[0]:
            worst = (e_ours, k)
    print("  vs fp64: logits %.2e (ref32 %.2e), dice %.2e, worst tensor %s %.2e, %d of %d tensors below 0.1 %% of max|g|"
          % (O.max_rel(pred, pred64), O.max_rel(ref_pred, pred64), abs(loss.item() - loss64), worst[1],
             worst[0], n_small, len(names)))
    assert O.max_rel(pred, pred64) <= LOGIT_TOL and abs(loss.item() - loss64) <= DICE_TOL


def _check_first_adam_step(before, after, grad, what, lr=1e-3, eps=1e-8):
    """the first Adam update is  w - lr * g / (|g| + eps)  (m_hat = g, v_hat = g^2): our step applied to
    OUR gradient, exactly - independent of how the reference's rounding noise falls"""
    g = grad.double().reshape(-1)
    want = before.double().reshape(-1) - lr * g / (g.abs() + eps)
    err = (after.double().reshape(-1) - want).abs().max().item()
    assert err <= 1e-6, (what, err)


BIG_TENSOR_KERNELS = ["conv3_s1_tma_merged", "wgrad3_s1", "wgrad3_s2_tma", "wgrad3_s2_tma_cs8",
                      "affine_sum_fwd_ring", "affine_sum_bwd_reduce_ring", "affine_sum_bwd_apply_ring",
                      "pointwise_fwd_ring", "pw_bwd_fused", "umma_conv_ws", "umma_wgrad", "conv3_s2_sfb",
                      "conv3_s2_bfs"]


def test_searched_128_batch2_train_step_matches_fp32_and_fp64_oracle():
    """BASELINE.json config #4 per-GPU shape at batch 2: searched-G0, 4x128^3, train mode with the
    shared Dropout3d draw.  Ring-staged streaming kernels, merged-dimension TMA tiles, persistent
    and TMA-staged wgrads, the warp-specialised tcgen05 convs, the tcgen05 wgrad and the fused 1x1
    backward must all have run."""
    _train_step_parity(2, 128, seed=41, fp64_twin=True, expect_engaged=BIG_TENSOR_KERNELS)


def test_searched_64_batch8_train_step_matches_oracle():
    """BASELINE.json config #2: searched-G0, 4x64^3, batch 8"""
    _train_step_parity(8, 64, seed=42, fp64_twin=False,
                       expect_engaged=["conv3_s1_tma_merged", "wgrad3_s1", "affine_sum_fwd_ring",
                                       "pw_bwd_fused", "umma_conv_ws", "wgrad3_s2_tma"])


def test_supernet_64_search_step_matches_oracle():
    """BASELINE.json config #1: one full search step of search.py:222-238 at 4x64^3, batch 1 - alpha
    step on a val batch, weight step on a train batch, torch Adam for both, Dropout3d(0.1) active
    with the shared draw.  Compared per half-step: loss, d(alpha) of all four alpha matrices, flat
    weight gradient, and the parameters after the Adam steps.  Adam's first update is
    lr*g/(|g|+1e-8): where |g| is at rounding-noise level the SIGN of the update is noise in the
    reference too, so post-step values are compared only where |g_ref| >= 1e-6."""
    from nas_3d_unet_b200.loss import WeightedDiceLoss
    s = make_supernet(random_alphas=False, train=True)
    sd = O.leaf_state(s.state_dict())
    alpha_keys = ['alpha1_down', 'alpha1_up', 'alpha2_down', 'alpha2_up']
    w_keys = [k for k in sd if k.startswith('kernel.') and sd[k].requires_grad]
    vx, vy = O.synthetic_batch(1, 64, seed=51, brain_like=True)
    x, y = O.synthetic_batch(1, 64, seed=52, brain_like=True)
    m1, m2 = _dropout_mask(5, 1, 0.1), _dropout_mask(6, 1, 0.1)

    s = s.cuda()
    lossf = WeightedDiceLoss().cuda()
    optim_shell = torch.optim.Adam(s.alphas())
    optim_kernel = torch.optim.Adam(s.kernel.parameters())
    ref_shell = torch.optim.Adam([sd[k] for k in alpha_keys])
    ref_kernel = torch.optim.Adam([sd[k] for k in w_keys])

    # ---- alpha step (val batch) -----------------------------------------------------------
    optim_shell.zero_grad()
    torch.manual_seed(5)
    vl = lossf(s(vx.cuda()), vy.cuda())
    vl.backward()
    rvl = O.dice_loss(O.shell_net(sd, vx, 4, 3, drop_mask=m1), vy)
    rvl.backward()
    assert abs(vl.item() - rvl.item()) <= DICE_TOL
    for k in alpha_keys:
        e = O.max_rel(getattr(s, k).grad, sd[k].grad)
        assert e <= GRAD_TOL, (k, e)
    galpha = {k: sd[k].grad.clone() for k in alpha_keys}
    a_before = {k: getattr(s, k).detach().clone() for k in alpha_keys}
    optim_shell.step()
    ref_shell.step()
    for k in alpha_keys:
        ours, ref = getattr(s, k).detach().cpu(), sd[k].detach()
        _check_first_adam_step(a_before[k].cpu(), ours, getattr(s, k).grad.cpu(), k)
        well = galpha[k].abs() >= 1e-6
        assert (ours - ref)[well].abs().max().item() <= 2e-6, k
        with torch.no_grad():
            sd[k].copy_(ours)       # second half-step starts from identical alphas on both sides
    # ---- weight step (train batch) --------------------------------------------------------
    for k in sd:
        if sd[k].requires_grad:
            sd[k].grad = None
    optim_kernel.zero_grad()
    torch.manual_seed(6)
    l = lossf(s(x.cuda()), y.cuda())
    l.backward()
    rl = O.dice_loss(O.shell_net(sd, x, 4, 3, drop_mask=m2), y)
    rl.backward()
    assert abs(l.item() - rl.item()) <= DICE_TOL
    named = dict(s.named_parameters())
    flat_o = torch.cat([named[k].grad.reshape(-1).cpu() for k in w_keys])
    flat_r = torch.cat([sd[k].grad.reshape(-1) for k in w_keys])
    e_flat = O.max_rel(flat_o, flat_r)
    print("supernet 64^3 search step: val loss d %.2e, train loss d %.2e, flat weight grad %.2e"
          % (abs(vl.item() - rvl.item()), abs(l.item() - rl.item()), e_flat))
    assert e_flat <= GRAD_TOL
    w_before = torch.cat([named[k].detach().reshape(-1).cpu() for k in w_keys])
    optim_kernel.step()
    ref_kernel.step()
    w_o = torch.cat([named[k].detach().reshape(-1).cpu() for k in w_keys])
    w_r = torch.cat([sd[k].detach().reshape(-1) for k in w_keys])
    _check_first_adam_step(w_before, w_o, flat_o, "kernel weights")
    well = flat_r.abs() >= 1e-6
    assert int(well.sum()) > 10000
    assert (w_o - w_r)[well].abs().max().item() <= 5e-6
