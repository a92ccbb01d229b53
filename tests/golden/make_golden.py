"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, which only exists in the build container) on seeded synthetic inputs.

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures pin (a) the oracle
(oracle/nas3d_oracle.py) and (b) the CUDA path against outputs of the reference itself.
Everything is CPU fp32, torch.manual_seed-determined, dropout disabled (eval mode or p=0).
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("NAS3D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import prim_ops as ref_prim  # noqa: E402
import cell as ref_cell  # noqa: E402
import nas as ref_nas  # noqa: E402
import searched as ref_searched  # noqa: E402
import loss as ref_loss  # noqa: E402
import genotype as ref_genotype  # noqa: E402

sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.nas3d_oracle import G0, synthetic_batch  # noqa: E402  (constants + data recipe only)

torch.set_num_threads(os.cpu_count())


def tstats(t):
    t = t.detach().double().reshape(-1)
    return np.array([t.sum().item(), t.norm().item(), t.abs().max().item(),
                     t[0].item(), t[t.numel() // 2].item(), t[-1].item()], dtype=np.float64)


def sample_idx(n, k=4096):
    if n <= k:
        return np.arange(n)
    return np.unique(np.linspace(0, n - 1, k).astype(np.int64))


def prim_cases():
    cases = []
    for name in ref_prim.OPS:
        for c in ((8, 32) if name in ('conv', 'identity', 'down_conv', 'up_conv', 'dep_conv',
                                      'se_conv', 'down_se_conv') else (8,)):
            cases.append((name, c))
    return cases


def golden_prims():
    out = {}
    meta = []
    for name, c in prim_cases():
        torch.manual_seed(100 + len(meta))
        op = ref_prim.OPS[name](c)
        # non-trivial affine params so GroupNorm gradients are exercised
        with torch.no_grad():
            for k, p in op.named_parameters():
                if 'norm' in k:
                    p.add_(0.3 * torch.randn_like(p))
        s = 4 if name.startswith('up_') else 6
        g = torch.Generator().manual_seed(7 + len(meta))
        x = torch.randn(2, c, s, s, s, generator=g).requires_grad_(True)
        y = op(x)
        r = torch.randn(y.shape, generator=g)
        (y * r).sum().backward()
        key = '%s_c%d' % (name, c)
        out[key + '/y'] = y.detach().numpy()
        out[key + '/dx'] = x.grad.numpy()
        for k, p in op.named_parameters():
            out[key + '/param/' + k] = p.detach().numpy()
            out[key + '/grad/' + k] = p.grad.numpy()
        meta.append({'name': name, 'c': c, 'size': s, 'model_seed': 100 + len(meta),
                     'data_seed': 7 + len(meta)})
    np.savez_compressed(os.path.join(HERE, 'prims.npz'), **out)
    with open(os.path.join(HERE, 'prims.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('prims:', len(meta), 'cases')


def net_record(model, x, y, prefix, out, extra_alpha=None):
    lossf = ref_loss.WeightedDiceLoss()
    pred = model(x)
    L = lossf(pred, y)
    L.backward()
    flat = pred.detach().reshape(-1)
    idx = sample_idx(flat.numel())
    out[prefix + '/pred_idx'] = idx
    out[prefix + '/pred'] = flat[idx].numpy()
    out[prefix + '/pred_stats'] = tstats(pred)
    out[prefix + '/loss'] = np.array([L.item()], dtype=np.float64)
    names, gstats, pstats = [], [], []
    flatg = []
    for k, p in model.named_parameters():
        names.append(k)
        pstats.append(tstats(p))
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        gstats.append(tstats(g))
        flatg.append(g.reshape(-1))
    flatg = torch.cat(flatg)
    gi = sample_idx(flatg.numel(), 16384)
    out[prefix + '/grad_idx'] = gi
    out[prefix + '/grad_sample'] = flatg[gi].numpy()
    out[prefix + '/grad_flat_stats'] = tstats(flatg)
    out[prefix + '/grad_stats'] = np.stack(gstats)
    out[prefix + '/param_stats'] = np.stack(pstats)
    out[prefix + '/param_names'] = np.array(names)
    out[prefix + '/param_shapes'] = np.array([json.dumps(list(p.shape)) for _, p in model.named_parameters()])
    return L.item()


def golden_nets():
    out = {}
    gene = ref_genotype.Genotype(down=G0.down, up=G0.up)
    # searched-G0, 32^3, batch 2, eval mode (dropout off; GroupNorm has no train/eval switch)
    torch.manual_seed(0)
    m = ref_searched.SearchedNet(4, 4, 3, 4, 3, True, gene)
    m.eval()
    x, y = synthetic_batch(2, 32, seed=1)
    net_record(m, x, y, 'searched32', out)
    # same net on "brain-like" input (0 or U(10,110)) - exercises GN statistics at scale
    m.zero_grad()
    x, y = synthetic_batch(1, 32, seed=2, brain_like=True)
    net_record(m, x, y, 'searched32_brain', out)

    # supernet, 32^3, batch 1
    torch.manual_seed(0)
    s = ref_nas.ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True)
    s.eval()
    with torch.no_grad():   # alphas away from the all-equal point
        ga = torch.Generator().manual_seed(5)
        for p in s.alphas():
            p.copy_(0.5 * torch.randn(p.shape, generator=ga))
    x, y = synthetic_batch(1, 32, seed=3)
    net_record(s, x, y, 'supernet32', out)
    for k in ('alpha1_down', 'alpha1_up', 'alpha2_down', 'alpha2_up'):
        out['supernet32/alpha/' + k] = getattr(s, k).detach().numpy()
        out['supernet32/dalpha/' + k] = getattr(s, k).grad.numpy()
    gene_s = s.get_gene()
    out['supernet32/gene'] = np.array([json.dumps({'down': gene_s.down, 'up': gene_s.up})])

    # zero-alpha genotype KAT (SURVEY App. B G_init)
    torch.manual_seed(0)
    s0 = ref_nas.ShellNet(4, 4, 3, 4, 3, channel_change=True)
    g0 = s0.get_gene()
    out['gene_init'] = np.array([json.dumps({'down': g0.down, 'up': g0.up})])

    # two search steps exactly as search.py:222-238 at 32^3 (Adam defaults, dropout p forced to 0)
    torch.manual_seed(0)
    s = ref_nas.ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True)
    s.kernel.last_conv[0].dropout.p = 0.0
    s.train()
    lossf = ref_loss.WeightedDiceLoss()
    optim_shell = torch.optim.Adam(s.alphas())
    optim_kernel = torch.optim.Adam(s.kernel.parameters())
    g = torch.Generator().manual_seed(1)
    losses = []
    for step in range(2):
        x = torch.randn(1, 4, 32, 32, 32, generator=g)
        y = (torch.rand(1, 3, 32, 32, 32, generator=g) > 0.7).float()
        vx = torch.randn(1, 4, 32, 32, 32, generator=g)
        vy = (torch.rand(1, 3, 32, 32, 32, generator=g) > 0.7).float()
        optim_shell.zero_grad()
        vl = lossf(s(vx), vy)
        vl.backward()
        optim_shell.step()
        optim_kernel.zero_grad()
        l = lossf(s(x), y)
        l.backward()
        optim_kernel.step()
        losses.append([vl.item(), l.item()])
    out['search2/losses'] = np.array(losses, dtype=np.float64)
    out['search2/alpha2_down'] = s.alpha2_down.detach().numpy()
    out['search2/alpha1_up'] = s.alpha1_up.detach().numpy()
    np.savez_compressed(os.path.join(HERE, 'nets.npz'), **out)
    print('nets: done; search2 losses', losses)


def golden_cells():
    """one supernet down cell and one up cell stand-alone (cell.py surface)"""
    out = {}
    for tag, downward in (('down', True), ('up', False)):
        torch.manual_seed(11)
        c = ref_cell.Cell(3, 12, 24, 8, downward=downward)
        g = torch.Generator().manual_seed(12)
        if downward:
            x0 = torch.randn(1, 12, 8, 8, 8, generator=g, requires_grad=True)
            x1 = torch.randn(1, 24, 4, 4, 4, generator=g, requires_grad=True)
            a2 = torch.softmax(torch.randn(9, 6, generator=g), -1).requires_grad_(True)
        else:
            x0 = torch.randn(1, 12, 8, 8, 8, generator=g, requires_grad=True)
            x1 = torch.randn(1, 24, 4, 4, 4, generator=g, requires_grad=True)
            a2 = torch.softmax(torch.randn(9, 4, generator=g), -1).requires_grad_(True)
        a1 = torch.softmax(torch.randn(9, 5, generator=g), -1).requires_grad_(True)
        y = c(x0, x1, a1, a2)
        r = torch.randn(y.shape, generator=g)
        (y * r).sum().backward()
        out[tag + '/y'] = y.detach().numpy()
        out[tag + '/dx0'] = x0.grad.numpy()
        out[tag + '/dx1'] = x1.grad.numpy()
        out[tag + '/da1'] = a1.grad.numpy()
        out[tag + '/da2'] = a2.grad.numpy()
        gs = [tstats(p.grad if p.grad is not None else torch.zeros_like(p)) for p in c.parameters()]
        out[tag + '/grad_stats'] = np.stack(gs)
    np.savez_compressed(os.path.join(HERE, 'cells.npz'), **out)
    print('cells: done')


def main():
    if '--permute' in sys.argv:
        golden_permute()
        return
    if '--datapath' not in sys.argv:
        golden_prims()
        golden_cells()
        golden_nets()
    golden_datapath()


def _stub_heavy_imports():
    import types
    for name in ("h5py", "nibabel", "nilearn", "nilearn.image", "tqdm.notebook", "ipywidgets"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.tqdm = lambda it=None, **k: it
            m.new_img_like = None
            m.resample_to_img = None
            m.reorder_img = None
            m.Nifti1Image = None
            sys.modules[name] = m


def golden_permute():
    """all 48 cube permutations of the reference's augment.permute_data (augment.py:105-132) on a
    labelled (2,5,5,5) cube, keys in sorted order -> tests/golden/permute.npz"""
    _stub_heavy_imports()
    import augment as ref_augment
    keys = sorted(ref_augment.generate_permutation_keys())
    assert len(keys) == 48
    data = np.arange(2 * 125, dtype=np.int16).reshape(2, 5, 5, 5)
    out = np.stack([np.ascontiguousarray(ref_augment.permute_data(data, k)) for k in keys])
    flat = np.array([[k[0][0], k[0][1], k[1], k[2], k[3], k[4]] for k in keys], dtype=np.int8)
    np.savez_compressed(os.path.join(HERE, 'permute.npz'), keys=flat, data=data, out=out)
    print('permute: done', out.shape)


def golden_datapath():
    """integer / data-path KATs from the reference's own patches.py / prediction.py / generator.py
    (their heavy imports - h5py, nibabel, tqdm.notebook ... - are stubbed; only pure-numpy
    functions are called)"""
    import types
    np.int = int      # removed from numpy >= 1.24 (patches.py:75,188,194)
    np.bool = bool
    for name in ("h5py", "nibabel", "nilearn", "nilearn.image", "tqdm.notebook", "ipywidgets"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.tqdm = lambda it=None, **k: it
            m.new_img_like = None
            m.resample_to_img = None
            m.reorder_img = None
            sys.modules[name] = m
    import patches as ref_patches
    out = {}
    cases = [((240, 240, 155), (128, 128, 128), None), ((240, 240, 155), (64, 64, 64), None),
             ((240, 240, 155), (128, 128, 128), 64), ((240, 240, 155), (64, 64, 64), 17),
             ((160, 192, 140), (128, 128, 128), None), ((160, 192, 140), (64, 64, 64), None),
             ((100, 120, 90), (128, 128, 128), None), ((138, 173, 141), (64, 64, 64), 23)]
    for i, (img, ps, ov) in enumerate(cases):
        c = ref_patches.patching(img, ps, overlap=ov)
        out['patching/%d/corners' % i] = np.asarray(c, dtype=np.int64)
        out['patching/%d/args' % i] = np.array(list(img) + list(ps) + [-1 if ov is None else ov])
    # OOB patch extraction (patches.py KAT of SURVEY App. D)
    data = np.arange(4 * 10 ** 3, dtype=np.float32).reshape(4, 10, 10, 10)
    out['get_patch/kat'] = ref_patches.get_patch_from_3d_data(data, (8, 8, 8), (-3, 5, 2))
    # stitch: random fp32 patches on a ragged brain box with autofit corners
    rng = np.random.default_rng(7)
    shape = (3, 21, 25, 19)
    corners = ref_patches.patching(shape[1:], (16, 16, 16))
    plist = [rng.random((3, 16, 16, 16), dtype=np.float32) for _ in corners]
    st = ref_patches.stitch(plist, [np.array(c) for c in corners], shape)
    out['stitch/corners'] = np.asarray(corners, dtype=np.int64)
    out['stitch/patches'] = np.stack(plist)
    out['stitch/result'] = st
    # label assembly both ways
    sys.modules.setdefault("yaml", __import__("yaml"))
    import importlib
    pred_src = open(os.path.join(REF, "prediction.py")).read()
    gen_src = open(os.path.join(REF, "generator.py")).read()

    def extract(src, name):
        """exec one method of the reference source as a plain function (the classes themselves
        need datasets to construct)"""
        lines = src.splitlines()
        i0 = next(i for i, l in enumerate(lines) if l.strip().startswith("def %s(" % name))
        indent = len(lines[i0]) - len(lines[i0].lstrip())
        body = [lines[i0][indent:]]
        for l in lines[i0 + 1:]:
            if l.strip() and (len(l) - len(l.lstrip())) <= indent:
                break
            body.append(l[indent:] if len(l) >= indent else l)
        ns = {"np": np}
        exec("\n".join(body), ns)
        return ns[name]
    get_tumor_pred = extract(pred_src, "get_tumor_pred")
    get_multi = extract(gen_src, "get_multi_class_labels")
    p = rng.random((3, 9, 10, 11))
    p[:, 0, 0, :3] = 0.5            # exact-threshold and tie cases
    p[0, 1, 1, 1] = p[1, 1, 1, 1] = 0.7
    p[0, 2, 2, 2] = p[2, 2, 2, 2] = 0.9
    out['tumor/pred'] = p
    out['tumor/inclusive'] = get_tumor_pred(None, p, 0.5, True)
    out['tumor/exclusive'] = get_tumor_pred(None, p, 0.5, False)
    truth = rng.choice(np.array([0, 1, 2, 4, 3], dtype=np.int16), size=(2, 1, 5, 6, 7))

    class _S:
        labels = (1, 2, 4)
    out['labels/truth'] = truth
    out['labels/inclusive'] = get_multi(_S(), truth, True)
    out['labels/exclusive'] = get_multi(_S(), truth, False)
    np.savez_compressed(os.path.join(HERE, 'datapath.npz'), **out)
    print('datapath: done;', {k: v.shape for k, v in out.items() if k.startswith('patching') and 'corners' in k})


if __name__ == '__main__':
    main()
