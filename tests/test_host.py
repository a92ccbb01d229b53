"""CPU: host-side logic, the drop-in surface and the C-ABI library (load + exported symbols;
no compute without a GPU)."""
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import make_searched, make_supernet
from oracle import nas3d_oracle as O


def test_library_builds_and_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "nas3d_b200.h")).read()
    declared = set(re.findall(r"\b(nas3d_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), "libnas3d_b200.so does not export %s" % name
    from nas_3d_unet_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.nas3d_version() >= 100


def test_sass_is_sm100a(lib):
    from nas_3d_unet_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_cpu_tensor_fails_loudly():
    from nas_3d_unet_b200.prim_ops import OPS
    from nas_3d_unet_b200.loss import WeightedDiceLoss
    from nas_3d_unet_b200.engine import Nas3dDeviceError
    op = OPS['conv'](4)
    with pytest.raises(Nas3dDeviceError):
        op(torch.randn(1, 4, 4, 4, 4))
    with pytest.raises(Nas3dDeviceError):
        WeightedDiceLoss()(torch.rand(1, 3, 4, 4, 4), torch.rand(1, 3, 4, 4, 4))


def test_reference_error_conventions():
    from nas_3d_unet_b200.prim_ops import OPS, PoolingOp, ConvOps
    from nas_3d_unet_b200.nas import KernelNet
    with pytest.raises(KeyError):
        OPS['no_such_op']
    with pytest.raises(NotImplementedError):
        PoolingOp(4, 4, pool_type='median')
    with pytest.raises(AssertionError):
        KernelNet(4, 4, 3, 1, 3, True)          # depth must be >= 2 (nas.py:31)
    assert ConvOps(4, 4, ops_order='weight_foo').ops_list == ['weight', 'foo']


def test_op_lists_and_alpha_shapes():
    from nas_3d_unet_b200 import prim_ops
    assert prim_ops.DownOps == O.DOWN_OPS and prim_ops.UpOps == O.UP_OPS and prim_ops.NormOps == O.NORM_OPS
    assert set(prim_ops.OPS) == set(O.PRIMS)
    s = make_supernet(random_alphas=False)
    assert tuple(s.alpha2_down.shape) == (9, 6) and tuple(s.alpha2_up.shape) == (9, 4)
    assert tuple(s.alpha1_down.shape) == (9, 5) and tuple(s.alpha1_up.shape) == (9, 5)
    assert len(list(s.alphas())) == 4
    assert sum(p.numel() for p in s.parameters()) == 6854184
    assert sum(p.numel() for p in s.kernel.parameters()) == 6854004
    assert len(s.state_dict()) == 1784
    m = make_searched()
    assert sum(p.numel() for p in m.parameters()) == 1242760


def test_normal_w_share_aliases():
    from nas_3d_unet_b200.nas import ShellNet
    s = ShellNet(4, 4, 3, 2, 2, normal_w_share=True, channel_change=False)
    assert s.alpha1_up is s.alpha1_down
    assert len(list(s.alphas())) == 3


def test_state_dict_keys_match_reference(golden_nets):
    G = golden_nets
    s = make_supernet()
    assert [k for k, _ in s.named_parameters()] == list(G['supernet32/param_names'])
    m = make_searched()
    assert [k for k, _ in m.named_parameters()] == list(G['searched32/param_names'])
    sd = m.state_dict()
    assert 'down_cells.0._ops.3.depth_conv.weight' in sd and 'last_conv.0.conv.weight' in sd
    assert tuple(sd['up_cells.0._ops.0.conv.weight'].shape) == (64, 64, 3, 3, 3)


def test_genotype_parse_matches_oracle_and_reference(golden_nets):
    from nas_3d_unet_b200.genotype import GenoParser, Genotype
    rng = np.random.default_rng(0)
    for trial in range(50):
        a1 = torch.softmax(torch.from_numpy(rng.normal(size=(9, 5)).astype(np.float32)), -1).numpy()
        a2d = torch.softmax(torch.from_numpy(rng.normal(size=(9, 6)).astype(np.float32)), -1).numpy()
        a2u = torch.softmax(torch.from_numpy(rng.normal(size=(9, 4)).astype(np.float32)), -1).numpy()
        if trial % 5 == 0:      # ties
            a1[:] = 0.2
            a2d[:] = 1 / 6
            a2u[:] = 0.25
        p = GenoParser(3)
        assert p.parse(a1, a2d) == O.geno_parse(a1, a2d, 3, True)
        assert p.parse(a1, a2u, downward=False) == O.geno_parse(a1, a2u, 3, False)
    s = make_supernet(random_alphas=False)
    g = s.get_gene()
    ref = json.loads(str(golden_nets['gene_init'][0]))
    assert isinstance(g, Genotype)
    assert [list(t) for t in g.down] == ref['down'] and [list(t) for t in g.up] == ref['up']
    s = make_supernet(random_alphas=True)
    g = s.get_gene()
    ref = json.loads(str(golden_nets['supernet32/gene'][0]))
    assert [list(t) for t in g.down] == ref['down'] and [list(t) for t in g.up] == ref['up']


def test_dropin_shims_import():
    d = os.path.join(ROOT, "nas_3d_unet_b200", "dropin")
    saved = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in ('nas', 'searched', 'loss', 'genotype', 'prim_ops', 'cell')
                  if k in sys.modules}
    sys.path.insert(0, d)
    try:
        import nas, searched, loss, genotype, prim_ops, cell   # noqa: E401
        from nas_3d_unet_b200 import nas as ours
        assert nas.ShellNet is ours.ShellNet and searched.SearchedNet.__module__ == 'nas_3d_unet_b200.searched'
        assert loss.WeightedDiceLoss and genotype.Genotype and prim_ops.OPS and cell.Cell
    finally:
        sys.path[:] = saved
        for k in ('nas', 'searched', 'loss', 'genotype', 'prim_ops', 'cell'):
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)


def test_flat_adam_chunk_table_and_cpu_refusal():
    """optim.FlatAdam: chunk table layout; no CPU fallback (step on CPU parameters raises)"""
    import torch
    from nas_3d_unet_b200.optim import FlatAdam, chunk_table
    rows, total, offs = chunk_table([5, 4096, 9000, 0, 3], 4096)
    assert offs == [0, 8, 4104, 13104, 13104] and total == 13108
    assert rows == [(0, 0, 0, 5), (1, 8, 0, 4096), (2, 4104, 0, 4096), (2, 8200, 4096, 4096),
                    (2, 12296, 8192, 808), (4, 13104, 0, 3)]
    assert all(o % 4 == 0 for o in offs)
    p = torch.nn.Parameter(torch.zeros(3))
    opt = FlatAdam([p], lr=1e-3)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(NotImplementedError):
        FlatAdam([p], amsgrad=True)


def test_compat_layer(tmp_path):
    """SURVEY App. F shims: verbose= on ReduceLROnPlateau, torch.load of pickled containers,
    np.int, tqdm.notebook, genotype file round trip in the drivers' (str, count) format"""
    import collections
    import pickle
    import torch
    from nas_3d_unet_b200 import compat
    from nas_3d_unet_b200.genotype import Genotype
    from oracle import nas3d_oracle as O
    orig_load, orig_sched = torch.load, torch.optim.lr_scheduler.ReduceLROnPlateau
    try:
        compat.install()
        compat.install()        # idempotent
        assert np.int is int
        from tqdm.notebook import tqdm as nb_tqdm
        assert list(nb_tqdm(range(3), disable=True)) == [0, 1, 2]
        from torch.optim.lr_scheduler import ReduceLROnPlateau
        w = torch.nn.Parameter(torch.zeros(2))
        opt = torch.optim.Adam([w])
        sch = ReduceLROnPlateau(opt, verbose=True, factor=0.5)     # search.py:105
        for _ in range(12):
            sch.step(1.0)
        assert opt.param_groups[0]["lr"] == 5e-4
        sd = sch.state_dict()
        ReduceLROnPlateau(opt, factor=0.5).load_state_dict(sd)
        ck = tmp_path / "last.pt"
        torch.save({"geno_count": collections.Counter(a=2), "hist": collections.defaultdict(list),
                    "w": w.detach()}, ck)
        back = torch.load(ck, map_location="cpu")                   # search.py:112
        assert back["geno_count"]["a"] == 2 and torch.equal(back["w"], w.detach())
    finally:
        compat.uninstall()
    assert torch.load is orig_load and torch.optim.lr_scheduler.ReduceLROnPlateau is orig_sched
    gene = Genotype(down=O.G0.down, up=O.G0.up)
    path = tmp_path / "best_genotype.pkl"
    compat.save_genotype(gene, path, count=7)
    text, count = pickle.load(open(path, "rb"))
    assert count == 7 and eval(text, {"Genotype": Genotype}) == gene        # what train.py:38 does
    g2, c2 = compat.load_genotype(path)
    assert g2 == gene and c2 == 7
    with pytest.raises(ValueError):
        compat.parse_genotype("__import__('os').system('true')")
