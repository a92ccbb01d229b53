"""shared test helpers (model construction recipes mirrored from tests/golden/make_golden.py)"""
import json

import numpy as np
import torch

from oracle import nas3d_oracle as O

CFG = dict(in_channels=4, init_n_kernels=4, out_channels=3, depth=4, n_nodes=3)


def tstats(t):
    t = t.detach().double().cpu().reshape(-1)
    return np.array([t.sum().item(), t.norm().item(), t.abs().max().item(),
                     t[0].item(), t[t.numel() // 2].item(), t[-1].item()], dtype=np.float64)


def our_gene():
    from nas_3d_unet_b200.genotype import Genotype
    return Genotype(down=O.G0.down, up=O.G0.up)


def make_searched():
    from nas_3d_unet_b200.searched import SearchedNet
    torch.manual_seed(0)
    m = SearchedNet(4, 4, 3, 4, 3, True, our_gene())
    m.eval()
    return m


def make_supernet(random_alphas=True, dropout0=False, train=False):
    from nas_3d_unet_b200.nas import ShellNet
    torch.manual_seed(0)
    s = ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True)
    if random_alphas:
        with torch.no_grad():
            ga = torch.Generator().manual_seed(5)
            for p in s.alphas():
                p.copy_(0.5 * torch.randn(p.shape, generator=ga))
    if dropout0:
        s.kernel.last_conv[0].dropout.p = 0.0
    s.train(train)
    return s


def make_prim(meta_entry):
    """our primitive for a golden prims case, parameters loaded from the fixture"""
    from nas_3d_unet_b200.prim_ops import OPS
    torch.manual_seed(meta_entry['model_seed'])
    return OPS[meta_entry['name']](meta_entry['c'])


def prim_inputs(meta_entry, y_shape=None):
    g = torch.Generator().manual_seed(meta_entry['data_seed'])
    c, s = meta_entry['c'], meta_entry['size']
    x = torch.randn(2, c, s, s, s, generator=g)
    r = torch.randn(y_shape, generator=g) if y_shape is not None else None
    return x, r, g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def gene_to_json(g):
    return json.dumps({'down': [list(t) for t in g.down], 'up': [list(t) for t in g.up]})


class variant:
    """run a block with kernel-selection options changed: engine options (config.EngineConfig
    fields) and C-ABI library options (nas3d_set_option) by name"""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        import contextlib
        from nas_3d_unet_b200 import _lib, config
        self.stack = contextlib.ExitStack()
        eng = {k: v for k, v in self.kw.items() if k in config.EngineConfig.__slots__}
        if eng:
            self.stack.enter_context(config.override(**eng))
        for k, v in self.kw.items():
            if k not in eng:
                self.stack.enter_context(_lib.option(k, int(v)))
        return self

    def __exit__(self, *exc):
        self.stack.close()
        return False
