"""The supernet (KernelNet) and its architecture parameters (ShellNet), B200-native.

Reference surface: nas.py:13-78 (KernelNet), :81-135 (ShellNet).  Constructor signatures,
attribute names (`kernel`, `stem0/1`, `down_cells`, `up_cells`, `last_conv`, `alpha{1,2}_{down,up}`)
and parameter shapes are the reference's, so its 1 784-entry state_dict loads unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .prim_ops import ConvOps, DownOps, UpOps, NormOps
from .cell import Cell
from .genotype import Genotype, GenoParser

FLAG_DEBUG = False


def _u_shape(n_nodes, init_n_kernels, depth, channel_change):
    """channel bookkeeping of the U (nas.py:25,36-49): yields ('down'|'up', c0, c1, c_node)"""
    c0 = c1 = n_nodes * init_n_kernels
    c_node = init_n_kernels
    skips = [c0, c1]
    plan = []
    for _ in range(depth):
        if channel_change:
            c_node *= 2
        plan.append(('down', c0, c1, c_node))
        c0, c1 = c1, n_nodes * c_node
        skips.append(c1)
    skips.pop()
    for _ in range(depth + 1):
        c0 = skips.pop()
        plan.append(('up', c0, c1, c_node))
        c1 = n_nodes * c_node
        if channel_change:
            c_node //= 2
    return plan, c1


class KernelNet(nn.Module):
    def __init__(self, in_channels, init_n_kernels, out_channels, depth, n_nodes, channel_change):
        '''
        The U-shaped supernet: 2 stems, `depth` down cells, `depth+1` up cells, 1x1 head.
        in_channels: number of MRI modalities;  out_channels: number of tumour regions.
        init_n_kernels: node width of the first cell;  channel_change: double/halve per level.
        '''
        super().__init__()
        c_stem = n_nodes * init_n_kernels
        self.stem0 = ConvOps(in_channels, c_stem, kernel_size=1, ops_order='weight_norm')
        self.stem1 = ConvOps(in_channels, c_stem, kernel_size=3, stride=2, ops_order='weight_norm')

        assert depth >= 2, 'depth must >= 2'

        self.down_cells = nn.ModuleList()
        self.up_cells = nn.ModuleList()
        plan, c_last = _u_shape(n_nodes, init_n_kernels, depth, channel_change)
        for kind, c0, c1, c_node in plan:
            if kind == 'down':
                self.down_cells.append(Cell(n_nodes, c0, c1, c_node))
            else:
                self.up_cells.append(Cell(n_nodes, c0, c1, c_node, downward=False))
        self.last_conv = nn.Sequential(ConvOps(c_last, out_channels, kernel_size=1,
                                               dropout_rate=0.1, ops_order='weight'),
                                       nn.Sigmoid())

    def forward(self, x, alpha1_down, alpha1_up, alpha2_down, alpha2_up):
        '''alphas are already softmaxed (ShellNet.forward does it)'''
        return engine.run_module(self, (x,), (alpha1_down, alpha1_up, alpha2_down, alpha2_up))

    def _run(self, ctx, x, a1d, a1u, a2d, a2u):
        s0 = engine.materialize(ctx, self.stem0._run(ctx, x))
        s1 = engine.materialize(ctx, self.stem1._run(ctx, x))
        skips = [s0, s1]
        for cell in self.down_cells:
            s0, s1 = s1, cell._run(ctx, s0, s1, a1d, a2d, virtual_cat=engine.virtual_cat_enabled())
            skips.append(s1)
        if FLAG_DEBUG:
            print('x.shape = ', (x.N, x.C, x.D, x.H, x.W))
            for s in skips:
                print((s.N, s.C, s.D, s.H, s.W))
        skips.pop()
        for cell in self.up_cells:
            s0 = skips.pop()
            s1 = cell._run(ctx, s0, s1, a1u, a2u, virtual_cat=engine.virtual_cat_enabled())
            if FLAG_DEBUG:
                print((s1.N, s1.C, s1.D, s1.H, s1.W))
        return engine.materialize(ctx, self.last_conv[0]._run(ctx, s1, sigmoid=True))


class ShellNet(nn.Module):
    def __init__(self, in_channels, init_n_kernels, out_channels, depth, n_nodes,
                 normal_w_share=False, channel_change=False):
        '''
        Holds the architecture parameters around a KernelNet.
        normal_w_share: if True the stride-1 alphas of up cells alias those of down cells.
        '''
        super().__init__()
        self.normal_w_share = normal_w_share
        self.n_nodes = n_nodes
        self.kernel = KernelNet(in_channels, init_n_kernels, out_channels, depth, n_nodes,
                                channel_change)
        self._init_alphas()

    def _init_alphas(self):
        '''one row per edge of a cell; columns follow DownOps / UpOps / NormOps'''
        n_edges = sum(range(2, 2 + self.n_nodes))
        self.alpha2_down = nn.Parameter(torch.zeros((n_edges, len(DownOps))))
        self.alpha2_up = nn.Parameter(torch.zeros((n_edges, len(UpOps))))
        self.alpha1_down = nn.Parameter(torch.zeros((n_edges, len(NormOps))))
        self.alpha1_up = self.alpha1_down if self.normal_w_share else nn.Parameter(
            torch.zeros((n_edges, len(NormOps))))
        self._alphas = [(name, p) for name, p in self.named_parameters() if 'alpha' in name]

    def alphas(self):
        for _, p in self._alphas:
            yield p

    def forward(self, x):
        # the four 9xK softmaxes stay in torch (180 floats)
        return self.kernel(x,
                           F.softmax(self.alpha1_down, dim=-1),
                           F.softmax(self.alpha1_up, dim=-1),
                           F.softmax(self.alpha2_down, dim=-1),
                           F.softmax(self.alpha2_up, dim=-1))

    def get_gene(self):
        parser = GenoParser(self.n_nodes)

        def host(p):
            return F.softmax(p, dim=-1).detach().cpu().numpy()
        down = parser.parse(host(self.alpha1_down), host(self.alpha2_down))
        up = parser.parse(host(self.alpha1_up), host(self.alpha2_up), downward=False)
        return Genotype(down=down, up=up)
