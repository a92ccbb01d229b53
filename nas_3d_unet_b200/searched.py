"""The derived (searched) network, B200-native (reference surface: searched.py:10-51, :54-111)."""
import torch.nn as nn

from . import engine
from .prim_ops import OPS, ConvOps
from .genotype import Genotype
from .nas import _u_shape

FLAG_DEBUG = False


class SearchedCell(nn.Module):
    _mutates_out_grad = True

    def __init__(self, n_nodes, c0, c1, c_node, gene, downward=True):
        '''
        gene: Genotype; its .down / .up list holds, per node, two (op name, input index) pairs.
        '''
        super().__init__()
        self.n_nodes = n_nodes
        self.c_node = c_node
        self.genolist = gene.down if downward else gene.up
        self.preprocess0 = ConvOps(c0, c_node, kernel_size=1, stride=2 if downward else 1,
                                   ops_order='act_weight_norm')
        self.preprocess1 = ConvOps(c1, c_node, kernel_size=1, ops_order='act_weight_norm')
        self._ops = nn.ModuleList([OPS[name](c_node) for name, _ in self.genolist])

    @property
    def out_channels(self):
        return self.n_nodes * self.c_node

    def forward(self, x0, x1):
        return engine.run_module(self, (x0, x1))

    def _run(self, ctx, x0, x1, virtual_cat=False):
        # independent branches (the two preprocess convs; the edges of a node when they read
        # different states) are spread over stream lanes, forward and backward
        ctx.join_lanes(mark=True)
        with ctx.on_lane(0):
            p0 = engine.materialize(ctx, self.preprocess0._run(ctx, x0))
        with ctx.on_lane(1 if x0 is not x1 else 0):
            p1 = engine.materialize(ctx, self.preprocess1._run(ctx, x1))
        ctx.join_lanes()
        states = [p0, p1]
        out = None
        nodes = []
        for j in range(self.n_nodes):
            terms = []
            srcs = [self.genolist[e][1] for e in (2 * j, 2 * j + 1)]
            ctx.join_lanes(mark=True)
            for i, e in enumerate((2 * j, 2 * j + 1)):
                with ctx.on_lane(i if len(set(srcs)) == len(srcs) else 0):
                    terms.append(self._ops[e]._run(ctx, states[self.genolist[e][1]]))
            ctx.join_lanes()
            t0 = terms[0].x
            for t in terms[1:]:
                if (t.x.C, t.x.D, t.x.H, t.x.W) != (t0.C, t0.D, t0.H, t0.W):
                    raise RuntimeError("genotype mixes resolutions inside node %d: %s vs %s"
                                       % (j, (t0.C, t0.D, t0.H, t0.W), (t.x.C, t.x.D, t.x.H, t.x.W)))
            if virtual_cat:
                node = engine.new_act(t0.N, self.c_node, t0.D, t0.H, t0.W, ctx.device)
            else:
                if out is None:
                    out = engine.new_act(t0.N, self.out_channels, t0.D, t0.H, t0.W, ctx.device)
                node = out.slice(j * self.c_node, (j + 1) * self.c_node)
            engine.affine_sum(ctx, terms, node)
            nodes.append(node)
            states.append(node)
        if virtual_cat:
            return engine.CatAct(nodes)
        engine.bind_concat(ctx, out, nodes, self.c_node)
        return out


class SearchedNet(nn.Module):
    def __init__(self, in_channels, init_n_kernels, out_channels, depth, n_nodes, channel_change,
                 gene):
        '''same U as the supernet with every MixedOp replaced by the genotype's choice'''
        super().__init__()
        c_stem = n_nodes * init_n_kernels
        self.stem0 = ConvOps(in_channels, c_stem, kernel_size=1, ops_order='weight_norm')
        self.stem1 = ConvOps(in_channels, c_stem, kernel_size=3, stride=2, ops_order='weight_norm')
        self.down_cells = nn.ModuleList()
        self.up_cells = nn.ModuleList()
        plan, c_last = _u_shape(n_nodes, init_n_kernels, depth, channel_change)
        for kind, c0, c1, c_node in plan:
            if kind == 'down':
                self.down_cells.append(SearchedCell(n_nodes, c0, c1, c_node, gene))
            else:
                self.up_cells.append(SearchedCell(n_nodes, c0, c1, c_node, gene, downward=False))
        # dropout 0.5 here vs 0.1 while searching (searched.py:91-92)
        self.last_conv = nn.Sequential(ConvOps(c_last, out_channels, kernel_size=1,
                                               dropout_rate=0.5, ops_order='weight'),
                                       nn.Sigmoid())

    def forward(self, x):
        return engine.run_module(self, (x,))

    def _run(self, ctx, x):
        s0 = engine.materialize(ctx, self.stem0._run(ctx, x))
        s1 = engine.materialize(ctx, self.stem1._run(ctx, x))
        skips = [s0, s1]
        for cell in self.down_cells:
            s0, s1 = s1, cell._run(ctx, s0, s1, virtual_cat=engine.virtual_cat_enabled())
            skips.append(s1)
        if FLAG_DEBUG:
            print('x.shape = ', (x.N, x.C, x.D, x.H, x.W))
            for s in skips:
                print((s.N, s.C, s.D, s.H, s.W))
        skips.pop()
        for cell in self.up_cells:
            s0 = skips.pop()
            s1 = cell._run(ctx, s0, s1, virtual_cat=engine.virtual_cat_enabled())
            if FLAG_DEBUG:
                print((s1.N, s1.C, s1.D, s1.H, s1.W))
        return engine.materialize(ctx, self.last_conv[0]._run(ctx, s1, sigmoid=True))
