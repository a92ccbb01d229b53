"""MixedOp and Cell of the supernet, B200-native (reference surface: cell.py:8-33, :35-82).

A MixedOp never materialises its candidates' normalised outputs: every candidate returns a
lazy Term and the *node* (sum over incoming edges of sum over candidates of w*op(x),
cell.py:30,32,81) is evaluated by ONE fused affine-sum launch that writes straight into the
node's channel slice of the cell output (the torch.cat of cell.py:82 costs nothing).
"""
import torch.nn as nn

from . import engine
from .prim_ops import OPS, DownOps, UpOps, NormOps, ConvOps


class MixedOp(nn.Module):
    def __init__(self, channels, stride, transposed=False):
        '''
        channels: in_channels == out_channels for MixedOp
        '''
        super().__init__()
        self._ops = nn.ModuleList()
        self.stride = stride
        if stride == 1:
            family = NormOps
        elif transposed:
            family = UpOps
        else:
            family = DownOps
        for name in family:
            self._ops.append(OPS[name](channels))

    def forward(self, x, alpha1, alpha2):
        '''
        alpha1: weights (one row of softmax(alpha)) used when stride == 1
        alpha2: weights used when stride == 2
        '''
        return engine.run_module(self, (x,), (alpha1, alpha2))

    def _terms(self, ctx, x, alpha1, alpha2, row):
        """candidate Terms tagged with their softmax weight (alpha matrix, row, column)"""
        alpha = alpha1 if self.stride == 1 else alpha2
        if alpha.ncol < len(self._ops):
            raise ValueError("alpha row has %d weights for %d candidate ops" % (alpha.ncol, len(self._ops)))
        terms = []
        # every family holds a candidate that needs the per-(n,c) moments of x (identity / se_conv /
        # down_se_conv / up_se_conv share them through x.S): produce them BEFORE the candidates fork
        # onto streams of their own
        if x.S is None:
            x.S = engine.moments(ctx, x)
        for k, op in enumerate(self._ops):
            with ctx.on_sublane(k):
                t = op._run(ctx, x)
                if t.alpha is not None:      # cannot happen for the 15 primitives
                    t = engine.Term(engine.materialize(ctx, t))
            t.alpha = (alpha, row, k)
            terms.append(t)
        return terms

    def _run(self, ctx, x, alpha1, alpha2):
        terms = self._terms(ctx, x, alpha1, alpha2, 0)
        ctx.join_lanes()             # the candidates' forward streams (engine.ExecCtx.on_sublane)
        t0 = terms[0].x
        out = engine.new_act(t0.N, t0.C, t0.D, t0.H, t0.W, ctx.device)
        return engine.affine_sum(ctx, terms, out)


class Cell(nn.Module):
    _mutates_out_grad = True

    def __init__(self, n_nodes, c0, c1, c_node, downward=True):
        '''
        n_nodes: How many nodes in a cell.
        c0, c1: in_channels for two inputs.
        c_node: out_channels for each node.
        downward: If True, this is a downward block, otherwise, an upward block.
        '''
        super().__init__()
        self.n_nodes = n_nodes
        self.c_node = c_node
        self.preprocess0 = ConvOps(c0, c_node, kernel_size=1, stride=2 if downward else 1,
                                   ops_order='act_weight_norm')
        self.preprocess1 = ConvOps(c1, c_node, kernel_size=1, ops_order='act_weight_norm')
        self._ops = nn.ModuleList()
        # node j has 2+j incoming edges; which of them change resolution (cell.py:55-59):
        #   down cell: the edges from the two cell inputs;  up cell: the edge from input 1 only
        for n_edges in range(2, 2 + n_nodes):
            for src in range(n_edges):
                if downward:
                    self._ops.append(MixedOp(c_node, stride=2 if src <= 1 else 1))
                else:
                    self._ops.append(MixedOp(c_node, stride=2 if src == 1 else 1, transposed=True))

    @property
    def out_channels(self):
        return self.n_nodes * self.c_node

    def forward(self, x0, x1, alpha1, alpha2):
        return engine.run_module(self, (x0, x1), (alpha1, alpha2))

    def _run(self, ctx, x0, x1, alpha1, alpha2, virtual_cat=False):
        # independent branches (the two preprocess convs; the edges of a node - each reads a
        # different state) are spread over stream lanes, forward and backward
        ctx.join_lanes(mark=True)
        with ctx.on_lane(0):
            p0 = engine.materialize(ctx, self.preprocess0._run(ctx, x0))
        with ctx.on_lane(1 if x0 is not x1 else 0):
            p1 = engine.materialize(ctx, self.preprocess1._run(ctx, x1))
        ctx.join_lanes()
        states = [p0, p1]
        out = None
        nodes = []
        edge = 0
        for j in range(self.n_nodes):
            terms = []
            ctx.join_lanes(mark=True)
            for i, x in enumerate(states):
                with ctx.on_lane(i):
                    terms += self._ops[edge]._terms(ctx, x, alpha1, alpha2, edge)
                edge += 1
            ctx.join_lanes()
            t0 = terms[0].x
            if virtual_cat:
                # inside a net the concat is never materialised: the node is its own dense tensor
                # and the next 1x1x1 conv reads the nodes as parts (engine.CatAct)
                node = engine.new_act(t0.N, self.c_node, t0.D, t0.H, t0.W, ctx.device)
            else:
                if out is None:
                    out = engine.new_act(t0.N, self.out_channels, t0.D, t0.H, t0.W, ctx.device)
                node = out.slice(j * self.c_node, (j + 1) * self.c_node)
            engine.affine_sum(ctx, terms, node)
            nodes.append(node)
            states.append(node)
        if len(states) - 2 != self.n_nodes:
            raise AssertionError
        if virtual_cat:
            return engine.CatAct(nodes)
        engine.bind_concat(ctx, out, nodes, self.c_node)
        return out
