"""The 15 candidate primitives of the search space, B200-native.

Same module surface as the reference's prim_ops.py (OPS / DownOps / UpOps / NormOps at :5-45,
BaseOp :48-83, ConvOps :85-117, SEConvOp :119-153, PoolingOp :155-168, IdentityOp :170-174):
class names, constructor signatures, sub-module attribute names and therefore state_dict keys
and default initialisation are identical, so checkpoints and seeds carry over.  What differs is
below the surface: no torch.nn kernel is ever executed.  Each module's `_run(ctx, x)` launches
the hand-written sm_100a kernels through the C-ABI (engine.py) and returns a lazy `Term` so that
GroupNorm-apply / ReLU / SE-scale fold into the consumer's fused affine-sum pass.
"""
from math import ceil

import torch
import torch.nn as nn

from . import engine
from .engine import ConvSpec, Term

# name -> factory(c); insertion order is irrelevant, the three lists below fix the alpha columns
OPS = {
    'identity':      lambda c: IdentityOp(c, c),
    'se_conv':       lambda c: SEConvOp(c, c),
    'dil_conv':      lambda c: ConvOps(c, c, dilation=2),
    'dep_conv':      lambda c: ConvOps(c, c, depthwised=True),
    'conv':          lambda c: ConvOps(c, c),
    'avg_pool':      lambda c: PoolingOp(c, c, pool_type='avg'),
    'max_pool':      lambda c: PoolingOp(c, c, pool_type='max'),
    'down_se_conv':  lambda c: SEConvOp(c, c, stride=2),
    'down_dil_conv': lambda c: ConvOps(c, c, stride=2, dilation=2),
    'down_dep_conv': lambda c: ConvOps(c, c, stride=2, depthwised=True),
    'down_conv':     lambda c: ConvOps(c, c, stride=2),
    'up_se_conv':    lambda c: SEConvOp(c, c, stride=2, transposed=True),
    'up_dep_conv':   lambda c: ConvOps(c, c, stride=2, depthwised=True, transposed=True),
    'up_conv':       lambda c: ConvOps(c, c, stride=2, transposed=True),
    'up_dil_conv':   lambda c: ConvOps(c, c, stride=2, dilation=2, transposed=True),
}

# column order of alpha2_down / alpha2_up / alpha1_* (prim_ops.py:23-45)
DownOps = ['avg_pool', 'max_pool', 'down_se_conv', 'down_dil_conv', 'down_dep_conv', 'down_conv']
UpOps = ['up_se_conv', 'up_dep_conv', 'up_conv', 'up_dil_conv']
NormOps = ['identity', 'se_conv', 'dil_conv', 'dep_conv', 'conv']


def _same_padding(kernel_size, stride, dilation):
    # prim_ops.py:91,131
    return max(0, ceil((dilation * (kernel_size - 1) - stride + 1) / 2))


class BaseOp(nn.Module):
    """Executes the `ops_order` string ('weight' / 'norm' / 'act' tokens joined by '_').

    GroupNorm grouping rule (prim_ops.py:57): 1 group unless out_channels is a multiple of 16,
    then 16 channels per group.  Dropout3d, if any, acts right before the weight op."""

    def __init__(self, in_channels, out_channels, dropout_rate=0, ops_order='weight_norm_act'):
        super().__init__()
        self.ops_list = ops_order.split('_')
        if 'norm' in self.ops_list:
            group = 1 if out_channels % 16 != 0 else out_channels // 16
            self.norm = nn.GroupNorm(group, out_channels)
        else:
            self.norm = None
        self.activation = nn.ReLU() if 'act' in self.ops_list else None
        self.dropout = nn.Dropout3d(dropout_rate) if dropout_rate > 0 else None

    # -- public torch surface ---------------------------------------------------------
    def forward(self, x):
        return engine.run_module(self, (x,))

    # -- B200 path --------------------------------------------------------------------
    def _weight_run(self, ctx, x, in_relu, in_scale, sigmoid):
        """subclass hook: apply the weight op to Act x; returns a Term"""
        raise NotImplementedError

    def _dropout_scale(self, ctx, x):
        """Dropout3d as a per-(n,c) scale drawn from torch's generator exactly like
        feature_dropout does (empty(N,C,1,1,1).bernoulli_(1-p).div_(1-p))."""
        if self.dropout is None or not ctx.training or self.dropout.p == 0:
            return None
        p = self.dropout.p
        noise = torch.empty((x.N, x.C, 1, 1, 1), device=ctx.device, dtype=torch.float32)
        if p >= 1:
            return noise.zero_().view(x.N, x.C)
        return noise.bernoulli_(1 - p).div_(1 - p).view(x.N, x.C)

    def _run(self, ctx, x, sigmoid=False):
        state = Term(x)
        for pos, op in enumerate(self.ops_list):
            if op == 'weight':
                scale = self._dropout_scale(ctx, state.x)
                relu_only = state.relu and state.a is None and state.b is None
                if state.is_identity or (relu_only and self._fuses_prologue):
                    in_relu = state.relu
                    src = state.x
                else:
                    in_relu = False
                    src = engine.materialize(ctx, state)
                if scale is not None and not self._fuses_prologue:
                    src = engine.materialize(ctx, Term(src, scale, None, in_relu, "coef"))
                    in_relu, scale = False, None
                last = pos == len(self.ops_list) - 1
                # GroupNorm right after the weight op: its statistics come out of the conv epilogue
                self._want_stats = (not last and self.ops_list[pos + 1] == 'norm'
                                    and self.norm is not None)
                state = self._weight_run(ctx, src, in_relu, scale, sigmoid and last)
            elif op == 'norm':
                if self.norm is not None:
                    src = engine.materialize(ctx, state)
                    state = engine.gn_term(ctx, src, self.norm, relu=False)
            elif op == 'act':
                if self.activation is not None:
                    if state.relu:
                        pass                      # relu(relu(.)) == relu(.)
                    elif state.kind in ("plain", "gn") and state.alpha is None:
                        state.relu = True         # folds into the pending affine term
                    else:
                        state = Term(engine.materialize(ctx, state), relu=True)
            else:
                raise Warning('Unrecognized op: %s' % op)
        return state

    _fuses_prologue = False
    _want_stats = False


class ConvOps(BaseOp):
    """dense / dilated / depthwise-separable / transposed convolution block (prim_ops.py:85-117)"""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1,
                 transposed=False, depthwised=False, dropout_rate=0, ops_order='weight_norm_act'):
        super().__init__(in_channels, out_channels, dropout_rate, ops_order)
        self.depthwised = depthwised
        padding = _same_padding(kernel_size, stride, dilation)
        out_pad = 0 if stride == 1 else 1
        if depthwised:
            # depthwise (no dilation, as in the reference) followed by a 1x1x1 pointwise conv
            if transposed:
                self.depth_conv = nn.ConvTranspose3d(in_channels, in_channels, kernel_size,
                                                     stride=stride, padding=padding,
                                                     groups=in_channels, output_padding=out_pad)
            else:
                self.depth_conv = nn.Conv3d(in_channels, in_channels, kernel_size, stride=stride,
                                            padding=padding, groups=in_channels)
            self.point_conv = nn.Conv3d(in_channels, out_channels, kernel_size=1)
            self._specs = (ConvSpec(self.depth_conv), ConvSpec(self.point_conv))
        else:
            if transposed:
                self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size,
                                               stride=stride, padding=padding, dilation=dilation,
                                               output_padding=out_pad)
            else:
                self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride,
                                      padding=padding, dilation=dilation)
            self._specs = (ConvSpec(self.conv),)
        self._fuses_prologue = (not depthwised) and (not transposed)

    def _weight_run(self, ctx, x, in_relu, in_scale, sigmoid):
        if self.depthwised:
            mid = engine.conv(ctx, x, self.depth_conv, self._specs[0])
            y = engine.conv(ctx, mid, self.point_conv, self._specs[1], sigmoid=sigmoid,
                            stats=self._want_stats)
        else:
            y = engine.conv(ctx, x, self.conv, self._specs[0], in_relu=in_relu, in_scale=in_scale,
                            sigmoid=sigmoid, stats=self._want_stats)
        return Term(y)


class SEConvOp(BaseOp):
    """squeeze-and-excitation scale, followed by a strided (transposed) conv + GN when stride>1
    (prim_ops.py:119-153).  With stride 1 the op is just x*s: no conv, no norm."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1,
                 transposed=False, dropout_rate=0, ops_order='weight_norm'):
        super().__init__(in_channels, out_channels, dropout_rate,
                         ops_order=ops_order if stride > 1 else 'weight')
        self.stride = stride
        padding = _same_padding(kernel_size, stride, dilation)
        self.avg_pool = nn.AdaptiveAvgPool3d(1)
        self.fc = nn.Sequential(nn.Linear(in_channels, 1), nn.ReLU(),
                                nn.Linear(1, out_channels), nn.Sigmoid())
        if stride > 1:
            # NB: like the reference, `dilation` only influences the padding here
            if transposed:
                self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size,
                                               stride=stride, padding=padding, output_padding=1)
            else:
                self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride,
                                      padding=padding)
            self._spec = ConvSpec(self.conv)

    def _weight_run(self, ctx, x, in_relu, in_scale, sigmoid):
        if in_relu or in_scale is not None:
            raise NotImplementedError("SEConvOp with a fused prologue")
        scaled = engine.se_term(ctx, x, self.fc)
        if self.stride < 2:
            return scaled
        xs = engine.materialize(ctx, scaled)
        return Term(engine.conv(ctx, xs, self.conv, self._spec, sigmoid=sigmoid,
                                stats=self._want_stats))


class PoolingOp(BaseOp):
    """2x2x2 stride-2 average / max pooling and nothing else (prim_ops.py:155-168)"""

    def __init__(self, in_channels, out_channels, pool_type, kernel_size=2, stride=2,
                 ops_order='weight'):
        super().__init__(in_channels, out_channels, ops_order=ops_order)
        if pool_type == 'avg':
            self.pool = nn.AvgPool3d(kernel_size, stride=stride)
            self._kind = 0
        elif pool_type == 'max':
            self.pool = nn.MaxPool3d(kernel_size, stride=stride)
            self._kind = 1
        else:
            raise NotImplementedError
        if kernel_size != 2 or stride != 2:
            raise NotImplementedError("only the 2x2x2 stride-2 pooling of the search space is built")

    def _weight_run(self, ctx, x, in_relu, in_scale, sigmoid):
        if in_relu or in_scale is not None:
            x = engine.materialize(ctx, Term(x, in_scale, None, in_relu, "coef"))
        return Term(engine.pool2(ctx, x, self._kind))


class IdentityOp(BaseOp):
    """'identity' still normalises and rectifies: x -> GroupNorm -> ReLU (prim_ops.py:170-174)"""

    def __init__(self, in_channels, out_channels, ops_order='weight_norm_act'):
        super().__init__(in_channels, out_channels, ops_order=ops_order)

    def _weight_run(self, ctx, x, in_relu, in_scale, sigmoid):
        if in_relu or in_scale is not None:
            x = engine.materialize(ctx, Term(x, in_scale, None, in_relu, "coef"))
        return Term(x)
