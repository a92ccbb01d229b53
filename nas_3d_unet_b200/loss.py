"""Multi-region soft-Dice loss on the B200 path (reference: loss.py:6-14).

loss = 1 - mean_{n,c} (2*sum(p*t) + eps) / (sum(p) + sum(t) + eps), sums over D,H,W.
One streaming kernel produces the three sums per (n,c) in fp64; the backward is elementwise.
Predictions may be NDHWC (what our nets return) or NCDHW, labels likewise - read in place."""
import torch
import torch.nn as nn

from . import _lib
from .engine import _require_cuda, _stream, get_lib


def _voxel_strides(t):
    """(sn, sc, sv) if (d,h,w) collapse to a single voxel stride, else None"""
    N, C, D, H, W = t.shape
    sn, sc, sd, sh, sw = t.stride()
    if (H == 1 or sh == W * sw) and (D == 1 or sd == H * W * sw):
        return sn, sc, sw
    return None


class _DiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_truth, smooth):
        lib = get_lib()
        if _voxel_strides(y_pred) is None:
            y_pred = y_pred.contiguous()
        if _voxel_strides(y_truth) is None:
            y_truth = y_truth.contiguous()
        N, C, D, H, W = y_pred.shape
        V = D * H * W
        ps, ts = _voxel_strides(y_pred), _voxel_strides(y_truth)
        sums = torch.empty((N, C, 3), device=y_pred.device, dtype=torch.float64)
        loss = torch.empty((), device=y_pred.device, dtype=torch.float32)
        with torch.cuda.device(y_pred.device):
            _lib.check(lib.nas3d_dice_fwd(y_pred.data_ptr(), ps[0], ps[1], ps[2],
                                          y_truth.data_ptr(), 0 if y_truth.dtype == torch.float32 else 1,
                                          ts[0], ts[1], ts[2], N, C, V,
                                          smooth, sums.data_ptr(), loss.data_ptr(), _stream()),
                       "dice_fwd")
        ctx.save_for_backward(y_truth, sums)
        ctx.geom = (N, C, V, ps, ts, smooth, y_pred.shape, y_pred.stride())
        return loss

    @staticmethod
    def backward(ctx, gout):
        y_truth, sums = ctx.saved_tensors
        N, C, V, ps, ts, smooth, shape, stride = ctx.geom
        lib = get_lib()
        gout = gout.contiguous()
        dpred = torch.empty_strided(shape, stride, device=y_truth.device, dtype=torch.float32)
        with torch.cuda.device(y_truth.device):
            _lib.check(lib.nas3d_dice_bwd(sums.data_ptr(), gout.data_ptr(), y_truth.data_ptr(),
                                          0 if y_truth.dtype == torch.float32 else 1,
                                          ts[0], ts[1], ts[2], dpred.data_ptr(), ps[0], ps[1],
                                          ps[2], N, C, V, smooth, _stream()), "dice_bwd")
        return dpred, None, None


class WeightedDiceLoss(nn.Module):
    def __init__(self, axis=(-1, -2, -3), smooth=1e-6):
        super().__init__()
        self.axis = axis
        self.smooth = smooth

    def forward(self, y_pred, y_truth):
        if sorted(a % 5 for a in self.axis) != [2, 3, 4]:
            raise NotImplementedError("the B200 Dice kernel reduces over the three spatial axes")
        if y_pred.shape != y_truth.shape or y_pred.dim() != 5:
            raise ValueError("y_pred %s and y_truth %s must be equal 5-D shapes"
                             % (tuple(y_pred.shape), tuple(y_truth.shape)))
        _require_cuda(y_pred, "y_pred")
        if not y_truth.is_cuda:
            _require_cuda(y_truth.float(), "y_truth")      # raises: there is no CPU path
        if y_truth.dtype not in (torch.float32, torch.int8):
            # int8 {0,1} masks (generator.py:230-248) are read as they are; anything else -> fp32
            y_truth = y_truth.to(torch.int8) if y_truth.dtype in (torch.uint8, torch.bool) else y_truth.float()
        return _DiceFn.apply(y_pred, y_truth, float(self.smooth))
