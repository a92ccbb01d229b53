"""Per-launch CUDA-event timing of the C-ABI calls, with the ALGORITHMIC bytes / flops of
each launch derived from its arguments.  Used by bench.py's roofline pass only.

Algorithmic bytes of a launch = unique input + output + parameter bytes it must move
(SURVEY.md §8d); algorithmic flops of a conv = 2*N*Vsmall*Cs*(Cb/groups)*k^3.
"""
import ctypes as C
from collections import defaultdict

import torch

from . import _lib


def _conv_work(d, wgrad=False):
    d = d._obj if hasattr(d, "_obj") else d
    taps = d.k ** 3
    vs = d.N * d.Ds * d.Hs * d.Ws
    vb = d.N * d.Db * d.Hb * d.Wb
    g = d.Cb if d.depthwise else 1
    flops = 2.0 * vs * d.Cs * (d.Cb / g) * taps
    wbytes = 4.0 * d.Cs * (d.Cb / g) * taps
    # a stride-2 dilation-2 conv touches only the even voxels of the big tensor
    frac = 1.0
    if d.stride == 2 and d.dil == 2:
        frac = 0.125
    elif d.stride == 2 and d.k == 1:
        frac = 0.125
    byts = 4.0 * (vs * d.Cs + vb * d.Cb * frac) + wbytes
    tag = "k%d s%d d%d %s C%d->%d @%dx%dx%d" % (
        d.k, d.stride, d.dil, "dw" if d.depthwise else "dense", d.Cb, d.Cs, d.Ds, d.Hs, d.Ws)
    return byts, flops, tag


def _work(name, args):
    """(algorithmic bytes, flops, tag) of one call"""
    if name in ("nas3d_conv_small_from_big", "nas3d_conv_big_from_small", "nas3d_conv_wgrad",
                "nas3d_conv_wgrad_ws"):
        return _conv_work(args[0])
    if name == "nas3d_conv1x1_bwd_fused":
        # reads dsmall + big once, writes dbig: the fused pair does the flops of dgrad AND wgrad
        d = args[0]._obj if hasattr(args[0], "_obj") else args[0]
        b, f, tag = _conv_work(d)
        vb = d.N * d.Db * d.Hb * d.Wb
        has_dx = bool(args[4])
        return b + (4.0 * vb * d.Cb if has_dx else 0.0), f * (2 if has_dx else 1), tag + " bwd-fused"
    if name.startswith("nas3d_conv1x1_cat_"):
        b, f, tag = _conv_work(args[0])
        return b, f, tag + " cat"
    if name == "nas3d_umma_conv":
        b, f, tag = _conv_work(args[0])
        return b, f, tag + (" umma-T" if args[1] else " umma")
    if name in ("nas3d_umma_pack_weights", "nas3d_umma_pack_weights_batch"):
        return 0.0, 0.0, "pack"
    if name == "nas3d_affine_sum_fwd":
        n, N, V, Cc = args[0], args[9], args[10], args[11]
        return 4.0 * (n + 1) * N * V * Cc, 0.0, "K=%d C=%d V=%d" % (n, Cc, V)
    if name == "nas3d_affine_sum_fwd_gn":
        n, N, V, Cc = args[0], args[15], args[16], args[17]
        return 4.0 * (n + 1) * N * V * Cc, 0.0, "K=%d C=%d V=%d" % (n, Cc, V)
    if name == "nas3d_affine_sum_bwd_apply_gn":
        n, N, V, Cc = args[0], args[24], args[25], args[26]
        return 4.0 * (2 * n + 1) * N * V * Cc, 0.0, "K=%d C=%d V=%d" % (n, Cc, V)
    if name == "nas3d_affine_sum_bwd_reduce":
        n, N, V, Cc = args[0], args[9], args[10], args[11]
        return 4.0 * (n + 1) * N * V * Cc, 0.0, "K=%d C=%d V=%d" % (n, Cc, V)
    if name == "nas3d_affine_sum_bwd_apply":
        n, N, V, Cc = args[0], args[15], args[16], args[17]
        # reads dout + x_k (when a mask/q term needs it; counted always), writes dx_k
        return 4.0 * (2 * n + 1) * N * V * Cc, 0.0, "K=%d C=%d V=%d" % (n, Cc, V)
    if name == "nas3d_moments_nc":
        N, V, Cc = args[1], args[2], args[3]
        return 4.0 * N * V * Cc, 0.0, "C=%d V=%d" % (Cc, V)
    if name == "nas3d_pool2_fwd":
        N, Do, Ho, Wo, Cc = args[5:10]
        return 4.0 * 9 * N * Do * Ho * Wo * Cc, 0.0, "C=%d" % Cc
    if name == "nas3d_pool2_bwd":
        N, Do, Ho, Wo, Cc = args[8:13]
        return 4.0 * (9 + (8 if args[0] == 1 else 0)) * N * Do * Ho * Wo * Cc, 0.0, "C=%d" % Cc
    if name == "nas3d_dice_fwd":
        N, Cc, V = args[9], args[10], args[11]
        return 8.0 * N * Cc * V, 0.0, ""
    if name == "nas3d_dice_bwd":
        N, Cc, V = args[11], args[12], args[13]
        return 8.0 * N * Cc * V, 0.0, ""
    if name == "nas3d_sigmoid_bwd":
        return 12.0 * args[3], 0.0, ""
    if name == "nas3d_add_inplace":
        return 12.0 * args[2], 0.0, ""
    if name == "nas3d_ncdhw_to_ndhwc":
        return 8.0 * args[2] * args[3] * args[4], 0.0, ""
    if name == "nas3d_adam_step":
        return 28.0 * args[4], 0.0, ""
    return 0.0, 0.0, ""


class ProfiledLib:
    """proxy around the ctypes library that brackets every call with CUDA events"""

    def __init__(self, lib):
        self._lib = lib
        self.records = []   # (name, tag, bytes, flops, ev0, ev1)

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("nas3d_") or name in ("nas3d_last_error", "nas3d_version",
                                                       "nas3d_launch_count", "nas3d_launch_count_of",
                                                       "nas3d_launch_labels", "nas3d_set_option",
                                                       "nas3d_get_option", "nas3d_probe_fma",
                                                       "nas3d_conv_wgrad_workspace_floats",
                                                       "nas3d_umma_packed_floats",
                                                       "nas3d_umma_pack_mode",
                                                       "nas3d_conv1x1_bwd_fused_supported"):
            return fn

        def wrapped(*args):
            byts, flops, tag = _work(name, args)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, tag, byts, flops, e0, e1))
            return rc
        return wrapped

    def summary(self):
        torch.cuda.synchronize()
        agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
        for name, tag, byts, flops, e0, e1 in self.records:
            a = agg[(name, tag)]
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += byts
            a[3] += flops
        rows = []
        for (name, tag), (cnt, ms, byts, flops) in agg.items():
            rows.append({"kernel": name, "shape": tag, "launches": cnt, "ms": ms,
                         "bytes": byts, "flops": flops})
        rows.sort(key=lambda r: -r["ms"])
        return rows


_active = None


def enable():
    """route every engine launch through a ProfiledLib; returns it"""
    global _active
    _active = ProfiledLib(_lib.load())
    return _active


def disable():
    global _active
    _active = None


def active():
    return _active
