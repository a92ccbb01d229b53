"""Micro-benchmark of single 3x3x3 convolution launches straight through the C-ABI
(nas3d_conv_small_from_big / big_from_small / wgrad): a short process for ncu captures and for
timing kernel variants without running a whole training step.

    python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1 [--stride 1] [--iters 10]
prints one JSON line per direction (us per launch, TFLOP/s, algorithmic GB/s)."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nas_3d_unet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c", type=int, default=4)
    ap.add_argument("--cs", type=int, default=0, help="small-side channels (default: = --c)")
    ap.add_argument("--s", type=int, default=128)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--dil", type=int, default=1)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--relu", type=int, default=0, help="k=1: relu prologue (preprocess convs)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--which", default="fwd,dgrad,wgrad")
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    cs = a.cs or a.c
    k = a.k
    pad = 0 if k == 1 else (a.dil if a.stride == 1 else (1 if a.dil == 1 else 2))
    S = a.s
    Ss = (S + 2 * pad - a.dil * (k - 1) - 1) // a.stride + 1
    d = _lib.ConvDesc(a.n, S, S, S, a.c, a.c, Ss, Ss, Ss, cs, cs, k, a.stride, a.dil, pad, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    big = torch.randn(a.n * S ** 3 * a.c, device=dev, generator=g)
    small = torch.randn(a.n * Ss ** 3 * cs, device=dev, generator=g)
    w = torch.randn(cs * a.c * k ** 3, device=dev, generator=g) * 0.1
    bias = torch.randn(cs, device=dev, generator=g)
    mask = torch.randn(a.n * S ** 3 * a.c, device=dev, generator=g) if a.relu else None
    dW = torch.zeros_like(w)
    db = torch.zeros(cs, device=dev)
    mom = torch.zeros(a.n * max(cs, a.c) * 2, device=dev, dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    flops = 2.0 * a.n * Ss ** 3 * cs * a.c * k ** 3
    bytes_ = 4.0 * (big.numel() + small.numel())

    def fwd():
        _lib.check(lib.nas3d_conv_small_from_big(C.byref(d), big.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                                 None, a.relu, 0, small.data_ptr(), 0, mom.data_ptr(), st), "fwd")

    def dgrad():
        _lib.check(lib.nas3d_conv_big_from_small(C.byref(d), small.data_ptr(), w.data_ptr(), None,
                                                 mask.data_ptr() if a.relu else None, a.c, None,
                                                 big.data_ptr(), 0, None, st), "dgrad")

    def wgrad():
        _lib.check(lib.nas3d_conv_wgrad(C.byref(d), small.data_ptr(), big.data_ptr(), None, a.relu, dW.data_ptr(),
                                        db.data_ptr(), None, st), "wgrad")

    for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
        if name not in a.which.split(","):
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.iters * 1e3
        print(json.dumps({"dir": name, "C": a.c, "Cs": cs, "S": S, "N": a.n, "k": k, "dil": a.dil, "stride": a.stride,
                          "us": round(us, 1), "TFLOPs": round(flops / us / 1e6, 2),
                          "GBps": round(bytes_ / us / 1e3, 1)}))


if __name__ == "__main__":
    main()
