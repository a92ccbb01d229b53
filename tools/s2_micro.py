"""micro-benchmark of the stride-2 3x3x3 family at the searched net's shapes (batch 8, 128^3):
weight gradient (TMA double-buffered vs cp.async kernel), forward (small-from-big) and data
gradient (big-from-small).  usage: python tools/s2_micro.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import _lib
from nas_3d_unet_b200._lib import ConvDesc
lib = _lib.load()
FMA_ROOF = None


def timed(call, reps=10):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def desc(cb, cs, S, N):
    d = ConvDesc()
    d.N = N; d.Db = d.Hb = d.Wb = S; d.Cb = cb; d.ld_big = cb
    d.Ds = d.Hs = d.Ws = S // 2; d.Cs = cs; d.ld_small = cs
    d.k, d.stride, d.dil, d.pad, d.depthwise = 3, 2, 1, 1, 0
    return d


def run(cb, cs, S, N=8):
    d = desc(cb, cs, S, N)
    so = S // 2
    big = torch.randn(N, S, S, S, cb, device="cuda"); small = torch.randn(N, so, so, so, cs, device="cuda")
    w = torch.randn(cs, cb, 27, device="cuda"); bias = torch.randn(cs, device="cuda")
    dW = torch.zeros(cs, cb, 27, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gflop = 2.0 * N * so ** 3 * 27 * cb * cs * 1e-9
    out = {}

    def wg():
        _lib.check(lib.nas3d_conv_wgrad_ws(C.byref(d), small.data_ptr(), big.data_ptr(), None, 0, dW.data_ptr(), None, None, None, 0, st), "wgrad")
    for name, v in (("wgrad_tma", 1), ("wgrad_cp", 0)):
        with _lib.option("s2_wgrad_tma", v):
            out[name] = timed(wg)
    def fwd():
        _lib.check(lib.nas3d_conv_small_from_big(C.byref(d), big.data_ptr(), w.data_ptr(), bias.data_ptr(), None, 0, 0, small.data_ptr(), 0, None, st), "fwd")
    def dgrad():
        _lib.check(lib.nas3d_conv_big_from_small(C.byref(d), small.data_ptr(), w.data_ptr(), None, None, 0, None, big.data_ptr(), 0, None, st), "dgrad")
    out["fwd"] = timed(fwd)
    if cb == cs:
        out["dgrad"] = timed(dgrad)
    print("C%d->%d @%d^3 N%d  %.2f GFLOP   " % (cb, cs, S, N, gflop) +
          "   ".join("%s %.1f us (%.1f TF/s)" % (k, t, gflop / t * 1e3) for k, t in out.items()), flush=True)


def run_dw(c, S, N=8):
    d = desc(c, c, S, N)
    d.depthwise = 1
    so = S // 2
    big = torch.randn(N, S, S, S, c, device="cuda"); small = torch.randn(N, so, so, so, c, device="cuda")
    dW = torch.zeros(c, 27, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    mb = 4e-6 * (big.numel() + small.numel())
    out = {}

    def wg():
        _lib.check(lib.nas3d_conv_wgrad_ws(C.byref(d), small.data_ptr(), big.data_ptr(), None, 0, dW.data_ptr(), None, None, None, 0, st), "dw wgrad")
    for name, v in (("dw_wgrad_tma", 1), ("dw_wgrad_cp", 0)):
        with _lib.option("s2_wgrad_tma", v):
            out[name] = timed(wg)
    w = torch.randn(c, 27, device="cuda")

    def fwd():
        _lib.check(lib.nas3d_conv_small_from_big(C.byref(d), big.data_ptr(), w.data_ptr(), None, None, 0, 0, small.data_ptr(), 0, None, st), "dw fwd")

    def dgrad():
        _lib.check(lib.nas3d_conv_big_from_small(C.byref(d), small.data_ptr(), w.data_ptr(), None, None, 0, None, big.data_ptr(), 0, None, st), "dw dgrad")
    out["dw_fwd"] = timed(fwd)
    out["dw_dgrad"] = timed(dgrad)
    print("depthwise C%d @%d^3 N%d  %.0f MB   " % (c, S, N, mb) +
          "   ".join("%s %.1f us (%.0f GB/s)" % (k, t, mb / t * 1e3) for k, t in out.items()), flush=True)


if __name__ == "__main__":
    run_dw(4, 128)
    run_dw(8, 64)
    run_dw(16, 32)
    if "--dw-only" in sys.argv:
        sys.exit(0)
    for cb, cs, S in ((4, 4, 128), (4, 12, 128), (8, 8, 64), (16, 16, 32)):
        run(cb, cs, S)
