"""nas3d_umma_conv straight through the C-ABI, CUDA-event timed back to back (no module overhead):
warp-specialised vs lock-step kernel at the searched net's deep-level shapes (batch 8)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import _lib
from nas_3d_unet_b200._lib import ConvDesc
lib = _lib.load()


def run(c, S, stride=1, dil=1, produce_big=0, N=8, reps=20):
    pad = dil if stride == 1 else 1
    so = (S + 2 * pad - dil * 2 - 1) // stride + 1
    d = ConvDesc(N, S, S, S, c, c, so, so, so, c, c, 3, stride, dil, pad, 0)
    big = torch.randn(N, S, S, S, c, device="cuda"); small = torch.randn(N, so, so, so, c, device="cuda")
    w = torch.randn(c, c, 27, device="cuda") * 0.1
    n = lib.nas3d_umma_packed_floats(C.byref(d), produce_big)
    wp = torch.empty(n, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.nas3d_umma_pack_weights(C.byref(d), w.data_ptr(), produce_big, wp.data_ptr(), st), "pack")
    src, dst = (small, big) if produce_big else (big, small)
    out = {}
    for name, v in (("ws", 1), ("lockstep", 0)):
        with _lib.option("umma_ws", v):
            def call():
                _lib.check(lib.nas3d_umma_conv(C.byref(d), produce_big, src.data_ptr(), wp.data_ptr(), None,
                                               dst.data_ptr(), 0, None, st), "umma_conv")
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                call()
            e1.record(); torch.cuda.synchronize()
            out[name] = e0.elapsed_time(e1) / reps * 1e3
    gflop = 2.0 * N * so ** 3 * 27 * c * c * 1e-9
    print("C%d @%d^3 s%d d%d %s: %.2f GFLOP  " % (c, S, stride, dil, "dgrad" if produce_big else "fwd", gflop) +
          "  ".join("%s %.1f us (%.1f TF/s)" % (k, t, gflop / t * 1e3) for k, t in out.items()), flush=True)


if __name__ == "__main__":
    for c, S in ((32, 16), (64, 8), (16, 32)):
        for pb in (0, 1):
            run(c, S, produce_big=pb)
            run(c, S, dil=2, produce_big=pb)
    for c, S in ((16, 32), (32, 16), (64, 8)):
        for pb in (0, 1):
            run(c, S, stride=2, produce_big=pb)
