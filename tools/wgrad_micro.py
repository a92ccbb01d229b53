"""micro-benchmark: nas3d_conv_wgrad_ws on the deep-level shapes of the searched net, tcgen05
split-K GEMM (conv_umma_wgrad.cu) against the CUDA-core kernels (library option umma_wgrad)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import _lib
from nas_3d_unet_b200._lib import ConvDesc
lib = _lib.load()
def run(c, S, N=8, stride=1):
    d = ConvDesc()
    d.N = N; d.Db = d.Hb = d.Wb = S; d.Cb = c; d.ld_big = c
    so = S // stride
    d.Ds = d.Hs = d.Ws = so; d.Cs = c; d.ld_small = c
    d.k, d.stride, d.dil, d.pad, d.depthwise = 3, stride, 1, 1, 0
    big = torch.randn(N, S, S, S, c, device="cuda"); small = torch.randn(N, so, so, so, c, device="cuda")
    dW = torch.zeros(c, c, 27, device="cuda")
    n = lib.nas3d_conv_wgrad_workspace_floats(C.byref(d), 0)
    ws = torch.empty(max(n, 1), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def call():
        _lib.check(lib.nas3d_conv_wgrad_ws(C.byref(d), small.data_ptr(), big.data_ptr(), None, 0, dW.data_ptr(), None, None, ws.data_ptr() if n else None, n, st), "wgrad")
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3
with _lib.option("umma_wgrad", 1), _lib.option("umma_wgrad_min_c", 16):
    print("tcgen05: C16@32^3: %.1f us  C32@16^3: %.1f us  C64@8^3: %.1f us  C16 s2 @32->16: %.1f us" % (run(16, 32), run(32, 16), run(64, 8), run(16, 32, stride=2)))
with _lib.option("umma_wgrad", 0):
    print("ffma: C16@32^3: %.1f us  C32@16^3: %.1f us  C64@8^3: %.1f us  C16 s2: %.1f us" % (run(16, 32), run(32, 16), run(64, 8), run(16, 32, stride=2)))
