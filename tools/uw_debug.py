"""debug aid: nas3d_conv_wgrad on constant inputs, tcgen05 path vs CUDA-core path"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import _lib
from nas_3d_unet_b200._lib import ConvDesc

lib = _lib.load()
c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, S = 2, 6
d = ConvDesc()
d.N = N; d.Db = d.Hb = d.Wb = S; d.Cb = c; d.ld_big = c
d.Ds = d.Hs = d.Ws = S; d.Cs = c; d.ld_small = c
d.k, d.stride, d.dil, d.pad, d.depthwise = 3, 1, 1, 1, 0
torch.manual_seed(0)
big = torch.randn(N, S, S, S, c, device="cuda")
small = torch.randn(N, S, S, S, c, device="cuda")
res = {}
for mode in (0, 1):
    dW = torch.zeros(c, c, 27, device="cuda")
    with _lib.option("umma_wgrad", mode), _lib.option("umma_wgrad_debug", mode):
        _lib.check(lib.nas3d_conv_wgrad(C.byref(d), small.data_ptr(), big.data_ptr(), None, 0, dW.data_ptr(),
                                        None, None, torch.cuda.current_stream().cuda_stream), "wgrad")
    torch.cuda.synchronize()
    res[mode] = dW.cpu()
    print("mode", mode, "nonzero", int((dW != 0).sum()), "sum", float(dW.sum()), "abs", float(dW.abs().sum()))
print("first entries ffma ", res[0].reshape(-1)[:6].tolist())
print("first entries umma ", res[1].reshape(-1)[:6].tolist())
print("max abs diff", float((res[0] - res[1]).abs().max()), "of", float(res[0].abs().max()))
# which (cs, cb, tap) entries agree?
ok = (res[0] - res[1]).abs() <= 1e-3 * res[0].abs().max()
print("agreeing entries", int(ok.sum()), "of", ok.numel())
print(_lib.launch_counts())
