"""micro-benchmark of single convolution calls through the module surface (tuning aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import config
from nas_3d_unet_b200.prim_ops import ConvOps
config.cfg.umma_min_c = 16

def bench(c, s, n=8, stride=1, dil=1, transposed=False, iters=20):
    op = ConvOps(c, c, stride=stride, dilation=dil, transposed=transposed, ops_order='weight').cuda()
    x = torch.randn(n, c, s, s, s, device='cuda').contiguous(memory_format=torch.channels_last_3d)
    res = {}
    for mode in ("umma", "ffma"):
        config.cfg.umma = (mode == "umma")
        with torch.no_grad():
            for _ in range(3):
                y = op(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                y = op(x)
            e1.record()
            torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / iters * 1e3
    config.cfg.umma = True
    flops = 2.0 * y.numel() * c * 27 if not transposed else 2.0 * x.numel() * c * 27
    print("C=%d S=%d stride=%d dil=%d T=%d: umma %.1f us (%.1f TF)  ffma %.1f us (%.1f TF)" % (
        c, s, stride, dil, transposed, res["umma"], flops / res["umma"] / 1e6, res["ffma"], flops / res["ffma"] / 1e6))

if __name__ == "__main__":
    for c, s in ((16, 32), (32, 16), (64, 8), (64, 16), (32, 32), (64, 32)):
        bench(c, s)
    bench(64, 8, stride=2)
    bench(16, 32, stride=2)
    bench(16, 16, stride=2, transposed=True)
