"""kernel-variant launch counts of ONE supernet forward+backward (or searched-net step), from the
library's own counters (nas3d_launch_labels).  usage: python tools/launch_mix.py [supernet|searched] [patch] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nas_3d_unet_b200 import _lib
from nas_3d_unet_b200.loss import WeightedDiceLoss

wl = sys.argv[1] if len(sys.argv) > 1 else "supernet"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
if wl == "supernet":
    from nas_3d_unet_b200.nas import ShellNet
    model = ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True).to(dev)
else:
    from nas_3d_unet_b200.searched import SearchedNet
    from nas_3d_unet_b200.genotype import G0
    model = SearchedNet(4, 4, 3, 4, 3, True, G0).to(dev)
model.train()
lossf = WeightedDiceLoss().to(dev)
x = torch.randn(B, 4, P, P, P, device=dev)
y = (torch.rand(B, 3, P, P, P, device=dev) > 0.5).float()
for it in range(2):
    before = _lib.launch_counts()
    model.zero_grad()
    loss = lossf(model(x), y)
    mid = _lib.launch_counts()
    loss.backward()
    torch.cuda.synchronize()
    after = _lib.launch_counts()
fw = {k: mid.get(k, 0) - before.get(k, 0) for k in mid}
bw = {k: after.get(k, 0) - mid.get(k, 0) for k in after}
print("%s patch %d batch %d: forward %d launches, backward %d" % (wl, P, B, sum(fw.values()), sum(bw.values())))
for k in sorted(after, key=lambda k: -(fw.get(k, 0) + bw.get(k, 0))):
    if fw.get(k, 0) + bw.get(k, 0):
        print("  %-28s fwd %4d  bwd %4d" % (k, fw.get(k, 0), bw.get(k, 0)))
