"""config #5: sliding-window whole-volume inference, synthetic 4x240x240x155 volumes, 128^3 patches
(9 per volume), searched-G0 net in eval mode.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nas_3d_unet_b200.searched import SearchedNet
from nas_3d_unet_b200.genotype import G0
from nas_3d_unet_b200.infer import SlidingWindowPredictor

def main():
    nvol = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.manual_seed(0)
    model = SearchedNet(4, 4, 3, 4, 3, True, G0).cuda().eval()
    rng = np.random.default_rng(0)
    vol = (rng.random((4, 240, 240, 155), dtype=np.float32) * 100 + 10)
    hv = torch.as_tensor(vol).pin_memory()
    pred = SlidingWindowPredictor(model, (128, 128, 128), batch=9)
    for _ in range(2):
        pred.predict(hv.cuda(non_blocking=True))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(nvol):
        lab = pred.predict(hv.cuda(non_blocking=True))     # H2D of the volume inside the timed loop
        lab_host = lab.cpu()                                # uint8 label volume back to the host
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / nvol
    print(json.dumps({"metric": "sliding-window inference", "volumes_per_s": 1 / dt, "patches_per_s": 9 / dt,
                      "ms_per_volume": dt * 1e3, "patch": 128, "patches_per_volume": 9,
                      "labels": sorted(int(v) for v in torch.unique(lab_host))}))

if __name__ == "__main__":
    main()
