"""Sliding-window whole-volume inference on the GPU (reference: prediction.py:121-170 with
patches.py): patch corners (host integers, bit-exact with the reference), patch extraction,
batched searched-net forward, float64 mean-stitch in patch order, label assembly, skull mask.

    pred = SlidingWindowPredictor(model, patch_shape=(128,128,128), batch=8)
    labels = pred.predict(volume, brain_width)        # uint8 (D,H,W) with values {0,1,2,4}

The reference does batch 1 with autograd on and a device->host copy per patch; here nothing
leaves the device until the uint8 label volume.  With torch.distributed initialised the patches
of a volume are sharded over the ranks and gathered on every rank before the (deterministic,
order-preserving) stitch.
"""
import numpy as np
import torch

from . import _lib
from .engine import _ndhwc_pitch, _stream, get_lib


# ---- host-side integer logic (patches.py:9-75), kept bit-exact incl. float -> int truncation ----
def _grid(start, stop, step):
    g = np.mgrid[start[0]:stop[0]:step[0], start[1]:stop[1]:step[1], start[2]:stop[2]:step[2]]
    return np.asarray(g.reshape(3, -1).T, dtype=int)


def _autofit(img_shape, patch_shape):
    n = np.ceil(img_shape / patch_shape)
    start, step = np.zeros(3), np.zeros(3)
    for d in range(3):
        if n[d] == 1:
            start[d] = -(patch_shape[d] - img_shape[d]) // 2
            step[d] = patch_shape[d]
        else:
            ov = np.floor(n[d] * patch_shape[d] - img_shape[d]) / (n[d] - 1)
            overflow = n[d] * patch_shape[d] - (n[d] - 1) * ov - img_shape[d]
            start[d] = -overflow // 2
            step[d] = patch_shape[d] - ov
    return np.vstack((_grid(start, start + n * step, step), (img_shape - patch_shape) // 2))


def patching(img_shape, patch_shape, overlap=None, both_ps=False):
    """corners of the patches covering img_shape: auto-fit grid + centre cube (overlap None), or
    centre cube + grid with the given overlap"""
    img_shape, patch_shape = np.asarray(img_shape), np.asarray(patch_shape)
    auto = _autofit(img_shape, patch_shape)
    if overlap is None:
        return auto
    overlap = np.asarray([overlap] * 3) if isinstance(overlap, int) else np.asarray(overlap)
    n = np.ceil(img_shape / (patch_shape - overlap))
    overflow = patch_shape * n - (n - 1) * overlap - img_shape
    start = -overflow // 2
    step = patch_shape - overlap
    ol = np.vstack(((img_shape - patch_shape) // 2, _grid(start, start + n * step, step)))
    return np.vstack((auto, ol)) if both_ps else ol


def seg_to_masks(seg, inclusive_label=True):
    """int16 (N,1,D,H,W) segmentation on the device -> float32 (N,3,D,H,W) region masks"""
    if not seg.is_cuda or seg.dtype != torch.int16:
        raise TypeError("seg_to_masks expects a CUDA int16 tensor")
    seg = seg.contiguous()
    N = seg.shape[0]
    V = seg[0].numel()
    out = torch.empty((N, 3) + tuple(seg.shape[2:]), device=seg.device, dtype=torch.float32)
    _lib.check(get_lib().nas3d_seg_to_masks(seg.data_ptr(), N, V, 1 if inclusive_label else 0,
                                            out.data_ptr(), _stream()), "seg_to_masks")
    return out


class SlidingWindowPredictor:
    def __init__(self, model, patch_shape=(128, 128, 128), batch=8, threshold=0.5,
                 inclusive_label=True, patch_overlap=None):
        self.model = model
        self.patch_shape = tuple(int(p) for p in patch_shape)
        self.batch = int(batch)
        self.threshold = float(threshold)
        self.inclusive = bool(inclusive_label)
        self.overlap = patch_overlap

    @torch.no_grad()
    def predict_patches(self, volume, corners):
        """volume: (C,D,H,W) float32 CUDA; corners: (B,3) ints in volume coordinates.
        Returns predictions as a (B,3,Pd,Ph,Pw) tensor (channels-last memory)."""
        lib = get_lib()
        C, D, H, W = volume.shape
        Pd, Ph, Pw = self.patch_shape
        dev = volume.device
        cdev = torch.as_tensor(np.ascontiguousarray(corners, dtype=np.int32), device=dev)
        B = cdev.shape[0]
        ld = (C + 3) // 4 * 4
        outs = []
        was_training = self.model.training
        self.model.eval()
        try:
            for b0 in range(0, B, self.batch):
                nb = min(self.batch, B - b0)
                x = torch.empty((nb, Pd, Ph, Pw, ld), device=dev, dtype=torch.float32)
                _lib.check(lib.nas3d_extract_patches(volume.data_ptr(), C, D, H, W,
                                                     cdev[b0:b0 + nb].data_ptr(), nb, Pd, Ph, Pw,
                                                     x.data_ptr(), ld, _stream()), "extract_patches")
                y = self.model(x.permute(0, 4, 1, 2, 3)[:, :C])
                # an all-zero patch is not run through the net by the reference: its prediction
                # is defined as 0 (prediction.py:133-136)
                empty = ~x.reshape(nb, -1).ne(0).any(dim=1)
                if bool(empty.any()):
                    y = y * (~empty).to(y.dtype).view(nb, 1, 1, 1, 1)
                outs.append(y)
        finally:
            self.model.train(was_training)
        return torch.cat(outs, 0) if len(outs) > 1 else outs[0]

    @torch.no_grad()
    def predict(self, volume, brain_width=None, skull_mask=None, return_stitched=False):
        """volume (C,D,H,W) float32 (CUDA tensor or numpy); brain_width [[d0,h0,w0],[d1,h1,w1]]
        inclusive bounds (default: whole volume).  Returns uint8 labels (D,H,W) on the device."""
        import torch.distributed as dist
        lib = get_lib()
        if isinstance(volume, np.ndarray):
            volume = torch.as_tensor(volume, dtype=torch.float32).cuda()
        volume = volume.contiguous()
        C, D, H, W = volume.shape
        bw = np.asarray(brain_width if brain_width is not None else [[0, 0, 0], [D - 1, H - 1, W - 1]])
        off = bw[0].astype(int)
        bshape = (bw[1] - bw[0] + 1).astype(int)
        corners = patching(bshape, self.patch_shape, overlap=self.overlap)    # brain coordinates
        B = len(corners)
        vol_corners = corners + off[None, :]
        # patches outside the brain box are zero-padded exactly as get_data_from_file crops first:
        brain = volume[:, off[0]:off[0] + bshape[0], off[1]:off[1] + bshape[1],
                       off[2]:off[2] + bshape[2]].contiguous()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        mine = list(range(rank, B, world))
        preds_mine = self.predict_patches(brain, corners[mine]) if mine else None
        Pd, Ph, Pw = self.patch_shape
        ld_pred = 3
        if world > 1:
            per = (B + world - 1) // world
            buf = torch.zeros((per, Pd, Ph, Pw, 3), device=volume.device, dtype=torch.float32)
            if mine:
                buf[:len(mine)] = preds_mine.permute(0, 2, 3, 4, 1)
            allb = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(allb, buf)
            preds = torch.empty((B, Pd, Ph, Pw, 3), device=volume.device, dtype=torch.float32)
            for r in range(world):
                idx = list(range(r, B, world))
                if idx:
                    preds[idx] = allb[r][:len(idx)]
        else:
            ld_pred = _ndhwc_pitch(preds_mine)              # the head output is stored at pitch 4
            preds = preds_mine.permute(0, 2, 3, 4, 1)       # (B,Pd,Ph,Pw,3) view of the NDHWC data
            if ld_pred is None:
                preds = preds.contiguous()
                ld_pred = 3
        cdev = torch.as_tensor(np.ascontiguousarray(corners, dtype=np.int32), device=volume.device)
        labels = torch.empty((D, H, W), device=volume.device, dtype=torch.uint8)
        stitched = (torch.empty((3,) + tuple(int(b) for b in bshape), device=volume.device,
                                dtype=torch.float64) if return_stitched else None)
        skull = None
        if skull_mask is not None:
            skull = torch.as_tensor(skull_mask).to(volume.device).to(torch.uint8).contiguous()
        _lib.check(lib.nas3d_stitch_labels(
            preds.data_ptr(), ld_pred, cdev.data_ptr(), B, Pd, Ph, Pw, int(bshape[0]), int(bshape[1]),
            int(bshape[2]), D, H, W, int(off[0]), int(off[1]), int(off[2]), self.threshold,
            1 if self.inclusive else 0, skull.data_ptr() if skull is not None else None,
            labels.data_ptr(), stitched.data_ptr() if stitched is not None else None, _stream()),
            "stitch_labels")
        del vol_corners
        return (labels, stitched, preds, corners) if return_stitched else labels
