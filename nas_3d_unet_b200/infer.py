"""Sliding-window whole-volume inference on the GPU (reference: prediction.py:121-170 with
patches.py): patch corners (host integers, bit-exact with the reference), patch extraction,
batched searched-net forward, float64 mean-stitch in patch order, label assembly, skull mask.

    pred = SlidingWindowPredictor(model, patch_shape=(128,128,128), batch=8)
    labels = pred.predict(volume, brain_width)        # uint8 (D,H,W) with values {0,1,2,4}

The reference does batch 1 with autograd on and a device->host copy per patch; here nothing
leaves the device until the uint8 label volume and the patch loop has no host synchronisation
(the reference's `np.all(data==0)` skip is a device flag consumed by the stitch kernel).

Multi-GPU (one process per GPU): `predict_many` shards whole VOLUMES over the ranks when
there are at least as many volumes as ranks (no communication); `predict` shards the PATCHES of one
volume round-robin over the ranks and all-gathers the predictions, so every rank runs the same
deterministic, order-preserving float64 stitch and holds the same labels.
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


def shard_indices(n, rank, world):
    """units (patches of a volume, or volumes of a list) owned by `rank`: round-robin"""
    return list(range(rank, n, world))


def gather_patch_predictions(local, n_total, group=None):
    """local: (n_local, ...) tensor of the predictions of THIS rank's patches (shard_indices order).
    Returns the n_total per-patch tensors in patch order, identical on every rank.  One
    all_gather of equal-sized (zero-padded) buffers; works on any backend / device (the CPU
    tests run it over gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    if local.shape[0] == per:
        buf = local.contiguous()
    else:
        buf = local.new_zeros((per,) + tuple(local.shape[1:]))
        buf[:local.shape[0]] = local
    allb = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf, group=group)
    out = [None] * n_total
    for r in range(world):
        for j, b in enumerate(shard_indices(n_total, r, world)):
            out[b] = allb[r][j]
    return out


class SlidingWindowPredictor:
    def __init__(self, model, patch_shape=(128, 128, 128), batch=8, threshold=0.5,
                 inclusive_label=True, patch_overlap=None, group=None, deterministic=False):
        self.model = model
        self.patch_shape = tuple(int(p) for p in patch_shape)
        self.batch = int(batch)
        self.threshold = float(threshold)
        self.inclusive = bool(inclusive_label)
        self.overlap = patch_overlap
        self.group = group
        # the reference's CPU inference is bit-reproducible; ours is too once the tcgen05 convs of the
        # deep levels stop splitting their taps over CTAs (fp32 atomics reorder sums at the 1e-7 level,
        # which flips a handful of labels per volume where a probability sits on the threshold).
        # deterministic=True costs ~10 % of a volume's time (18.0 -> 20.0 ms at 240x240x155, measured)
        self.deterministic = bool(deterministic)

    @torch.no_grad()
    def predict_patches(self, volume, corners):
        """volume: (C,D,H,W) float32 CUDA; corners: (B,3) ints in volume coordinates.
        Returns (chunks, ld): the raw network outputs, one (nb,3,Pd,Ph,Pw) tensor (NDHWC memory,
        voxel pitch ld) per forward batch.  All-zero patches are NOT special-cased here
        (predict() hands their flags to the stitch kernel)."""
        lib = get_lib()
        C, D, H, W = volume.shape
        Pd, Ph, Pw = self.patch_shape
        dev = volume.device
        cdev = torch.as_tensor(np.ascontiguousarray(corners, dtype=np.int32), device=dev)
        B = cdev.shape[0]
        ld = (C + 3) // 4 * 4
        chunks, ld_pred = [], None
        was_training = self.model.training
        self.model.eval()
        split_k = _lib.option("umma_split_k", 0 if self.deterministic else 1)
        split_k.__enter__()
        try:
            for b0 in range(0, B, self.batch):
                nb = min(self.batch, B - b0)
                x = torch.empty((nb, Pd, Ph, Pw, ld), device=dev, dtype=torch.float32)
                _lib.check(lib.nas3d_extract_patches(volume.data_ptr(), C, D, H, W,
                                                     cdev[b0:b0 + nb].data_ptr(), nb, Pd, Ph, Pw,
                                                     x.data_ptr(), ld, _stream()), "extract_patches")
                y = self.model(x.permute(0, 4, 1, 2, 3)[:, :C])
                l = _ndhwc_pitch(y)
                if l is None or (ld_pred is not None and l != ld_pred):
                    y = y.contiguous(memory_format=torch.channels_last_3d)
                    l = 3
                ld_pred = l
                chunks.append(y)
        finally:
            split_k.__exit__(None, None, None)
            self.model.train(was_training)
        return chunks, ld_pred

    @torch.no_grad()
    def predict(self, volume, brain_width=None, skull_mask=None, return_stitched=False,
                shard_patches=True):
        """volume (C,D,H,W) float32 (CUDA tensor or numpy); brain_width [[d0,h0,w0],[d1,h1,w1]]
        inclusive bounds (default: whole volume).  Returns uint8 labels (D,H,W) on the device.
        With torch.distributed initialised (and shard_patches) the patches are sharded over the
        ranks of `group`; every rank returns the same labels."""
        import torch.distributed as dist
        lib = get_lib()
        if isinstance(volume, np.ndarray):
            volume = torch.as_tensor(volume, dtype=torch.float32).cuda()
        volume = volume.contiguous()
        dev = volume.device
        C, D, H, W = volume.shape
        bw = np.asarray(brain_width if brain_width is not None else [[0, 0, 0], [D - 1, H - 1, W - 1]])
        off = bw[0].astype(int)
        bshape = (bw[1] - bw[0] + 1).astype(int)
        corners = patching(bshape, self.patch_shape, overlap=self.overlap)    # brain coordinates
        B = len(corners)
        Pd, Ph, Pw = self.patch_shape
        # patches outside the brain box are zero-padded exactly as get_data_from_file crops first:
        brain = volume[:, off[0]:off[0] + bshape[0], off[1]:off[1] + bshape[1],
                       off[2]:off[2] + bshape[2]].contiguous()
        cdev = torch.as_tensor(np.ascontiguousarray(corners, dtype=np.int32), device=dev)
        # prediction.py:133-136: an all-zero patch is never run through the net, its prediction is
        # zeros.  Here: a device flag per patch (every rank computes all of them - one pass over
        # the volume), consumed by the stitch kernel; no host round trip.
        nonzero = torch.empty((B,), device=dev, dtype=torch.int32)
        _lib.check(lib.nas3d_patch_nonzero(brain.data_ptr(), C, int(bshape[0]), int(bshape[1]),
                                           int(bshape[2]), cdev.data_ptr(), B, Pd, Ph, Pw,
                                           nonzero.data_ptr(), _stream()), "patch_nonzero")
        world = (dist.get_world_size(self.group)
                 if shard_patches and dist.is_available() and dist.is_initialized() else 1)
        rank = dist.get_rank(self.group) if world > 1 else 0
        mine = shard_indices(B, rank, world)
        if mine:
            chunks, ld_pred = self.predict_patches(brain, corners[mine])
        else:
            chunks, ld_pred = [], 4
        if world > 1:
            # every rank must agree on the pitch: the gathered buffers are (.., 4) NDHWC patches
            local = torch.zeros((len(mine), Pd, Ph, Pw, 4), device=dev, dtype=torch.float32)
            j = 0
            for y in chunks:
                local[j:j + y.shape[0], ..., :3] = y.permute(0, 2, 3, 4, 1)
                j += y.shape[0]
            per_patch = gather_patch_predictions(local, B, self.group)
            ld_pred = 4
        else:
            per_patch = [y[i] for y in chunks for i in range(y.shape[0])]
        keep = (chunks, per_patch)            # the pointer table below does not own its targets
        table = torch.tensor([t.data_ptr() for t in per_patch], dtype=torch.int64).to(dev, non_blocking=True)
        labels = torch.empty((D, H, W), device=dev, dtype=torch.uint8)
        stitched = (torch.empty((3,) + tuple(int(b) for b in bshape), device=dev, dtype=torch.float64)
                    if return_stitched else None)
        skull = None
        if skull_mask is not None:
            skull = torch.as_tensor(skull_mask).to(dev).to(torch.uint8).contiguous()
        _lib.check(lib.nas3d_stitch_labels(
            table.data_ptr(), nonzero.data_ptr(), ld_pred, cdev.data_ptr(), B, Pd, Ph, Pw,
            int(bshape[0]), int(bshape[1]), int(bshape[2]), D, H, W, int(off[0]), int(off[1]),
            int(off[2]), self.threshold, 1 if self.inclusive else 0,
            skull.data_ptr() if skull is not None else None, labels.data_ptr(),
            stitched.data_ptr() if stitched is not None else None, _stream()), "stitch_labels")
        if not return_stitched:
            del keep
            return labels
        # for the tests: the predictions as the reference would hold them (zeros for empty patches)
        preds = torch.stack([t[..., :3] if world > 1 else t.permute(1, 2, 3, 0) for t in per_patch])
        preds = preds * nonzero.ne(0).to(preds.dtype).view(B, 1, 1, 1, 1)
        return labels, stitched, preds, corners

    @torch.no_grad()
    def predict_many(self, volumes, brain_widths=None, skull_masks=None):
        """a list of volumes: sharded over the ranks volume by volume when there are at least as
        many volumes as ranks (each rank runs whole volumes, no communication; returns
        {volume index: labels} for the volumes THIS rank owns), otherwise patch-sharded one volume
        at a time (every rank returns every volume's labels)."""
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        n = len(volumes)
        by_volume = n >= world
        out = {}
        for i in (shard_indices(n, rank, world) if by_volume else range(n)):
            out[i] = self.predict(volumes[i], None if brain_widths is None else brain_widths[i],
                                  None if skull_masks is None else skull_masks[i],
                                  shard_patches=not by_volume)
        return out
