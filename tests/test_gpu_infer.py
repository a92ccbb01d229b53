"""GPU: sliding-window inference chain (patch extraction, stitch, label assembly) - bit-exact
against the numpy oracle of patches.py / prediction.py semantics."""
import numpy as np
import pytest
import torch

from oracle import nas3d_oracle as O
from helpers import make_searched

pytestmark = pytest.mark.gpu


def _volume(seed, shape=(4, 40, 44, 36)):
    rng = np.random.default_rng(seed)
    v = (rng.random(shape, dtype=np.float32) * 100 + 10).astype(np.float32)
    zz, yy, xx = np.meshgrid(*[np.arange(s) for s in shape[1:]], indexing="ij")
    c = [(s - 1) / 2 for s in shape[1:]]
    r2 = ((zz - c[0]) / (shape[1] / 2)) ** 2 + ((yy - c[1]) / (shape[2] / 2)) ** 2 + ((xx - c[2]) / (shape[3] / 2)) ** 2
    return v * (r2 < 0.8)[None].astype(np.float32)


def test_extract_patches_bit_exact(lib):
    from nas_3d_unet_b200 import _lib
    from nas_3d_unet_b200.engine import _stream
    vol = _volume(1)
    corners = np.array([[-3, 5, 2], [20, 30, 10], [0, 0, 0], [-40, 0, 0]], dtype=np.int32)
    P = (16, 18, 20)
    vg = torch.as_tensor(vol).cuda()
    cg = torch.as_tensor(corners).cuda()
    out = torch.empty((len(corners),) + P + (4,), device='cuda')
    _lib.check(lib.nas3d_extract_patches(vg.data_ptr(), 4, *vol.shape[1:], cg.data_ptr(), len(corners),
                                         *P, out.data_ptr(), 4, _stream()), "extract")
    got = out.permute(0, 4, 1, 2, 3).cpu().numpy()
    for b, c in enumerate(corners):
        np.testing.assert_array_equal(got[b], O.get_patch(vol, P, c))


@pytest.mark.parametrize("inclusive", [True, False])
def test_sliding_window_chain_matches_oracle(inclusive):
    from nas_3d_unet_b200.infer import SlidingWindowPredictor
    model = make_searched().cuda()
    vol = _volume(2)
    bw = np.array([[3, 2, 1], [37, 41, 33]])
    skull = (vol.sum(0) != 0).astype(np.uint8)
    pred = SlidingWindowPredictor(model, patch_shape=(32, 32, 32), batch=4, inclusive_label=inclusive)
    labels, stitched, preds, corners = pred.predict(torch.as_tensor(vol).cuda(), bw, skull_mask=skull,
                                                    return_stitched=True)
    bshape = tuple(bw[1] - bw[0] + 1)
    np.testing.assert_array_equal(corners, O.patching(bshape, (32, 32, 32)))
    # the device stitch on the device predictions == numpy float64 stitch on the same predictions
    pl = [preds[b].permute(3, 0, 1, 2).cpu().numpy() for b in range(preds.shape[0])]
    ref_st = O.stitch(pl, corners, (3,) + bshape)
    np.testing.assert_array_equal(stitched.cpu().numpy(), ref_st)
    full = np.zeros((3,) + vol.shape[1:])
    full[:, bw[0, 0]:bw[1, 0] + 1, bw[0, 1]:bw[1, 1] + 1, bw[0, 2]:bw[1, 2] + 1] = ref_st
    ref_lab = O.tumor_pred(full, 0.5, inclusive) * skull
    got = labels.cpu().numpy()
    assert got.dtype == np.uint8 and set(np.unique(got)) <= {0, 1, 2, 4}
    np.testing.assert_array_equal(got, ref_lab)
    # and the predictions themselves are the reference network's (fp tolerance)
    brain = vol[:, bw[0, 0]:bw[1, 0] + 1, bw[0, 1]:bw[1, 1] + 1, bw[0, 2]:bw[1, 2] + 1]
    sd = O.leaf_state(model.state_dict())
    x0 = torch.as_tensor(O.get_patch(brain, (32, 32, 32), corners[0]))[None]
    with torch.no_grad():
        ref0 = O.searched_net(sd, x0, 4, 3, O.G0)[0]
    assert O.max_rel(preds[0].permute(3, 0, 1, 2), ref0) <= 1e-3


def test_seg_to_masks_bit_exact():
    from nas_3d_unet_b200.infer import seg_to_masks
    rng = np.random.default_rng(3)
    t = rng.choice(np.array([0, 1, 2, 4, 3], dtype=np.int16), size=(2, 1, 9, 10, 11))
    for inc in (True, False):
        got = seg_to_masks(torch.as_tensor(t).cuda(), inc).cpu().numpy()
        np.testing.assert_array_equal(got, O.multi_class_labels(t, inc).astype(np.float32))


def test_all_zero_patches_count_as_zero_predictions():
    """prediction.py:133-136: a patch whose data is all zero is not run through the net, its
    prediction is zeros (and it still counts in the overlap mean).  Here the flag is computed on
    the device and consumed by the stitch kernel - same stitched volume, bit for bit."""
    from nas_3d_unet_b200.infer import SlidingWindowPredictor
    model = make_searched().cuda()
    vol = _volume(4, (4, 70, 40, 36))
    vol[:, 30:] = 0                       # the patches of the second half are empty
    pred = SlidingWindowPredictor(model, patch_shape=(32, 32, 32), batch=3)
    labels, stitched, preds, corners = pred.predict(torch.as_tensor(vol).cuda(), return_stitched=True)
    empty = [b for b, c in enumerate(corners) if np.all(O.get_patch(vol, (32, 32, 32), c) == 0)]
    assert empty, "the test volume must contain an all-zero patch"
    for b in range(len(corners)):
        assert bool((preds[b] == 0).all()) == (b in empty)
    pl = [preds[b].permute(3, 0, 1, 2).cpu().numpy() for b in range(preds.shape[0])]
    ref_st = O.stitch(pl, corners.copy(), (3,) + vol.shape[1:])
    np.testing.assert_array_equal(stitched.cpu().numpy(), ref_st)
    np.testing.assert_array_equal(labels.cpu().numpy(), O.tumor_pred(ref_st, 0.5, True))


def test_brats_shape_volume_at_128_matches_oracle():
    """BASELINE.json config #5 at full size: 4x240x240x155 volume, 128^3 auto-fit patching = 9
    patches (corners = SURVEY App. D known answer), batch 4 + 4 + 1 forwards; float64 stitch and
    the label volume bit-equal to the numpy oracle on the same predictions; patch 0 against the
    CPU oracle network (logits tolerance 1e-3)."""
    from nas_3d_unet_b200.infer import SlidingWindowPredictor
    model = make_searched().cuda()
    shape = (4, 240, 240, 155)
    rng = np.random.default_rng(7)
    vol = (rng.random(shape, dtype=np.float32) * 100 + 10).astype(np.float32)
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape[1:]], indexing="ij")
    brain = (((zz - 120) / 85) ** 2 + ((yy - 120) / 100) ** 2 + ((xx - 77) / 65) ** 2) < 1.0   # App. H
    vol *= brain[None]
    skull = brain.astype(np.uint8)
    pred = SlidingWindowPredictor(model, patch_shape=(128, 128, 128), batch=4)
    labels, stitched, preds, corners = pred.predict(torch.as_tensor(vol).cuda(), skull_mask=skull,
                                                    return_stitched=True)
    assert len(corners) == 9
    np.testing.assert_array_equal(corners, O.patching(shape[1:], (128, 128, 128)))
    assert corners[3].tolist() == [0, 112, 27] and corners[8].tolist() == [56, 56, 13]
    pl = [preds[b].permute(3, 0, 1, 2).cpu().numpy() for b in range(9)]
    ref_st = O.stitch(pl, corners.copy(), (3,) + shape[1:])
    np.testing.assert_array_equal(stitched.cpu().numpy(), ref_st)
    ref_lab = O.tumor_pred(ref_st, 0.5, True) * skull
    got = labels.cpu().numpy()
    assert set(np.unique(got)) <= {0, 1, 2, 4}
    np.testing.assert_array_equal(got, ref_lab)
    sd = O.leaf_state(model.state_dict())
    x0 = torch.as_tensor(O.get_patch(vol, (128, 128, 128), corners[0]))[None]
    with torch.no_grad():
        ref0 = O.searched_net(sd, x0, 4, 3, O.G0)[0]
    assert O.max_rel(preds[0].permute(3, 0, 1, 2), ref0) <= 1e-3


def test_deterministic_predictor_is_bit_reproducible():
    """SlidingWindowPredictor(deterministic=True): no split-K atomics in the tcgen05 convs - two runs
    give identical float64 stitched probabilities and labels (the reference's CPU path is
    bit-reproducible; the default predictor trades that for ~10 % speed)"""
    from nas_3d_unet_b200.infer import SlidingWindowPredictor
    model = make_searched().cuda()
    vol = torch.as_tensor(_volume(6, (4, 70, 66, 40))).cuda()
    pred = SlidingWindowPredictor(model, patch_shape=(32, 32, 32), batch=5, deterministic=True)
    l1, s1, _, _ = pred.predict(vol, return_stitched=True)
    l2, s2, _, _ = pred.predict(vol, return_stitched=True)
    assert torch.equal(s1, s2) and torch.equal(l1, l2)
