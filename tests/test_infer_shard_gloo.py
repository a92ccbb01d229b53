"""CPU, world_size 2 and 3, gloo: the host-side logic of sharded sliding-window inference
(nas_3d_unet_b200/infer.py; reference loop prediction.py:121-148, stitch patches.py:172-206).

The patches of one volume are dealt round-robin to the ranks, each rank 'predicts' its own, one
all_gather hands every rank every patch, and the order-preserving float64 stitch must then be
bit-equal to the single-process result on EVERY rank.  The per-patch prediction is a stand-in
function of (patch index, voxel) - what is under test is ownership, padding of the uneven last
round (9 patches over 2 ranks = 5 + 4) and the re-interleave after the gather; the CUDA kernels
themselves are covered by tests/test_gpu_infer.py and the 2-GPU run of tools/bench_inference.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_patch_prediction(b, P):
    g = torch.Generator().manual_seed(1000 + b)
    return torch.rand((P, P, P, 4), generator=g)


def _worker(rank, world, port, q, img_shape, P):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from nas_3d_unet_b200 import infer
        from oracle import nas3d_oracle as O
        corners = infer.patching(img_shape, (P, P, P))
        B = len(corners)
        mine = infer.shard_indices(B, rank, world)
        local = (torch.stack([_fake_patch_prediction(b, P) for b in mine]) if mine
                 else torch.zeros((0, P, P, P, 4)))
        per_patch = infer.gather_patch_predictions(local, B)
        assert len(per_patch) == B
        got = [t[..., :3].permute(3, 0, 1, 2).numpy() for t in per_patch]
        st = O.stitch(got, corners.copy(), (3,) + tuple(img_shape))
        # single-process truth
        ref = [_fake_patch_prediction(b, P)[..., :3].permute(3, 0, 1, 2).numpy() for b in range(B)]
        st_ref = O.stitch(ref, corners.copy(), (3,) + tuple(img_shape))
        ok = all(np.array_equal(a, b) for a, b in zip(got, ref)) and np.array_equal(st, st_ref)
        lab = O.tumor_pred(st, 0.5, True)
        q.put((rank, bool(ok), B, len(mine), int(lab.astype(np.int64).sum())))
    finally:
        dist.destroy_process_group()


def _run(world, img_shape, P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, img_shape, P)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_patch_sharded_stitch_is_rank_invariant_world2():
    # (60,60,39) @ 32^3: the BraTS 240x240x155 @ 128^3 geometry at quarter scale -> 9 patches
    res = _run(2, (60, 60, 39), 32)
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == 9 and sorted(r[3] for r in res) == [4, 5]
    assert res[0][4] == res[1][4]          # both ranks assemble the same label volume


def test_patch_sharded_stitch_is_rank_invariant_world3_with_idle_round():
    # 2 patches over 3 ranks: one rank owns nothing and still takes part in the gather
    res = _run(3, (20, 30, 30), 32)
    assert all(r[1] for r in res)
    assert sorted(r[3] for r in res) == [0, 1, 1]


def test_volume_sharding_plan():
    sys.path.insert(0, ROOT)
    from nas_3d_unet_b200.infer import shard_indices
    for n, world in ((8, 8), (17, 8), (3, 2), (1, 4)):
        owned = [shard_indices(n, r, world) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
