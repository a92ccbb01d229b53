"""CPU, world_size 2, gloo: the host-side data-parallel logic (flat gradient bucket mean
all-reduce issued from the module backward; rank-0-only reference arm of bench.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from nas_3d_unet_b200 import engine

        class Ctx:      # the two fields _dp_allreduce touches
            pass

        class Extra:
            pass
        ctx = Ctx()
        ctx.bucket = torch.arange(16, dtype=torch.float32) * (rank + 1)
        e = Extra()
        e.g = torch.full((9, 5), float(rank + 1))
        e2 = Extra()
        e2.g = None
        engine.enable_data_parallel()
        engine._dp_allreduce(ctx, [e, e2])
        engine.enable_data_parallel(enabled=False)
        expect = torch.arange(16, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok = torch.allclose(ctx.bucket, expect) and torch.allclose(
            e.g, torch.full((9, 5), sum(range(1, world + 1)) / world))
        # disabled => untouched
        b2 = torch.ones(4) * (rank + 1)
        ctx.bucket = b2.clone()
        if engine._dp_state["enabled"]:
            engine._dp_allreduce(ctx, [])
        ok = ok and torch.equal(ctx.bucket, b2)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bucket_mean_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    outs = []
    for rank in (0, 1):
        env["RANK"] = str(rank)
        env["LOCAL_RANK"] = str(rank)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                            "--gpus", "2", "--steps", "1", "--warmup", "0", "--patch", "32"],
                           capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["value"] > 0
    # the reference's own modules where they (or their compiled copy, oracle/_ref) exist, else the port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["batch_per_timed_step"] in (1, 2, 4, 8)      # the batch it timed, stated
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2
