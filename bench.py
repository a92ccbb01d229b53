#!/usr/bin/env python
"""bench.py - searched-net 128^3 training throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--patch 128] [--batch 8] [--workload searched|supernet]

One "step" = train.py:121-128 on one batch: zero_grad, forward, Dice loss, backward, Adam step
(for --workload supernet: search.py:222-238, one alpha step + one weight step).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "searched-net 128^3 train patches/s"
UNIT = "patches/s"
H2D_DELAY_MS_DEFAULT = -1.0     # < 0: GraphedStep.stream() paces the prefetch from its own measurements


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8, help="patches per GPU per step")
    ap.add_argument("--workload", default="searched", choices=["searched", "supernet"])
    ap.add_argument("--optimizer", choices=["flat", "torch"], default="flat",
                    help="flat = nas_3d_unet_b200.optim.FlatAdam (one launch); torch = torch fused Adam")
    ap.add_argument("--ref-device", choices=["cpu", "cuda"], default="cpu",
                    help="--impl reference only: cuda = the same torch ops through torch's CUDA "
                         "kernels (informational second baseline, BASELINE.md section 4)")
    ap.add_argument("--ref-batch", type=int, default=0,
                    help="--impl reference: patches per timed CPU step; 0 = the workload's own batch if "
                         "(steps + warmup) of them fit --ref-budget seconds, else the largest power of two that "
                         "does (bounded sample of the workload; the line says which)")
    ap.add_argument("--ref-budget", type=float, default=240.0,
                    help="--impl reference: wall-clock budget in seconds used to size the CPU step's batch")
    ap.add_argument("--no-batch-probe", action="store_true",
                    help="--impl reference: skip the single extra step at the full batch size")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the step as one CUDA graph (nas_3d_unet_b200.graph.GraphedStep)")
    ap.add_argument("--labels", default="seg8", choices=["seg8", "masks"],
                    help="what crosses PCIe next to x: seg8 = the int8 segmentation (1 byte per voxel), region "
                         "masks assembled and x re-laid out on the device by data.stage_batch (SURVEY 8 f2); "
                         "masks = three int8 {0,1} masks per voxel as generator.py:230-248 yields them")
    ap.add_argument("--graph-buffers", type=int, default=2, choices=[1, 2],
                    help="static input sets / captured graphs (2: H2D straight into the idle set, GraphedStep.stream)")
    ap.add_argument("--drivers-loop", action="store_true",
                    help="additionally time the loop body of the UNCHANGED reference drivers (train.py:113-128 / "
                         "search.py:211-238: numpy batch -> torch.as_tensor(device, float) every step, eager "
                         "modules, torch.optim.Adam, fp32 labels, loss.item() in the step) and add it to the "
                         "line as `drivers_loop`")
    ap.add_argument("--e2e-probe", action="store_true", help="print an e2e overhead breakdown to stderr")
    ap.add_argument("--h2d-delay-ms", type=float, default=H2D_DELAY_MS_DEFAULT,
                    help="e2e: issue the prefetch of batch i+1 this long after step i started (GraphedStep.stream); < 0 = measured")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel table (json) here")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic data: background 0 / brain U(10,110); labels nested blobs
# ------------------------------------------------------------------------------------------
def synthetic_host_batch(batch, patch, seed, with_seg=False):
    """x (B,4,P,P,P) fp32; y (B,3,P,P,P) fp32 region masks = get_multi_class_labels(seg,
    inclusive_label=True) (generator.py:230-248, with its logical_or quirk) of a segmentation of nested
    blobs 2 > 1 > 4; with_seg also returns that segmentation (B,1,P,P,P) int8"""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 4, patch, patch, patch, generator=g) * 100 + 10
    zz, yy, xx = torch.meshgrid(*[torch.arange(patch, dtype=torch.float32)] * 3, indexing="ij")
    c = (patch - 1) / 2
    r2 = ((zz - c) ** 2 + (yy - c) ** 2 + (xx - c) ** 2) / (patch / 2) ** 2
    x = x * (r2 < 0.9).float()
    seg = torch.zeros_like(r2, dtype=torch.int8)
    seg[r2 < 0.15] = 2
    seg[r2 < 0.05] = 1
    seg[r2 < 0.02] = 4
    y = torch.stack([(seg == 1) | (seg == 4), (seg == 1) | (seg == 2), seg == 4]).float()
    y = y.unsqueeze(0).repeat(batch, 1, 1, 1, 1).contiguous()
    if with_seg:
        return x.contiguous(), y, seg.reshape(1, 1, patch, patch, patch).repeat(batch, 1, 1, 1, 1).contiguous()
    return x.contiguous(), y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------
def _reference_modules():
    """the UNMODIFIED reference's hot-path modules (searched / nas / loss / genotype), imported from
    /root/reference where it exists or from the bytecode oracle/build_ref.py compiled from it
    (oracle/_ref travels to the GPU box); None when neither is available"""
    from oracle import build_ref
    d = build_ref.reference_dir()
    if d is None:
        return None
    import importlib
    sys.path.insert(0, d)
    try:
        mods = {n: importlib.import_module(n) for n in ("prim_ops", "cell", "genotype", "nas", "searched", "loss")}
    finally:
        sys.path.remove(d)
    if not all(os.path.dirname(os.path.abspath(m.__file__)) == os.path.abspath(d) for m in mods.values()):
        return None           # something else named nas / searched / loss shadowed them
    return mods


def cpu_reference_steps(workload, patch, steps, warmup, batch=1, device="cpu", probe_batch=0,
                        full_batch=None, budget_s=240.0):
    """times `steps` training steps of the reference path on the host cores; returns (patches/s,
    s/step, cores, kind, sample description, extra).  kind "reference": the reference's own modules
    (train.py:121-128 / search.py:222-238 loop bodies around them); kind "port": the oracle port
    when the reference modules are unavailable.  device="cuda" runs the oracle's torch-op graph
    through torch's own CUDA kernels (cuDNN) - the reference's intended GPU path (search.py:66-69);
    informational only, never the baseline.  probe_batch > 0 additionally times ONE step at that
    batch size, to show how CPU patches/s depends on the batch."""
    from oracle import nas3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    mods = _reference_modules() if device == "cpu" else None
    p_drop = 0.5 if workload == "searched" else 0.1
    if mods is not None:
        kind = "reference"
        if workload == "searched":
            gene = mods["genotype"].Genotype(down=O.G0.down, up=O.G0.up)
            model = mods["searched"].SearchedNet(4, 4, 3, 4, 3, True, gene)
            optims = [torch.optim.Adam(model.parameters())]
        else:
            model = mods["nas"].ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True)
            optims = [torch.optim.Adam(model.alphas()), torch.optim.Adam(model.kernel.parameters())]
        model.train()
        lossf = mods["loss"].WeightedDiceLoss()

        def make_step(b):
            x, y = synthetic_host_batch(b, patch, seed=1234)

            def one():
                last = 0.0
                for opt in optims:          # search: alpha step then weight step (search.py:222-238)
                    opt.zero_grad()
                    loss = lossf(model(x), y)
                    last = loss.item()
                    loss.backward()
                    opt.step()
                return last
            return one
    else:
        kind = "port"
        if workload == "searched":
            from nas_3d_unet_b200.searched import SearchedNet
            from nas_3d_unet_b200.genotype import Genotype
            m = SearchedNet(4, 4, 3, 4, 3, True, Genotype(down=O.G0.down, up=O.G0.up))
        else:
            from nas_3d_unet_b200.nas import ShellNet
            m = ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True)
        sd = O.leaf_state(m.state_dict())       # parameters only; the modules themselves never run
        if device != "cpu":
            sd = {k: (v.detach().to(device).requires_grad_(True) if v.is_floating_point() else v.to(device))
                  for k, v in sd.items()}
        opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3)
        passes = 2 if workload == "supernet" else 1

        def make_step(b):
            x, y = synthetic_host_batch(b, patch, seed=1234)
            x, y = x.to(device), y.to(device)

            def one():
                last = 0.0
                for _ in range(passes):
                    opt.zero_grad()
                    mask = torch.empty((b, 12, 1, 1, 1), device=device).bernoulli_(1 - p_drop).div_(1 - p_drop)
                    pred = (O.searched_net(sd, x, 4, 3, O.G0, drop_mask=mask) if workload == "searched"
                            else O.shell_net(sd, x, 4, 3, drop_mask=mask))
                    loss = O.dice_loss(pred, y)
                    loss.backward()
                    opt.step()
                    last = loss.item()
                return last
            return one

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()

    if batch <= 0:
        # size the sample: one batch-1 step (it doubles as a warm-up), then the largest batch whose
        # (steps + warmup) steps fit the budget.  Measured on the 16-thread GPU box: a batch-8 step costs
        # 4.6x a batch-1 step (0.63 -> 1.09 patches/s), hence the 0.6 per extra patch.
        t0 = time.perf_counter()
        make_step(1)()
        sync()
        t1 = time.perf_counter() - t0
        batch = 1
        b = full_batch or 1
        while b > 1:
            if (steps + warmup) * t1 * (1 + 0.6 * (b - 1)) <= budget_s:
                batch = b
                break
            b //= 2
    one = make_step(batch)
    for _ in range(warmup):
        one()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    sync()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    extra = {}
    if probe_batch and probe_batch != batch:
        big = make_step(probe_batch)
        t0 = time.perf_counter()
        big()
        sync()
        dtb = time.perf_counter() - t0
        extra["batch_scaling"] = {"batch_%d_patches_per_s" % batch: batch / dt,
                                  "batch_%d_patches_per_s" % probe_batch: probe_batch / dtb,
                                  "note": "one untimed-warm-up-free step at batch %d" % probe_batch}
    what = {"reference": "the reference's own modules (unmodified; %s)" % (
                "imported from /root/reference" if os.path.isdir("/root/reference")
                else "bytecode compiled from /root/reference by oracle/build_ref.py"),
            "port": "oracle port"}[kind]
    sample = ("%s net, %d step(s) of batch %d at %d^3 (fwd + Dice + bwd + Adam), %s on %s"
              % (workload, steps, batch, patch, what,
                 "%d host threads" % torch.get_num_threads() if device == "cpu" else "torch eager CUDA (cuDNN)"))
    extra["batch"] = batch
    return batch / dt, dt, torch.get_num_threads(), kind, sample, extra


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the workload's own batch per timed step when (steps + warmup) of them fit the time budget
    # (--ref-budget, default 4 minutes), else a bounded sample at a smaller batch - CPU patches/s is NOT
    # flat in the batch (0.63 at batch 1, 1.09 at batch 8 on the 16-thread box), so the line states the
    # batch it timed and, when that is not the full one, a single probe step at the full batch
    on_gpu = args.ref_device == "cuda"
    v, dt, cores, kind, sample, extra = cpu_reference_steps(
        args.workload, args.patch, args.steps, args.warmup, batch=args.batch if on_gpu else args.ref_batch,
        device=args.ref_device, probe_batch=0 if (on_gpu or args.no_batch_probe) else args.batch,
        full_batch=args.batch, budget_s=args.ref_budget)
    b = extra.pop("batch")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                              "batch_per_timed_step": b, "device": args.ref_device}, **extra),
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    """BASELINE.json's metric for the default workload; the secondary workloads say what they are"""
    if args.workload == "searched" and args.patch == 128:
        return METRIC
    if args.workload == "searched":
        return "searched-net %d^3 train patches/s" % args.patch
    return "supernet %d^3 search patches/s (train+val pair = 1 patch)" % args.patch


def workload_config(args):
    """the workload both arms are quoted on (identical dict in both lines); the reference arm times a
    bounded sample of it and says which in cpu_baseline.sample / batch_per_timed_step"""
    cfg = {
        "workload": ("%s-G0 U-Net training step (fwd + Dice + bwd + Adam), 4x%d^3 patches, "
                     "batch %d per GPU" % (args.workload, args.patch, args.batch))
        if args.workload == "searched" else
        ("supernet search step (alpha step + weight step), 4x%d^3 patches, batch %d per GPU"
         % (args.patch, args.batch)),
        "patch": args.patch, "batch_per_gpu": args.batch,
        "global_batch": args.batch * args.gpus,
        "parallelism": "dp%d" % args.gpus,
        "l2": "inputs larger than L2 (%.0f MB of x (fp32) + %s per step per GPU vs 126 MB L2)"
              % (((4 * 4 + (1 if args.labels == "seg8" else 3)) * args.patch ** 3 * args.batch) / 1e6,
                 "int8 segmentation (masks assembled on the device)" if args.labels == "seg8" else "y (int8 masks)"),
    }
    return cfg


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from nas_3d_unet_b200 import _lib, engine, profiling
    from nas_3d_unet_b200.loss import WeightedDiceLoss

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        engine.enable_data_parallel()
    _lib.load()

    def make_adam(params):
        if args.optimizer == "flat":   # our one-launch flat-arena Adam
            from nas_3d_unet_b200.optim import FlatAdam
            return FlatAdam(params, lr=1e-3)
        return torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)

    torch.manual_seed(0)
    lossf = WeightedDiceLoss().to(dev)
    if args.workload == "searched":
        from nas_3d_unet_b200.searched import SearchedNet
        from nas_3d_unet_b200.genotype import G0
        model = SearchedNet(4, 4, 3, 4, 3, True, G0).to(dev)
        opts = [make_adam(model.parameters())]
    else:
        from nas_3d_unet_b200.nas import ShellNet
        model = ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True).to(dev)
        opts = [make_adam(model.alphas()), make_adam(model.kernel.parameters())]
    model.train()

    B, P = args.batch, args.patch
    staged = args.labels == "seg8"
    hx, hmask, hseg = synthetic_host_batch(B, P, seed=1234 + rank, with_seg=True)
    # labels travel as the int8 segmentation (the device assembles the masks) or as int8 {0,1} masks,
    # which is what generator.py:230-248 hands the drivers
    hx, hy = hx.pin_memory(), (hseg if staged else hmask.to(torch.int8)).pin_memory()
    if args.workload == "supernet":
        hvx, hvmask, hvseg = synthetic_host_batch(B, P, seed=4321 + rank, with_seg=True)
        hvx, hvy = hvx.pin_memory(), (hvseg if staged else hvmask.to(torch.int8)).pin_memory()

    def prep(x, y):
        """device-side batch assembly (data.stage_batch: NCDHW -> channels-last, int8 seg -> int8 masks)"""
        if not staged:
            return x, y
        from nas_3d_unet_b200.data import stage_batch
        return stage_batch(x, y, None, True)
    dx, dy = hx.to(dev), hy.to(dev)
    if args.workload == "supernet":
        dvx, dvy = hvx.to(dev), hvy.to(dev)
    patches_per_step = B * world

    def fwd_loss(x, y):
        xs, ys = prep(x, y)
        return lossf(model(xs), ys)

    def step_resident():
        if args.workload == "searched":
            opts[0].zero_grad()
            loss = fwd_loss(dx, dy)
            loss.backward()
            opts[0].step()
        else:
            opts[0].zero_grad()
            vl = fwd_loss(dvx, dvy)
            vl.backward()
            opts[0].step()
            opts[1].zero_grad()
            loss = fwd_loss(dx, dy)
            loss.backward()
            opts[1].step()
        return loss

    def step_fn(*batch):
        """one complete step on device tensors; returns the (last) loss tensor"""
        if args.workload == "searched":
            x, y = prep(*batch)
            opts[0].zero_grad(set_to_none=True)
            loss = lossf(model(x), y)
            loss.backward()
            opts[0].step()
            return loss
        x, y = prep(*batch[:2])
        vx, vy = prep(*batch[2:])
        opts[0].zero_grad(set_to_none=True)
        vl = lossf(model(vx), vy)
        vl.backward()
        opts[0].step()
        opts[1].zero_grad(set_to_none=True)
        loss = lossf(model(x), y)
        loss.backward()
        opts[1].step()
        return torch.stack([vl, loss])

    graphed = None
    if args.graph == "on":
        from nas_3d_unet_b200.graph import GraphedStep
        ex = (dx, dy) if args.workload == "searched" else (dx, dy, dvx, dvy)
        graphed = GraphedStep(step_fn, ex, warmup=max(args.warmup, 3), optimizers=opts,
                              # two captured supernet-128^3 steps (2 x ~12 000 kernel nodes) crash the driver in
                              # cudaGraphLaunch (B200, driver 580): the supernet keeps one graph; its batch of 1
                              # makes the H2D of a step 33 MB, so the double buffer would buy nothing there
                              buffers=args.graph_buffers if args.workload == "searched" else 1)

    def host_batches(n):
        """what a data pipeline hands the step loop: pinned host tensors"""
        for _ in range(n):
            if args.workload == "searched":
                yield hx, hy
            else:
                yield hx, hy, hvx, hvy

    def run_e2e(nsteps):
        """the loop a user of train.py / search.py runs: host batches in (H2D every step, overlapped
        with the previous step by nas_3d_unet_b200.data.DevicePrefetcher), loss value out (D2H)"""
        from nas_3d_unet_b200.data import DevicePrefetcher
        last = 0.0
        if graphed is not None:
            # GraphedStep.stream: batch i+1 goes H2D straight into the idle static input set while
            # graph i runs; every step's loss comes back to the host (4 bytes), read one step late
            for last in graphed.stream(host_batches(nsteps), h2d_delay_s=None if args.h2d_delay_ms < 0 else args.h2d_delay_ms * 1e-3):
                pass
            return last
        for batch in DevicePrefetcher(host_batches(nsteps), dev):
            if args.workload == "searched":
                x, y = prep(*batch)
                opts[0].zero_grad()
                loss = lossf(model(x), y)
                last = loss.item()
                loss.backward()
                opts[0].step()
            else:
                x, y = prep(*batch[:2])
                vx, vy = prep(*batch[2:])
                opts[0].zero_grad()
                vl = lossf(model(vx), vy)
                last = vl.item()
                vl.backward()
                opts[0].step()
                opts[1].zero_grad()
                loss = lossf(model(x), y)
                last = loss.item()
                loss.backward()
                opts[1].step()
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # kernels per step (counted on an eager step; a graph replay launches the same kernels)
    n0 = _lib.launch_count()
    step_resident()
    launches = (_lib.launch_count() - n0) * args.steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if graphed is not None:
        graphed.load(*((dx, dy) if args.workload == "searched" else (dx, dy, dvx, dvy)))
        for _ in range(3):
            graphed.replay()
        ms = timed(graphed.replay, args.steps)
    else:
        ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = patches_per_step / (ms_per_step * 1e-3)

    if args.e2e_probe and graphed is not None and rank == 0:
        # where the gap between `value` and `e2e` goes (stderr only): alternating replays of the two
        # graphs without any copy, the stream() loop with device-resident "batches" (D2D instead of
        # PCIe: same host-side work, no DMA from the host), and the real thing
        nset = len(graphed.sets)

        def alt():
            for i in range(args.steps):
                graphed.sets[i % nset][1].replay()
        dev_batches = [tuple(t.clone() for t in graphed.static_in)]

        def stream_dev():
            for _ in graphed.stream(dev_batches * args.steps):
                pass
        side = torch.cuda.Stream(device=dev)
        scratch = torch.empty_like(hx, device=dev)

        def alt_with_copy(frac, from_host):
            n = int(hx.shape[0] * frac) or 1
            src = hx[:n] if from_host else dx[:n]

            def fn():
                for i in range(args.steps):
                    with torch.cuda.stream(side):
                        scratch[:n].copy_(src, non_blocking=True)
                    graphed.sets[i % nset][1].replay()
                torch.cuda.current_stream().wait_stream(side)
            return fn
        for name, fn in (("alternating replays, no copies", alt), ("stream(), device-resident batches", stream_dev),
                         ("replays + unrelated H2D of x (268 MB)", alt_with_copy(1.0, True)),
                         ("replays + unrelated H2D of x/4 (67 MB)", alt_with_copy(0.25, True)),
                         ("replays + unrelated D2D of x (268 MB)", alt_with_copy(1.0, False)),
                         ("stream(), pinned host batches", lambda: run_e2e(args.steps))):
            fn()
            print("e2e-probe %-36s %.3f ms/step" % (name, timed(fn, 1) / args.steps), file=sys.stderr)

    # end to end: pinned host batch -> device, loss value -> host, every step
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1) / args.steps
    n_in = 2 if args.workload == "supernet" else 1
    h2d = n_in * (hx.numel() * hx.element_size() + hy.numel() * hy.element_size())
    d2h = 4 * n_in
    e2e = {"value": patches_per_step / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    drivers_loop = None
    if args.drivers_loop:
        # what a user gets WITHOUT touching the reference drivers: their own step loop, verbatim, on
        # fresh replicas of the modules (eager, torch Adam, pageable numpy batches, fp32 labels)
        import numpy as np
        torch.manual_seed(0)
        if args.workload == "searched":
            from nas_3d_unet_b200.searched import SearchedNet
            from nas_3d_unet_b200.genotype import G0
            dmodel = SearchedNet(4, 4, 3, 4, 3, True, G0).to(dev)
            dopts = [torch.optim.Adam(dmodel.parameters())]
        else:
            from nas_3d_unet_b200.nas import ShellNet
            dmodel = ShellNet(4, 4, 3, 4, 3, normal_w_share=False, channel_change=True).to(dev)
            dopts = [torch.optim.Adam(dmodel.alphas()), torch.optim.Adam(dmodel.kernel.parameters())]
        dmodel.train()
        nx, ny = hx.numpy().copy(), hmask.numpy().astype(np.int8)     # what generator.convert_data yields

        def driver_step():
            last = 0.0
            for opt in dopts:
                x = torch.as_tensor(nx, device=dev, dtype=torch.float)
                y_truth = torch.as_tensor(ny, device=dev, dtype=torch.float)
                opt.zero_grad()
                y_pred = dmodel(x)
                loss = lossf(y_pred, y_truth)
                last += loss.item()
                loss.backward()
                opt.step()
            return last
        for _ in range(3):
            driver_step()
        nst = max(3, args.steps // 2)
        ms_d = timed(driver_step, nst) / nst
        drivers_loop = {"value": patches_per_step / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d,
                        "what": "reference driver loop body unchanged: eager modules, torch.optim.Adam, "
                                "torch.as_tensor of numpy batches (fp32 x, int8 -> fp32 labels), loss.item() per step"}
        del dmodel, dopts

    roofline = None
    if not args.no_roofline and rank == 0:
        # per-kernel timing of rank 0's own replica: the gradient all-reduce must be off here, the
        # other ranks are not stepping (they wait at the barrier below)
        engine.enable_data_parallel(enabled=False)
        roofline = roofline_pass(step_resident, profiling, args, ms_per_step)
        engine.enable_data_parallel(enabled=world > 1)
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample for the default run: about 20 s of CPU work at whatever batch fits
        v, dt, cores, kind, sample, ex = cpu_reference_steps(args.workload, P, steps=2, warmup=0, batch=0,
                                                             full_batch=B, budget_s=20.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
               "batch_per_timed_step": ex["batch"], "s_per_step": dt}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), cuda_graph=(graphed is not None)),
            "voxels_per_s": value * P ** 3,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "drivers_loop": drivers_loop,
            # the per-family table again at top level (survives parsers that flatten `roofline`)
            "extra": {"families": roofline["by_kernel"], "summed_roofline": roofline["summed"],
                      "summed_roofline_vs_graph_step": roofline["summed_vs_graph_step"]} if roofline else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear down in dependency order: the captured graph holds NCCL kernels, and destroying the
        # communicator under a live graph can block forever - drop the graph, drain, then leave
        # without the collective teardown (every rank has already passed the last barrier)
        graphed = None
        import gc
        gc.collect()
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def ncu_traffic(kernel, shape):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this (C-ABI entry, shape) from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json: a profiler figure taken once per
    round, NOT re-measured by this run), or None if not captured"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["%s|%s" % (kernel, shape)]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def measure_fma_roof(dev):
    """fp32 FMA roof of THIS run: the library's probe kernel (packed-FMA chains, no memory traffic)
    timed with CUDA events, best of 5"""
    from nas_3d_unet_b200 import _lib
    lib = _lib.load()
    n = 148 * 8 * 256
    out = torch.empty(n, device=dev, dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.nas3d_probe_fma(out.data_ptr(), n, 2048, st)
        e1.record()
        torch.cuda.synchronize()
        if flops <= 0:
            return None
        tf = flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
        best = tf if best is None or tf > best else best
    return best


def roofline_pass(step, profiling, args, ms_per_step):
    """2 extra eager steps with every C-ABI launch bracketed by CUDA events on the launch stream.
    Reported: the DOMINANT C-ABI FAMILY (largest share of the step's kernel time) against its
    binding roof, the per-family table, and the summed roofline of SURVEY 8d:
        summed = sum_k max(F_k / compute_roof_k, B_k / HBM) / sum_k t_k
    with F_k / B_k the algorithmic flops / bytes of launch k (profiling.py) and compute_roof_k the
    pipe the kernel runs on: fp32 FMA for the CUDA-core convolutions, TF32 tensor (= measured bf16
    / 2) for the tcgen05 ones; everything else is HBM-only."""
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            j = json.load(f)
        peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                 "src": "MEASURED_PEAKS.json (sustained bf16)"}
    fma_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    fma_measured = measure_fma_roof(torch.device("cuda", torch.cuda.current_device()))
    fma_roof = fma_measured or fma_nominal
    tf32_roof = peaks["bf16_tflops"] / 2.0
    hbm = peaks["hbm_gbs"]
    prof = profiling.enable()
    try:
        nsteps = 2
        for _ in range(nsteps):
            step()
        rows = prof.summary()
    finally:
        profiling.disable()

    def roofs(r):
        """(roof seconds, bound) of one (kernel, shape) row"""
        t_mem = r["bytes"] / (hbm * 1e9)
        if r["flops"] <= 0:
            return t_mem, "hbm"
        if r["kernel"] == "nas3d_umma_conv":
            t_c, b = r["flops"] / (tf32_roof * 1e12), "tensor"
        else:
            t_c, b = r["flops"] / (fma_roof * 1e12), "fp32_fma"
        return (t_c, b) if t_c >= t_mem else (t_mem, "hbm")

    total_ms = sum(r["ms"] for r in rows)
    fam = {}
    for r in rows:
        t_roof, bound = roofs(r)
        a = fam.setdefault(r["kernel"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0,
                                         "roof_ms": 0.0, "bound_ms": {}})
        a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["flops"] += r["flops"]; a["launches"] += r["launches"]
        a["roof_ms"] += t_roof * 1e3
        a["bound_ms"][bound] = a["bound_ms"].get(bound, 0.0) + t_roof * 1e3
    roof_ms_total = sum(a["roof_ms"] for a in fam.values())
    order = sorted(fam.items(), key=lambda kv: -kv[1]["ms"])
    name, top = order[0]
    bound = max(top["bound_ms"].items(), key=lambda kv: kv[1])[0]
    ach_gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    ach_tf = top["flops"] / (top["ms"] * 1e-3) / 1e12
    if bound == "hbm":
        achieved, peak, unit, src = ach_gbs, hbm, "GB/s", peaks["src"]
    elif bound == "tensor":
        achieved, peak, unit, src = ach_tf, tf32_roof, "TFLOP/s", peaks["src"] + " / 2 for TF32"
    else:
        achieved, peak, unit = ach_tf, fma_roof, "TFLOP/s"
        src = ("fp32 FMA roof measured in this run by nas3d_probe_fma (packed FFMA2 chains)"
               if fma_measured else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz")
    # the family's heaviest shape: per-launch figures (what one ncu capture of it shows)
    trow = max((r for r in rows if r["kernel"] == name), key=lambda r: r["ms"])
    t_roof, t_bound = roofs(trow)
    top_shape = {"shape": trow["shape"], "launches_per_step": trow["launches"] // nsteps,
                 "avg_launch_ms": trow["ms"] / trow["launches"], "bound": t_bound,
                 "algorithmic_bytes_per_launch": trow["bytes"] / trow["launches"],
                 "algorithmic_flops_per_launch": trow["flops"] / trow["launches"],
                 "achieved_tflops": trow["flops"] / (trow["ms"] * 1e-3) / 1e12,
                 "achieved_gbps": trow["bytes"] / (trow["ms"] * 1e-3) / 1e9,
                 "frac_of_binding_roof": t_roof * 1e3 / trow["ms"],
                 "traffic": ncu_traffic(name, trow["shape"])}
    table = {}
    for k, v in order:
        b = max(v["bound_ms"].items(), key=lambda kv: kv[1])[0]
        table[k] = {"ms_per_step": v["ms"] / nsteps, "launches_per_step": v["launches"] // nsteps,
                    "share": v["ms"] / total_ms if total_ms else None,
                    "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0,
                    "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0,
                    "bound": b, "frac_of_binding_roof": (v["roof_ms"] / v["ms"]) if v["ms"] else None}
    out = {
        "bound": bound, "kernel": name, "scope": "all launches of this C-ABI entry point in the step",
        "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
        "frac_of_binding_roof_per_launch_sum": top["roof_ms"] / top["ms"],
        "peak_source": src, "traffic": top_shape["traffic"], "top_shape": top_shape,
        "share_of_step": top["ms"] / total_ms if total_ms else None,
        "launches_per_step": top["launches"] // nsteps,
        "avg_launch_ms": top["ms"] / top["launches"],
        "achieved_tflops": ach_tf, "achieved_gbps": ach_gbs,
        "fp32_fma_roof_tflops": {"measured_this_run": fma_measured, "nominal": fma_nominal},
        "tf32_tensor_roof_tflops": tf32_roof, "hbm_roof_gbps": hbm,
        "frac_of_hbm": ach_gbs / hbm, "frac_of_tf32_tensor": ach_tf / tf32_roof,
        # whole step against the sum of its kernels' binding roofs
        "summed": roof_ms_total / total_ms if total_ms else None,
        "summed_vs_graph_step": (roof_ms_total / nsteps) / ms_per_step if ms_per_step else None,
        "summed_roof_ms_per_step": roof_ms_total / nsteps,
        "step_hbm_frac": (sum(r["bytes"] for r in rows) / (total_ms * 1e-3) / 1e9) / hbm,
        "kernel_ms_per_step": total_ms / nsteps,
        "by_kernel": table,
    }
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        for r in rows:
            t_roof, b = roofs(r)
            r["roof_ms"], r["bound"] = t_roof * 1e3, b
        with open(args.profile_out, "w") as f:
            json.dump({"steps": nsteps, "rows": rows[:80]}, f, indent=1)
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
