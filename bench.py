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
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the step as one CUDA graph (nas_3d_unet_b200.graph.GraphedStep)")
    ap.add_argument("--e2e-probe", action="store_true", help="print an e2e overhead breakdown to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel table (json) here")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic data: background 0 / brain U(10,110); labels nested blobs
# ------------------------------------------------------------------------------------------
def synthetic_host_batch(batch, patch, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 4, patch, patch, patch, generator=g) * 100 + 10
    zz, yy, xx = torch.meshgrid(*[torch.arange(patch, dtype=torch.float32)] * 3, indexing="ij")
    c = (patch - 1) / 2
    r2 = ((zz - c) ** 2 + (yy - c) ** 2 + (xx - c) ** 2) / (patch / 2) ** 2
    x = x * (r2 < 0.9).float()
    y = torch.stack([(r2 < 0.05), (r2 < 0.15), (r2 < 0.02)]).float().unsqueeze(0).repeat(batch, 1, 1, 1, 1)
    return x.contiguous(), y.contiguous()


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
# reference arm / cpu baseline: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_steps(workload, patch, steps, warmup, batch=1, device="cpu"):
    """times `steps` training steps of the oracle (CPU restatement of the reference) at batch 1;
    returns (patches_per_s, seconds_per_step, cores, sample description).
    device="cuda" runs the same torch-op graph through torch's own CUDA kernels (cuDNN) - the
    reference's intended GPU path (search.py:66-69); informational only, never the baseline."""
    from oracle import nas3d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
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
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    x, y = synthetic_host_batch(batch, patch, seed=1234)
    x, y = x.to(device), y.to(device)
    p_drop = 0.5 if workload == "searched" else 0.1

    def one():
        opt.zero_grad()
        mask = torch.empty((batch, 12, 1, 1, 1), device=device).bernoulli_(1 - p_drop).div_(1 - p_drop)
        if workload == "searched":
            pred = O.searched_net(sd, x, 4, 3, O.G0, drop_mask=mask)
        else:
            pred = O.shell_net(sd, x, 4, 3, drop_mask=mask)
        loss = O.dice_loss(pred, y)
        loss.backward()
        opt.step()
        return loss.item()

    # a search step is two passes (alpha step on the val batch + weight step on the train batch,
    # search.py:222-238); both cost the same here, so the same pass is timed twice
    passes = 2 if workload == "supernet" else 1
    for _ in range(warmup):
        one()
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps * passes):
        one()
    if device != "cpu":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt, torch.get_num_threads(), (
        "%s net, %d step(s) of batch %d at %d^3 (fwd+bwd+Adam), oracle port on %s"
        % (workload, steps, batch, patch, "CPU" if device == "cpu" else "torch eager CUDA (cuDNN)"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: batch 1 per step (--ref-device cuda: informational torch-eager GPU run, batch
    # as given)
    on_gpu = args.ref_device == "cuda"
    v, dt, cores, sample = cpu_reference_steps(args.workload, args.patch, args.steps, args.warmup,
                                               batch=args.batch if on_gpu else 1,
                                               device=args.ref_device)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, reference=True),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "device": args.ref_device},
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


def workload_config(args, reference=False):
    """same workload description for both arms; the reference arm times a bounded sample of it
    (one patch per step - GroupNorm and Dice are per-sample, so CPU patches/s is flat in the batch)"""
    cfg = {
        "workload": ("%s-G0 U-Net training step (fwd + Dice + bwd + Adam), 4x%d^3 patches, "
                     "batch %d per GPU" % (args.workload, args.patch, args.batch))
        if args.workload == "searched" else
        ("supernet search step (alpha step + weight step), 4x%d^3 patches, batch %d per GPU"
         % (args.patch, args.batch)),
        "patch": args.patch, "batch_per_gpu": args.batch,
        "global_batch": args.batch * args.gpus,
        "parallelism": "dp%d" % args.gpus,
        "l2": "inputs larger than L2 (%.0f MB of x (fp32) + y (int8 masks) per step per GPU vs 126 MB L2)"
              % (((4 * 4 + 3) * args.patch ** 3 * args.batch) / 1e6),
    }
    if reference:
        cfg["reference_sample"] = "each timed step = 1 patch of this workload on the host cores"
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
    hx, hy = synthetic_host_batch(B, P, seed=1234 + rank)
    # labels travel as int8 {0,1} masks, which is what generator.py:230-248 hands the drivers
    hx, hy = hx.pin_memory(), hy.to(torch.int8).pin_memory()
    if args.workload == "supernet":
        hvx, hvy = synthetic_host_batch(B, P, seed=4321 + rank)
        hvx, hvy = hvx.pin_memory(), hvy.to(torch.int8).pin_memory()
    dx, dy = hx.to(dev), hy.to(dev)
    if args.workload == "supernet":
        dvx, dvy = hvx.to(dev), hvy.to(dev)
    patches_per_step = B * world

    def step_resident():
        if args.workload == "searched":
            opts[0].zero_grad()
            loss = lossf(model(dx), dy)
            loss.backward()
            opts[0].step()
        else:
            opts[0].zero_grad()
            vl = lossf(model(dvx), dvy)
            vl.backward()
            opts[0].step()
            opts[1].zero_grad()
            loss = lossf(model(dx), dy)
            loss.backward()
            opts[1].step()
        return loss

    def step_fn(*batch):
        """one complete step on device tensors; returns the (last) loss tensor"""
        if args.workload == "searched":
            x, y = batch
            opts[0].zero_grad(set_to_none=True)
            loss = lossf(model(x), y)
            loss.backward()
            opts[0].step()
            return loss
        x, y, vx, vy = batch
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
        graphed = GraphedStep(step_fn, ex, warmup=max(args.warmup, 3), optimizers=opts)

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
        for batch in DevicePrefetcher(host_batches(nsteps), dev):
            if graphed is not None:
                # the prefetched device batch is copied into the graph's static inputs (D2D, cheap);
                # the NEXT batch's PCIe transfer overlaps this replay on the copy stream
                last = graphed(*batch).reshape(-1)[-1].item()
                continue
            if args.workload == "searched":
                x, y = batch
                opts[0].zero_grad()
                loss = lossf(model(x), y)
                last = loss.item()
                loss.backward()
                opts[0].step()
            else:
                x, y, vx, vy = batch
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
        # where the gap between `value` and `e2e` goes (stderr only)
        src = tuple(t.clone() for t in graphed.static_in)

        def b_fn():
            graphed(*src)

        def c_fn():
            graphed(*src).reshape(-1)[-1].item()
        for name, fn in (("replay", graphed.replay), ("d2d+replay", b_fn), ("d2d+replay+item", c_fn)):
            fn()
            print("e2e-probe %-18s %.3f ms/step" % (name, timed(lambda: [fn() for _ in range(args.steps)], 1)
                                                    / args.steps), file=sys.stderr)
        run_e2e(2)
        for k in (args.steps, 4 * args.steps):
            print("e2e-probe full e2e, %3d steps  %.3f ms/step"
                  % (k, timed(lambda: run_e2e(k), 1) / k), file=sys.stderr)

    # end to end: pinned host batch -> device, loss value -> host, every step
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1) / args.steps
    n_in = 2 if args.workload == "supernet" else 1
    h2d = n_in * (hx.numel() * hx.element_size() + hy.numel() * hy.element_size())
    d2h = 4 * n_in
    e2e = {"value": patches_per_step / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    roofline = None
    if not args.no_roofline and rank == 0:
        # per-kernel timing of rank 0's own replica: the gradient all-reduce must be off here, the
        # other ranks are not stepping (they wait at the barrier below)
        engine.enable_data_parallel(enabled=False)
        roofline = roofline_pass(step_resident, profiling, args)
        engine.enable_data_parallel(enabled=world > 1)
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores, sample = cpu_reference_steps(args.workload, P, steps=3, warmup=1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "s_per_step": dt}

    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), cuda_graph=(graphed is not None)),
            "voxels_per_s": value * P ** 3,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this (kernel, shape) from the
    committed ncu --set full capture (profiles/ncu_traffic.json), or None if not captured"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["%s|%s" % (kernel, shape)]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def roofline_pass(step, profiling, args):
    """2 extra steps with every C-ABI launch bracketed by CUDA events on the launch stream;
    the dominant kernel (largest share of the step) is reported against its binding roof."""
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            j = json.load(f)
        peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                 "src": "measured"}
    prof = profiling.enable()
    try:
        nsteps = 2
        for _ in range(nsteps):
            step()
        rows = prof.summary()
    finally:
        profiling.disable()
    total_ms = sum(r["ms"] for r in rows)
    by_kernel = {}
    for r in rows:
        a = by_kernel.setdefault(r["kernel"], {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0})
        a["ms"] += r["ms"]; a["bytes"] += r["bytes"]; a["flops"] += r["flops"]; a["launches"] += r["launches"]
    top = rows[0]
    # fp32 FFMA roof of the CUDA cores: 148 SMs x 128 lanes x 2 flop x 1.965 GHz
    fp32_tflops = 148 * 128 * 2 * 1.965e9 / 1e12
    ach_gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    ach_tf = top["flops"] / (top["ms"] * 1e-3) / 1e12
    hbm_frac = ach_gbs / peaks["hbm_gbs"]
    out = {
        "bound": "hbm", "kernel": top["kernel"], "shape": top["shape"],
        "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_frac,
        "peak_source": peaks["src"], "traffic": ncu_traffic(top["kernel"], top["shape"]),
        "share_of_step": top["ms"] / total_ms if total_ms else None,
        # all shapes of the same kernel: the figure the ncu launch list (profiles/*_ncu_launch_summary.csv)
        # reports for this kernel name
        "kernel_share_of_step": by_kernel[top["kernel"]]["ms"] / total_ms if total_ms else None,
        "avg_launch_ms": top["ms"] / top["launches"],
        "achieved_tflops": ach_tf, "fp32_ffma_peak_tflops": fp32_tflops,
        "frac_of_fp32_ffma": ach_tf / fp32_tflops,
        "step_hbm_frac": (sum(r["bytes"] for r in rows) / (total_ms * 1e-3) / 1e9) / peaks["hbm_gbs"],
        "kernel_ms_per_step": total_ms / nsteps,
        "by_kernel": {k: {"ms_per_step": v["ms"] / nsteps, "launches_per_step": v["launches"] // nsteps,
                          "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0,
                          "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0}
                      for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1]["ms"])},
    }
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        with open(args.profile_out, "w") as f:
            json.dump({"steps": nsteps, "rows": rows[:60]}, f, indent=1)
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
