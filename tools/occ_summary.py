"""per-kernel table (time-sorted) of occupancy limiters from `ncu --page raw --csv` over one eager step.
usage: python tools/occ_summary.py gpurun_out/TAG_occ.csv"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[start + 2:]

    def f(r, k):
        try:
            return float(r[ix[k]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
    agg = defaultdict(lambda: dict(n=0, t=0.0, occ=0.0, issue=0.0, fma=0.0, dram=0.0))
    for r in body:
        name = r[ix["Kernel Name"]].replace("void ", "").replace("nas3d::", "")
        name = name.split("(")[0] if "<" not in name else name[:name.rfind(">") + 1]
        key = (name, int(f(r, "launch__block_size")), int(f(r, "launch__registers_per_thread")),
               int(f(r, "launch__occupancy_limit_registers")), int(f(r, "launch__occupancy_limit_shared_mem")),
               int(f(r, "launch__occupancy_limit_warps")))
        t = f(r, "gpu__time_duration.sum")
        a = agg[key]
        a["n"] += 1
        a["t"] += t
        a["occ"] += t * f(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["issue"] += t * f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        a["fma"] += t * f(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")
        a["dram"] += t * f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    tot = sum(a["t"] for a in agg.values())
    print("total %.3f ms over %d launches" % (tot / 1e3 if tot > 1e4 else tot, sum(a["n"] for a in agg.values())))
    print("%-62s %5s %4s %9s %4s | %8s %5s | %5s %5s %5s %5s" % (
        "kernel", "block", "regs", "lim r/s/w", "n", "time", "share", "occ%", "iss%", "fma%", "dram%"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"])[:60]:
        name, blk, regs, lr, ls, lw = key
        t = a["t"]
        print("%-62s %5d %4d %3d/%2d/%2d %4d | %8.1f %4.1f%% | %5.1f %5.1f %5.1f %5.1f" % (
            name[:62], blk, regs, lr, ls, lw, a["n"], t, 100 * t / tot, a["occ"] / t, a["issue"] / t,
            a["fma"] / t, a["dram"] / t))


if __name__ == "__main__":
    main(sys.argv[1])
