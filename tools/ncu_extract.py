"""Key metrics of every launch in an .ncu-rep (read with `ncu -i X --page raw --csv`) as CSV.
usage: ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_extract.py > profiles/x.csv"""
import csv
import sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
     "tcgen05_tf32_ops_pct_of_peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("smsp__inst_executed.sum", "instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(h, n) for h, n in WANT if h in idx]
    w = csv.writer(sys.stdout)
    w.writerow([n for _, n in cols])
    for r in rows[2:]:
        out = []
        for h, n in cols:
            v = r[idx[h]]
            if n == "kernel":
                v = v.replace("void ", "").replace("nas3d::", "").split("(")[0]
            elif n in ("dram_read_MB", "dram_write_MB"):
                u = units[idx[h]]
                f = float(v.replace(",", ""))
                v = "%.1f" % (f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0))
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main()
