#!/bin/bash
# Round r1f evidence run on ONE B200: tests, bench lines of every config, ncu launch list and full
# captures of the kernels that changed in r1f.  Outputs land in gpurun_out/ (copied to profiles/ here).
TAG=${1:-r1f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/${TAG}_kernels_per_shape.json > gpurun_out/${TAG}_bench_searched128_b8_n1.json 2> gpurun_out/${TAG}_bench.err
tail -1 gpurun_out/${TAG}_bench_searched128_b8_n1.json | cut -c1-300
timeout 600 python bench.py --workload supernet --batch 1 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_supernet128_b1.json
timeout 600 python bench.py --workload supernet --batch 1 --patch 64 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_supernet64_b1.json
timeout 600 python bench.py --patch 64 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_searched64_b8.json
timeout 600 python tools/bench_inference.py > gpurun_out/${TAG}_bench_inference.json 2>/dev/null
for f in supernet128_b1 supernet64_b1 searched64_b8 inference; do echo $f; tail -1 gpurun_out/${TAG}_bench_$f.json | cut -c1-200; done
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -s 1200 -c 1900 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --graph off --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > /dev/null 2>&1
wc -l gpurun_out/${TAG}_launches.csv
for k in pw_bwd_fused_kernel bwd_reduce_kernel affine_sum_bwd_apply_kernel; do
  timeout 600 $NCU --set full --import-source on -k regex:$k -s 0 -c 3 -o gpurun_out/${TAG}_full_$k -f python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > /dev/null 2>&1
  ls -la gpurun_out/${TAG}_full_$k.ncu-rep | awk '{print $5, $9}'
done
