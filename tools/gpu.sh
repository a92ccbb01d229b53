#!/bin/bash
# The ONE script run on the GPU box (replaces the per-step scripts of round 1):
#   gpurun --timeout 900 -- 'bash tools/gpu.sh TAG cmd [cmd ...]'
# Every cmd writes its full log under gpurun_out/TAG_*; nothing is piped through tail only.
#   test            whole GPU suite, no -x, -rA log kept
#   test:EXPR       pytest -k EXPR
#   smoke           __graft_entry__.smoke()
#   bench           headline bench line (20 steps) + per-(kernel, shape) table
#   ab:"ENV=V ..."  short bench (10 steps, no cpu baseline) under the given environment
#   configs         one line per BASELINE.json config (#1, #2, #3, #5; #4 is `bench`)
#   refarm          bench.py --impl reference (CPU arm) with 2 steps
#   launches        ncu launch list of one eager step (gpu__time_duration.sum)
#   occ             ncu occupancy / issue / pipe metrics of every launch of one eager step, by kernel
#   full:REGEX      ncu --set full capture of kernels matching REGEX (3 launches)
#   py:"ARGS"       python ARGS   (micro-benchmarks under tools/)
TAG=$1; shift
mkdir -p gpurun_out
NCU="ncu --clock-control none"
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rf = j.get("roofline") or {}
    fam = rf.get("by_kernel") or {}
    print("  value %.2f %s  %.3f ms/step  e2e %.2f  launches %s  summed_roofline %s" % (
        j["value"], j["unit"], j["ms_per_step"], (j.get("e2e") or {}).get("value", float("nan")),
        j.get("gpu_launches"), rf.get("summed")))
    print("  " + str({k[6:]: round(v["ms_per_step"], 2) for k, v in list(fam.items())[:12]}))
except Exception as e:
    print("  FAILED to parse", sys.argv[1], e)
PY
}
i=0
for cmd in "$@"; do
  arg="${cmd#*:}"; [ "$arg" == "$cmd" ] && arg=""
  case "$cmd" in
    test)    timeout 900 python -m pytest tests -m gpu -q -rA > gpurun_out/${TAG}_pytest_gpu.log 2>&1
             echo "pytest rc=$?"; grep -v "^PASSED" gpurun_out/${TAG}_pytest_gpu.log | tail -30 | cut -c1-400 ;;
    test:*)  timeout 900 python -m pytest tests -m gpu -q -rA -k "$arg" > gpurun_out/${TAG}_pytest_k$i.log 2>&1
             echo "pytest -k '$arg' rc=$?"; grep -v "^PASSED" gpurun_out/${TAG}_pytest_k$i.log | tail -40 | cut -c1-400 ;;
    smoke)   timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 ;;
    bench)   timeout 900 python bench.py --steps 20 --warmup 5 --drivers-loop --profile-out gpurun_out/${TAG}_kernels_per_shape.json \
                 > gpurun_out/${TAG}_bench_searched128_b8_n1.json 2> gpurun_out/${TAG}_bench.err
             echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err | cut -c1-300; summ gpurun_out/${TAG}_bench_searched128_b8_n1.json ;;
    ab:*)    env $arg timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
                 --profile-out gpurun_out/${TAG}_ab${i}_prof.json > gpurun_out/${TAG}_ab$i.json 2> gpurun_out/${TAG}_ab$i.err
             echo "ab [$arg] rc=$?"; tail -2 gpurun_out/${TAG}_ab$i.err | cut -c1-300; summ gpurun_out/${TAG}_ab$i.json ;;
    configs) timeout 600 python bench.py --workload supernet --batch 1 --patch 64 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>gpurun_out/${TAG}_cfg1.err | tail -1 > gpurun_out/${TAG}_bench_cfg1_supernet64_b1.json
             timeout 600 python bench.py --patch 64 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>gpurun_out/${TAG}_cfg2.err | tail -1 > gpurun_out/${TAG}_bench_cfg2_searched64_b8.json
             timeout 600 python bench.py --workload supernet --batch 1 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>gpurun_out/${TAG}_cfg3.err | tail -1 > gpurun_out/${TAG}_bench_cfg3_supernet128_b1.json
             timeout 600 python tools/bench_inference.py > gpurun_out/${TAG}_bench_cfg5_inference.json 2>gpurun_out/${TAG}_cfg5.err
             for f in cfg1_supernet64_b1 cfg2_searched64_b8 cfg3_supernet128_b1; do echo $f; summ gpurun_out/${TAG}_bench_$f.json; done
             echo cfg5_inference; tail -1 gpurun_out/${TAG}_bench_cfg5_inference.json | cut -c1-400 ;;
    refarm)  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_refarm.err
             tail -1 gpurun_out/${TAG}_bench_reference_arm.json | cut -c1-600 ;;
    launches) timeout 900 $NCU --metrics gpu__time_duration.sum -s 1200 -c 1900 --csv --log-file gpurun_out/${TAG}_launches.csv \
                 python bench.py --graph off --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > /dev/null 2>&1
             python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_ncu_launch_summary.csv; head -12 gpurun_out/${TAG}_ncu_launch_summary.csv ;;
    occ)     timeout 900 $NCU --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__occupancy_limit_warps,launch__registers_per_thread,launch__block_size,launch__grid_size,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
                 -s 1200 -c 1900 --csv --page raw --log-file gpurun_out/${TAG}_occ.csv \
                 python bench.py --graph off --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > /dev/null 2>&1
             python tools/occ_summary.py gpurun_out/${TAG}_occ.csv > gpurun_out/${TAG}_occ_summary.txt; head -50 gpurun_out/${TAG}_occ_summary.txt ;;
    full:*)  timeout 600 $NCU --set full --import-source on -k regex:$arg -s 0 -c 3 -o gpurun_out/${TAG}_full_$arg -f \
                 python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > /dev/null 2>&1
             ls -la gpurun_out/${TAG}_full_$arg.ncu-rep | awk '{print $5, $9}' ;;
    py:*)    timeout 600 python $arg 2>&1 | tail -40 | cut -c1-300 ;;
    *)       echo "unknown cmd $cmd" ;;
  esac
  i=$((i+1))
done
