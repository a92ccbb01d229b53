#!/bin/bash
# First GPU call of round 2: whole GPU suite WITHOUT -x (log kept), the gated tests of the opt-in
# variants, then an A/B of each variant on the headline bench and the batch sweep (L2 blocking).
#   gpurun --timeout 900 -- 'bash tools/r2_first.sh'
mkdir -p gpurun_out
NAS3D_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests -m gpu -q -rA > gpurun_out/r2a_pytest_full.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest_full.log | cut -c1-300
i=0
for cfg in "X=1" "NAS3D_APPLY_RING=1" "NAS3D_PW_FWD_RING=1" "NAS3D_APPLY_RING=1 NAS3D_PW_FWD_RING=1" \
           "NAS3D_SAMPLE_BLOCK=1" "NAS3D_SAMPLE_BLOCK=1 NAS3D_RING_MIN_LOG2=21"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
      --profile-out gpurun_out/r2a_prof$i.json > gpurun_out/r2a_bench_$i.json 2> gpurun_out/r2a_bench_$i.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2a_bench_$i.json").read().strip().splitlines()[-1])
    bk = j["roofline"]["by_kernel"]
    print("bench [$cfg]", round(j["value"], 2), round(j["ms_per_step"], 3), round(j["e2e"]["value"], 2),
          {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:10]})
except Exception as e:
    print("bench [$cfg] FAILED", e)
PY
  i=$((i+1))
done
for b in 1 2 4; do
  out=$(timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1)
  echo "$out" > gpurun_out/r2a_batch$b.json
  echo "batch $b => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), 'patches/s', round(j['ms_per_step'],3), 'ms/step')" "$out")"
done
