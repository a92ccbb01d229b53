#!/bin/bash
# First GPU call of the next round: validate the opt-in ring kernels that were written after the
# round-1 GPU budget was spent, then A/B them on the headline bench.
#   gpurun --timeout 600 -- 'bash tools/r2_first.sh'
mkdir -p gpurun_out
NAS3D_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "ring_staged_affine_bwd_apply or ring_staged_pointwise_forward or sample_blocked" 2>&1 | tail -15
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
i=0
for cfg in "X=1" "NAS3D_APPLY_RING=1" "NAS3D_PW_FWD_RING=1" "NAS3D_APPLY_RING=1 NAS3D_PW_FWD_RING=1" \
           "NAS3D_SAMPLE_BLOCK=1" "NAS3D_SAMPLE_BLOCK=1 NAS3D_RING_MIN_LOG2=21"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
      --profile-out gpurun_out/r2_first_prof$i.json > gpurun_out/r2_first_$i.json 2> gpurun_out/r2_first_$i.err
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2_first_$i.json").read().strip().splitlines()[-1])
bk = j["roofline"]["by_kernel"]
print("bench [$cfg]", round(j["value"], 2), round(j["ms_per_step"], 3), round(j["e2e"]["value"], 2),
      {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:10]})
PY
  i=$((i+1))
done
# L2 blocking over samples: does a smaller batch already run faster per patch?
for b in 1 2 4; do
  out=$(timeout 300 python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1)
  echo "batch $b => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), 'patches/s', round(j['ms_per_step'],3), 'ms/step')" "$out")"
done
