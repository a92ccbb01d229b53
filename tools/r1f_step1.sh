#!/bin/bash
# r1f step 1: fused 1x1 backward (dgrad + wgrad in one pass): parity, then A/B on the headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 1; do
NAS3D_PW_FUSED_BWD=$v timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1f_s1_prof$v.json > gpurun_out/r1f_s1_$v.json 2> gpurun_out/r1f_s1_$v.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r1f_s1_$v.json").read().strip().splitlines()[-1])
print("bench FUSED_BWD=$v", round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
bk = j["roofline"]["by_kernel"]
print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:14]})
PY
done
