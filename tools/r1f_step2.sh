#!/bin/bash
# r1f step 2: parity, then the headline bench (+ optional env settings given as arguments)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TAG=${TAG:-r1f_s2}
i=0
for cfg in "${@:-X=1}"; do
env $cfg timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${TAG}_prof$i.json > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_$i.json").read().strip().splitlines()[-1])
print("bench $cfg", round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
bk = j["roofline"]["by_kernel"]
print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:16]})
PY
i=$((i+1))
done
