#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/umma_micro.py 2>&1 | tail -12
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_s7_prof.json > gpurun_out/r1e_s7.json 2> gpurun_out/r1e_s7.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r1e_s7.json").read().strip().splitlines()[-1])
print("bench", round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
bk = j["roofline"]["by_kernel"]
print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:14]})
PY
