#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_s10_prof.json > gpurun_out/r1e_s10.json 2> gpurun_out/r1e_s10.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r1e_s10.json").read().strip().splitlines()[-1])
print("bench", round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
bk = j["roofline"]["by_kernel"]
print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:14]})
a = json.load(open("gpurun_out/r1e_s10_prof.json"))
for r in a["rows"]:
    if "affine" in r["kernel"] and r["ms"] / a["steps"] > 0.12:
        print(r["kernel"][6:], r["shape"], round(r["ms"] / r["launches"], 3), "ms", round(r["bytes"] / r["ms"] / 1e6), "GB/s")
PY
