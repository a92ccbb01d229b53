#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/conv_micro.py --k 1 --c 4 --cs 12 --s 128 --n 8 --which wgrad
python tools/conv_micro.py --k 1 --c 12 --cs 4 --s 128 --n 8 --relu 1 --which wgrad
python tools/conv_micro.py --k 1 --c 12 --cs 8 --s 64 --n 8 --relu 1 --which wgrad
for v in 4 2; do
NAS3D_PW_VPT_SFB=$v timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_s12_prof$v.json > gpurun_out/r1e_s12_$v.json 2> gpurun_out/r1e_s12.err
python - <<PY
import json
j = json.loads(open("gpurun_out/r1e_s12_$v.json").read().strip().splitlines()[-1])
print("bench VPT_SFB=$v", round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
bk = j["roofline"]["by_kernel"]
print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:12]})
PY
done
