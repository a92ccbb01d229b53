#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 1 0; do
  echo "NAS3D_TMA_MERGED=$m"
  NAS3D_TMA_MERGED=$m python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1
  NAS3D_TMA_MERGED=$m python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 2 --which fwd,dgrad
done
python tools/conv_micro.py --c 8 --s 64 --n 8 --dil 1
python tools/conv_micro.py --c 4 --cs 12 --s 128 --n 8 --stride 2
python tools/conv_micro.py --c 8 --s 64 --n 8 --stride 2
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_s3_prof.json > gpurun_out/r1e_s3.json 2> gpurun_out/r1e_s3.err
python - <<'PY'
import json
for f in ("r1e_s3",):
    try:
        j = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
        bk = j["roofline"]["by_kernel"]
        print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:10]})
    except Exception as e:
        print(f, "ERR", e)
PY
