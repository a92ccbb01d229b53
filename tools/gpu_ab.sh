#!/bin/bash
# A/B helper run on the GPU box: GPU parity suite, then headline + supernet bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/prof_ab.json > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
tail -2 gpurun_out/bench_ab.err | cut -c1-300
timeout 600 python bench.py --workload supernet --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_ab_supernet128.json 2>gpurun_out/bench_ab_supernet128.err
tail -2 gpurun_out/bench_ab_supernet128.err | cut -c1-300
python - <<'PY'
import json
for f in ("bench_ab", "bench_ab_supernet128"):
    try:
        j = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
