#!/bin/bash
# like gpu_env_ab.sh but for the supernet workload (batch 1)
for cfg in "$@"; do
  out=$(env $cfg timeout 300 python bench.py --workload supernet --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1)
  echo "supernet $cfg => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), round(j['ms_per_step'],3))" "$out")"
done
