#!/bin/bash
# usage: tools/gpu_env_ab.sh "ENV1=a ENV2=b" "ENV1=c" ...   -> one short bench line per setting
for cfg in "$@"; do
  out=$(env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1)
  echo "$cfg => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), round(j['ms_per_step'],3))" "$out")"
done
