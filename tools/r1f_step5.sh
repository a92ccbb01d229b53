#!/bin/bash
# env A/B of the stream settings after the 1x1 wgrad moved into the fused main-stream kernel
for cfg in "X=1" "NAS3D_WGRAD_STREAMS=2" "NAS3D_WGRAD_STREAM=0" "NAS3D_PW_VPT_SFB=2"; do
  out=$(env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1)
  echo "$cfg => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), round(j['ms_per_step'],3), round(j['e2e']['value'],2))" "$out")"
done
