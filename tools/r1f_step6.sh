#!/bin/bash
# A/B of build variants of the fused 1x1 backward (ring depth / CTAs per SM)
for lib in libnas3d_b200.so libnas3d_b200_pbs4m4.so libnas3d_b200_pbs8m3.so; do
  out=$(NAS3D_LIB=/root/repo/nas_3d_unet_b200/lib/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1f_s6_$lib.json 2>/dev/null | tail -1)
  echo "$lib => $(python -c "import json,sys; j=json.loads(sys.argv[1]); print(round(j['value'],2), round(j['ms_per_step'],3), round(j['e2e']['value'],2), round(j['roofline']['by_kernel']['nas3d_conv1x1_bwd_fused']['ms_per_step'],3))" "$out")"
done
