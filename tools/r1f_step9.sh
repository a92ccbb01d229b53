#!/bin/bash
# ring-staged affine_sum_fwd: parity of the variant path, then the headline bench with it on
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ring_staged_affine" 2>&1 | tail -3
out=$(NAS3D_AFFINE_RING=1 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1f_s9_prof.json 2>/dev/null | tail -1)
echo "$out" > gpurun_out/r1f_s9.json
python -c "import json,sys; j=json.loads(sys.argv[1]); bk=j['roofline']['by_kernel']; print('AFFINE_RING=1', round(j['value'],2), round(j['ms_per_step'],3), round(j['e2e']['value'],2), round(bk['nas3d_affine_sum_fwd']['ms_per_step'],3), round(bk['nas3d_affine_sum_bwd_reduce']['ms_per_step'],3))" "$out"
