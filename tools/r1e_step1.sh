#!/bin/bash
# r1e step 1 on ONE B200: FFMA vs FFMA2 micro-benchmark, GPU parity suite, headline bench with
# the packed-FMA kernels and (A/B) the scalar-FFMA build of the same sources, ncu full captures
# of the C=4 conv kernels.
mkdir -p gpurun_out
tools/micro/ffma2_micro | tee gpurun_out/r1e_ffma2_micro.json
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_ab_ffma2_prof.json > gpurun_out/r1e_ab_ffma2.json 2> gpurun_out/r1e_ab_ffma2.err
NAS3D_LIB=$PWD/nas_3d_unet_b200/lib/libnas3d_b200_noffma2.so timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/r1e_ab_noffma2_prof.json > gpurun_out/r1e_ab_noffma2.json 2> gpurun_out/r1e_ab_noffma2.err
python - <<'PY'
import json
for f in ("r1e_ab_ffma2", "r1e_ab_noffma2"):
    try:
        j = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["value"], 2), round(j["ms_per_step"], 2), round(j["e2e"]["value"], 2), j.get("gpu_launches"))
        bk = j["roofline"]["by_kernel"]
        print("   ", {k[6:]: round(v["ms_per_step"], 2) for k, v in list(bk.items())[:8]})
    except Exception as e:
        print(f, "ERR", e)
PY
NCU="ncu --clock-control none --kernel-name-base demangled"
for k in "conv3_s1_kernel<4, 1" "wgrad3_s1_kernel<1, 12"; do
  tag=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
  timeout 400 $NCU --set full --import-source on -k "regex:$k" -s 0 -c 2 -o gpurun_out/r1e_full_$tag -f python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > /dev/null 2>&1
  ls -la gpurun_out/r1e_full_$tag.ncu-rep | awk '{print $5, $9}'
done
