#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k "regex:pointwise_kernel" -s 2 -c 1 -o gpurun_out/r1e_full_pw_fwd_4_12 -f python tools/conv_micro.py --k 1 --c 4 --cs 12 --s 128 --n 8 --iters 1 --which fwd > /dev/null 2>&1
timeout 600 ncu --clock-control none --set full --import-source on -k "regex:pointwise_kernel" -s 2 -c 1 -o gpurun_out/r1e_full_pw_dgrad_12_4 -f python tools/conv_micro.py --k 1 --c 12 --cs 4 --s 128 --n 8 --relu 1 --iters 1 --which dgrad > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
