#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --set full --import-source on -k "regex:wgrad3_s1" -s 2 -c 1 -o gpurun_out/r1e_full_wgrad_c4_v2 -f python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1 --iters 1 --which wgrad > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
