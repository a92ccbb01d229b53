#!/bin/bash
# ncu full captures of the C=4 128^3 conv kernels through the C-ABI micro driver + plain timings
mkdir -p gpurun_out
python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1 | tee gpurun_out/r1e_conv_micro.jsonl
python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 2 | tee -a gpurun_out/r1e_conv_micro.jsonl
python tools/conv_micro.py --c 8 --s 64 --n 8 --dil 1 | tee -a gpurun_out/r1e_conv_micro.jsonl
timeout 600 ncu --clock-control none --set full --import-source on -k "regex:conv3_s1|wgrad3_s1" -s 4 -c 3 -o gpurun_out/r1e_full_conv_c4 -f python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
