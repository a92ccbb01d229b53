#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 1 --which wgrad
python tools/conv_micro.py --c 4 --s 128 --n 8 --dil 2 --which wgrad
python tools/conv_micro.py --c 8 --s 64 --n 8 --dil 1 --which wgrad
python tools/conv_micro.py --c 8 --s 64 --n 8 --dil 2 --which wgrad
python tools/conv_micro.py --c 16 --s 32 --n 8 --dil 1 --which wgrad
python tools/conv_micro.py --c 32 --s 16 --n 8 --dil 1 --which wgrad
