#!/bin/bash
for v in 4 2 1; do
  echo "== VPT_MOM=$v"
  NAS3D_PW_VPT_MOM=$v python tools/conv_micro.py --k 1 --c 4 --cs 12 --s 128 --n 8 --which fwd
  NAS3D_PW_VPT_MOM=$v python tools/conv_micro.py --k 1 --c 12 --cs 4 --s 128 --n 8 --relu 1 --which fwd
  NAS3D_PW_VPT_MOM=$v python tools/conv_micro.py --k 1 --c 12 --cs 8 --s 64 --n 8 --relu 1 --which fwd
  NAS3D_PW_VPT_MOM=$v python tools/conv_micro.py --k 1 --c 24 --cs 4 --s 64 --n 8 --relu 1 --which fwd
done
for v in 4 2 1; do
  echo "== VPT_SFB=$v (no moments: via dgrad of transposed? skipped) VPT_BFS=$v"
  NAS3D_PW_VPT_BFS=$v python tools/conv_micro.py --k 1 --c 12 --cs 4 --s 128 --n 8 --relu 1 --which dgrad
  NAS3D_PW_VPT_BFS=$v python tools/conv_micro.py --k 1 --c 4 --cs 12 --s 128 --n 8 --which dgrad
done
