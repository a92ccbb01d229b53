#!/bin/bash
# last check of the round on HEAD: GPU tests, smoke(), the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py --profile-out gpurun_out/r1f2_kernels_per_shape.json > gpurun_out/r1f2_bench_searched128_b8_n1.json 2> gpurun_out/r1f2_bench.err
tail -1 gpurun_out/r1f2_bench_searched128_b8_n1.json | cut -c1-400
