#!/bin/bash
# Multi-GPU evidence on ONE box:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_multi.sh TAG N [what ...]'
#   what: searched | supernet | inference   (default: all three)
# One process per GPU (torchrun, NCCL); every line lands in gpurun_out/TAG_*_nN.json.
TAG=$1; N=$2; shift; shift
WHAT="${@:-searched supernet inference}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in $WHAT; do
  case "$w" in
    searched)  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-roofline \
                   > gpurun_out/${TAG}_bench_searched128_b8_n$N.json 2> gpurun_out/${TAG}_searched_n$N.err
               echo "searched rc=$?"; tail -1 gpurun_out/${TAG}_bench_searched128_b8_n$N.json | cut -c1-420 ;;
    supernet)  timeout 600 $TR bench.py --gpus $N --workload supernet --batch 1 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline \
                   > gpurun_out/${TAG}_bench_supernet128_b1_n$N.json 2> gpurun_out/${TAG}_supernet_n$N.err
               echo "supernet rc=$?"; tail -1 gpurun_out/${TAG}_bench_supernet128_b1_n$N.json | cut -c1-420 ;;
    inference) timeout 600 $TR tools/bench_inference.py $((N * 2)) \
                   > gpurun_out/${TAG}_bench_inference_n$N.json 2> gpurun_out/${TAG}_inference_n$N.err
               echo "inference rc=$?"; tail -1 gpurun_out/${TAG}_bench_inference_n$N.json | cut -c1-900 ;;
  esac
done
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/${TAG}_topo_n$N.txt
