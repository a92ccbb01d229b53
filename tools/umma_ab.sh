# tcgen05 fwd kernel: warp-specialised (default) vs lock-step, module-surface micro-benchmark, then the
# searched-net step under both
for V in 1 0; do
  echo "== NAS3D_UMMA_WS=$V"
  NAS3D_UMMA_WS=$V python tools/umma_micro.py 2>&1 | cut -c1-160
done
