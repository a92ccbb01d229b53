# stride-1 weight gradient: TMA tile ring (default) vs the cp.async kernel, searched-net shapes
for V in 1 0; do
  echo "== NAS3D_S1_WGRAD_TMA=$V"
  for cfg in "4 128 1" "4 128 2" "8 64 1" "8 64 2" "16 32 1" "16 32 2"; do
    set -- $cfg
    NAS3D_S1_WGRAD_TMA=$V python tools/conv_micro.py --c $1 --s $2 --dil $3 --which wgrad 2>&1 | cut -c1-170
  done
done
