for e in "X=1" "NAS3D_UMMA_SPLIT_K=0"; do
  env $e python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(\"searched\", \"$e\", round(j[\"ms_per_step\"],3), \"ms\", round(j[\"value\"],1))"
  for p in 64 128; do env $e python bench.py --workload supernet --batch 1 --patch $p --steps 10 --warmup 3 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(\"supernet\", \"$e\", $p, round(j[\"ms_per_step\"],2), \"ms\")"; done
  env $e python bench.py --patch 64 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(\"searched64\", \"$e\", round(j[\"ms_per_step\"],3), \"ms\")"
done
