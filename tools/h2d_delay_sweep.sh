for d in -1 0 10; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --h2d-delay-ms $d 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('delay $d ms: value %.1f  e2e %.1f (%.3f ms)' % (j['value'], j['e2e']['value'], j['e2e']['ms_per_step']))"
done
