for fl in "" "--no-flush"; do
for d in 0 1 2 4 5 8 9 15; do
  OEA_DIAG=$d python bench.py --steps 60 --warmup 8 --no-cpu-baseline $fl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('diag=$d flush=%s: score %.1f us step %.1f us' % ('$fl'=='', d['roofline']['kernel_ms_median']*1e3, d['ms_per_step']*1e3))"
done; done
