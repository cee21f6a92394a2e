#!/bin/bash
# A/B of k_score_sampled_oct build variants (warps per CTA, min CTAs per SM, accumulators in smem) on the bench.
# Build here (no GPU needed):  scripts/ab_score.sh build      Run on the box:  scripts/ab_score.sh run
set -u
VARIANTS="8:2:0 4:5:0 4:6:1 4:8:1"
DIR=openea_b200/_lib/variants
if [ "${1:-run}" = build ]; then
  mkdir -p $DIR
  for v in $VARIANTS; do IFS=: read w b e <<<"$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --compiler-options -fPIC -shared \
      -DOEA_OCT_WARPS=$w -DOEA_OCT_MINB=$b -DOEA_OCT_E_SMEM=$e -o $DIR/liboea_w${w}b${b}e${e}.so openea_b200/csrc/*.cu || exit 1
  done; ls -la $DIR; exit 0
fi
mkdir -p gpurun_out
for wl in bootea_15k bootea_100k; do
for v in $VARIANTS; do IFS=: read w b e <<<"$v"
  OEA_LIB_PATH=$PWD/$DIR/liboea_w${w}b${b}e${e}.so python bench.py --workload $wl --steps 60 --warmup 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl w$w b$b e$e: score %.1f us step %.1f us value %.3e frac %.2f' % (d['roofline']['kernel_ms_median']*1e3, d['ms_per_step']*1e3, d['value'], d['roofline']['frac']))" | tee -a gpurun_out/ab_score.txt
done; done
