#!/bin/bash
mkdir -p gpurun_out
scripts/ubench_ffma2.bin > gpurun_out/ubench_ffma2.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/f32x2_tests.txt
rm -f gpurun_out/fuse_summary.txt
for wl in bootea_15k bootea_100k; do
  python bench.py --workload $wl --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/fuse_bench_$wl.json 2> gpurun_out/fuse_bench_$wl.err
  python - <<PY | tee -a gpurun_out/fuse_summary.txt
import json
d=json.loads(open("gpurun_out/fuse_bench_$wl.json").read().strip().splitlines()[-1])
r=d["roofline"]; s=r["score_kernel_alone"]
print("$wl fused: step %.1f us kernel %.1f us value %.3e frac %.2f | two-launch: step %.1f us score %.1f us frac %.2f | graph %.1f us/step | e2e %.3e | csls %s" % (
  d["ms_per_step"]*1e3, r["kernel_ms_median"]*1e3, d["value"], r["frac"], s["two_launch_ms_per_step"]*1e3, s["ms_median"]*1e3, s["frac"],
  d["epoch_graph"]["ms_per_step"]*1e3, d["e2e"]["value"], json.dumps(d["csls"])))
PY
done
