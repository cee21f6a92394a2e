#!/bin/bash
# r01 late profile: fused step kernel + short-K similarity kernel
mkdir -p gpurun_out/prof3
python -m pytest tests/test_finding_gpu.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/prof3/tests.txt
for wl in bootea_15k bootea_100k; do
  python bench.py --workload $wl --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/prof3/bench_$wl.json 2> gpurun_out/prof3/bench_$wl.err
done
OEA_SIM_NO_SHORTK=1 python bench.py --workload bootea_100k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/prof3/bench_100k_noshortk.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof3/launches_15k.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof3/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_sampled_oct -s 4 -c 1 -o gpurun_out/prof3/step_oct_15k python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof3/ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_store_shortk -c 1 -o gpurun_out/prof3/sim_shortk_15k python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof3/ncu_full2.log 2>&1
ls -la gpurun_out/prof3
