#!/bin/bash
# Round 2, first single-GPU call: full GPU test suite (quarantine retired), the new default bench line (100K + 15K block),
# the reference arm, the launch list and the full ncu capture of the headline kernel at both shapes, the gather ubench.
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt
tail -25 $O/tests.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_100k.json 2> $O/ref.err; echo "ref rc=$?"
timeout 300 python bench.py --impl reference --workload bootea_15k --steps 20 --warmup 5 > $O/ref_15k.json 2>> $O/ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bootea100k.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/bench_under_ncu.log 2>&1
for wl in bootea_100k bootea_15k; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step_sampled_oct -s 6 -c 1 -o $O/step_oct_$wl python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_$wl.log 2>&1
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_gather.bin scripts/ubench_gather.cu && timeout 180 scripts/ubench_gather.bin > $O/ubench_gather.txt 2>&1
ls -la $O
