#!/bin/bash
# Round 2, fourth single-GPU call: tensor-core similarity tests + measurement, Hits parity (regenerated oracle curves),
# the bench line, ncu of the duo step kernel at both shapes + launch list.
O=gpurun_out/r02d; mkdir -p $O
timeout 900 python -m pytest tests/test_sim_tc_gpu.py tests/test_hits_parity_gpu.py tests/test_stable_matching_gpu.py -q -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -12 $O/tests.txt
timeout 600 python scripts/sim_tc_measure.py 70000 100 > $O/sim_tc_70000.json 2> $O/sim_tc.err; echo "tc measure rc=$?"; cat $O/sim_tc_70000.json; tail -3 $O/sim_tc.err
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bootea100k.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/bench_under_ncu.log 2>&1
for wl in bootea_100k bootea_15k; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_step_sampled_duo -s 6 -c 1 -o $O/step_duo_$wl python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_$wl.log 2>&1
done
ls -la $O
