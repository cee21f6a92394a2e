#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out.
#   gpurun --timeout 900 -- 'bash scripts/r02_first_gpu_call.sh'          (1 GPU)
#   gpurun --gpus 2 --timeout 600 -- 'bash scripts/r02_first_gpu_call.sh multi'   (the NCCL parity tests)
mkdir -p gpurun_out/r02a
if [ "$1" = "multi" ]; then
  python -m pytest tests/test_multigpu.py tests/test_parallel_gloo.py -q -m gpu -p no:cacheprovider > gpurun_out/r02a/multigpu.txt 2>&1
  tail -15 gpurun_out/r02a/multigpu.txt
  exit 0
fi
# 1. the tests marked first_hw_run (collected last) + the whole suite in front of them
python -m pytest tests -q -m gpu -rxX -p no:cacheprovider > gpurun_out/r02a/tests.txt 2>&1
tail -60 gpurun_out/r02a/tests.txt
# 2. the bench lines, 15K and 100K
python bench.py > gpurun_out/r02a/bench_15k.json 2> gpurun_out/r02a/bench.err
python bench.py --workload bootea_100k --steps 40 --warmup 8 > gpurun_out/r02a/bench_100k.json 2>> gpurun_out/r02a/bench.err
#    + the launch list of the same command (kernel share of the step; the one-launch step has no list from round 1)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a/launches_bootea15k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02a/bench_under_ncu.log 2>&1
# 3. the approaches through the lifecycle (AliNet's epoch was 92 ms with host-side sampling; BootEA's iteration with host bootstrapping)
python scripts/bench_approaches.py > gpurun_out/r02a/approaches.json 2>> gpurun_out/r02a/bench.err
python scripts/bench_ext.py > gpurun_out/r02a/score_family.jsonl 2>> gpurun_out/r02a/bench.err
python scripts/bench_fed_grouped.py > gpurun_out/r02a/fed_grouped.jsonl 2>> gpurun_out/r02a/bench.err
python scripts/bench_weighted.py > gpurun_out/r02a/weighted.jsonl 2>> gpurun_out/r02a/bench.err
# 4. ncu of the score family's scorer (TransD limited loss is the heaviest instantiation)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_model_fed -s 20 -c 1 -o gpurun_out/r02a/model_fed python scripts/bench_ext.py > gpurun_out/r02a/ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_fed_grouped -s 4 -c 1 -o gpurun_out/r02a/step_fed_grouped python scripts/bench_fed_grouped.py > gpurun_out/r02a/ncu_fed.log 2>&1
# 5. what bounds K1: gather throughput vs rows in flight, dependent-chain latency, red.v4 throughput, grid-barrier cost
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_gather.bin scripts/ubench_gather.cu && timeout 180 scripts/ubench_gather.bin > gpurun_out/r02a/ubench_gather.txt 2>&1
ls -la gpurun_out/r02a
