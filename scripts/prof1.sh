set -x
mkdir -p gpurun_out/prof
python bench.py --steps 200 --warmup 10 > gpurun_out/prof/bench_15k.json 2> gpurun_out/prof/bench_15k.err; tail -3 gpurun_out/prof/bench_15k.err
python bench.py --steps 100 --warmup 10 --workload bootea_100k > gpurun_out/prof/bench_100k.json 2> gpurun_out/prof/bench_100k.err; tail -3 gpurun_out/prof/bench_100k.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 90 --csv --log-file gpurun_out/prof/launches_15k.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/prof/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_sampled -s 8 -c 2 -o gpurun_out/prof/score_sampled_15k python bench.py --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/prof/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sim_tile -c 3 -o gpurun_out/prof/sim_tile_15k python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof/ncu_full2.log 2>&1
ls -la gpurun_out/prof
