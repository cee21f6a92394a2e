set -x
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/final/bench_ref_15k.json 2> gpurun_out/final/err.log
python bench.py > gpurun_out/final/bench_15k.json 2>> gpurun_out/final/err.log
python bench.py --workload bootea_100k --steps 40 --warmup 8 > gpurun_out/final/bench_100k.json 2>> gpurun_out/final/err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 80 --csv --log-file gpurun_out/final/launches_15k.csv python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/final/ncu_bench.log 2>&1
tail -3 gpurun_out/final/err.log
