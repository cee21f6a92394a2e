set -x
mkdir -p gpurun_out/prof2
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-flush > gpurun_out/prof2/warm_v2.json 2> gpurun_out/prof2/err.log
OEA_SCORE_V1=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-flush > gpurun_out/prof2/warm_v1.json 2>> gpurun_out/prof2/err.log
ncu --set full --clock-control none --import-source on -k regex:k_score_sampled_oct -s 8 -c 2 -o gpurun_out/prof2/score_oct_15k python bench.py --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/prof2/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rowopt_pair -s 8 -c 2 -o gpurun_out/prof2/rowopt_15k python bench.py --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/prof2/ncu2.log 2>&1
ls -la gpurun_out/prof2
