#!/bin/bash
# r01 last GPU call: parity tests of the score-family kernels + batch producer, then (if time remains) their first timings
mkdir -p gpurun_out/ext
timeout 170 python -m pytest tests/test_zz_triple_ext_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/ext/tests.txt 2>&1
tail -25 gpurun_out/ext/tests.txt
timeout 60 python scripts/bench_ext.py > gpurun_out/ext/bench_ext.jsonl 2> gpurun_out/ext/bench_ext.err
tail -12 gpurun_out/ext/bench_ext.jsonl
