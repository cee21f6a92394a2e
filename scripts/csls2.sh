#!/bin/bash
mkdir -p gpurun_out/csls2
python -m pytest tests/test_finding_gpu.py tests/test_e2e_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/csls2/tests.txt
python scripts/prof_csls.py 10500 4 > gpurun_out/csls2/csls_15k_plain.txt 2>&1
python scripts/prof_csls.py 70000 3 > gpurun_out/csls2/csls_100k_plain.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/csls2/csls_15k.csv python scripts/prof_csls.py 10500 1 > gpurun_out/csls2/l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/csls2/csls_100k.csv python scripts/prof_csls.py 70000 1 > gpurun_out/csls2/l2.log 2>&1
