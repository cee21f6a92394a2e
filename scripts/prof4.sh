#!/bin/bash
mkdir -p gpurun_out/prof4
python scripts/prof_csls.py 10500 4 > gpurun_out/prof4/csls_15k_plain.txt 2>&1
python scripts/prof_csls.py 70000 3 > gpurun_out/prof4/csls_100k_plain.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof4/csls_15k.csv python scripts/prof_csls.py 10500 1 > gpurun_out/prof4/l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof4/csls_100k.csv python scripts/prof_csls.py 70000 1 > gpurun_out/prof4/l2.log 2>&1
