#!/bin/bash
# Round 2, sixth single-GPU call: tensor-core similarity v3 (warp-specialised pipeline) tests + measurement + ncu.
O=gpurun_out/r02f; mkdir -p $O
timeout 300 python -m pytest tests/test_sim_tc_gpu.py -q -p no:cacheprovider -rfEX -x > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -6 $O/tests.txt
timeout 300 python scripts/sim_tc_measure.py 70000 100 > $O/sim_tc_70000.json 2> $O/sim_tc.err; echo "tc measure rc=$?"; cat $O/sim_tc_70000.json; tail -3 $O/sim_tc.err
timeout 300 python scripts/sim_tc_measure.py 40000 300 > $O/sim_tc_40000_d300.json 2>> $O/sim_tc.err; cat $O/sim_tc_40000_d300.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sim_store_tc3 -s 1 -c 1 -o $O/sim_store_tc3 python scripts/sim_tc_measure.py 30000 100 > $O/ncu_tc.log 2>&1
