#!/bin/bash
# Round 2, last single-GPU call: tensor-core kernel with the coalesced epilogue (tests + measurement), the final default bench line.
O=gpurun_out/r02h; mkdir -p $O
timeout 300 python -m pytest tests/test_sim_tc_gpu.py tests/test_gnn_gpu.py -q -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -4 $O/tests.txt
timeout 300 python scripts/sim_tc_measure.py 70000 100 > $O/sim_tc_70000.json 2> $O/sim_tc.err; echo "tc measure rc=$?"; cat $O/sim_tc_70000.json; tail -2 $O/sim_tc.err
timeout 300 python scripts/sim_tc_measure.py 40000 300 > $O/sim_tc_40000_d300.json 2>> $O/sim_tc.err; cat $O/sim_tc_40000_d300.json
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -2 $O/bench_default.err
