#!/bin/bash
# Micro call with the round's last GPU seconds: the shared-memory-transposed epilogue of the tensor-core kernel (never run before).
O=gpurun_out/r02i; mkdir -p $O
export OEA_SIM_TC_EPI=smem
timeout 40 python -m pytest tests/test_sim_tc_gpu.py -q -x -p no:cacheprovider -k "v3 or materialised" > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -3 $O/tests.txt
timeout 30 python scripts/sim_tc_measure.py 70000 100 > $O/sim_tc_70000_smem_epilogue.json 2> $O/sim_tc.err; echo "measure rc=$?"; cat $O/sim_tc_70000_smem_epilogue.json; tail -2 $O/sim_tc.err
