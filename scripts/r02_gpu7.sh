#!/bin/bash
# Round 2, seventh single-GPU call: the FULL GPU suite on the final code (duo scorer default, TC GEMM in AliNet / RDGCN, TC v3 with 8
# producer warps), smoke(), the tensor-core measurement, the GNN bench at 100K with the TC GEMM.
O=gpurun_out/r02g; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -8 $O/tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 300 python scripts/sim_tc_measure.py 70000 100 > $O/sim_tc_70000.json 2> $O/sim_tc.err; echo "tc measure rc=$?"; cat $O/sim_tc_70000.json
timeout 300 python scripts/sim_tc_measure.py 40000 300 > $O/sim_tc_40000_d300.json 2>> $O/sim_tc.err; cat $O/sim_tc_40000_d300.json
timeout 300 python scripts/bench_gnn.py --config alinet --shape 100K > $O/gnn_alinet_100k_n1_tc.json 2> $O/gnn.err; cat $O/gnn_alinet_100k_n1_tc.json
OEA_GNN_TC=0 timeout 300 python scripts/bench_gnn.py --config alinet --shape 100K > $O/gnn_alinet_100k_n1_cublas.json 2>> $O/gnn.err; cat $O/gnn_alinet_100k_n1_cublas.json
timeout 300 python scripts/bench_gnn.py --config rdgcn --shape 100K > $O/gnn_rdgcn_100k_n1_tc.json 2>> $O/gnn.err; cat $O/gnn_rdgcn_100k_n1_tc.json
tail -3 $O/gnn.err
