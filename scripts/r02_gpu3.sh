#!/bin/bash
# Round 2, third single-GPU call: device Gale-Shapley tests, the GNN bench script at the 15K shape (script validation) and at 100K on one GPU.
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests/test_stable_matching_gpu.py tests/test_host_modules.py -q -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -8 $O/tests.txt
timeout 300 python scripts/bench_gnn.py --config alinet --shape 15K > $O/gnn_alinet_15k_n1.json 2> $O/gnn_alinet_15k.err; echo "alinet 15k rc=$?"; tail -3 $O/gnn_alinet_15k.err; cat $O/gnn_alinet_15k_n1.json
timeout 300 python scripts/bench_gnn.py --config rdgcn --shape 15K > $O/gnn_rdgcn_15k_n1.json 2> $O/gnn_rdgcn_15k.err; echo "rdgcn 15k rc=$?"; tail -3 $O/gnn_rdgcn_15k.err; cat $O/gnn_rdgcn_15k_n1.json
timeout 600 python scripts/bench_gnn.py --config alinet --shape 100K > $O/gnn_alinet_100k_n1.json 2> $O/gnn_alinet_100k.err; echo "alinet 100k rc=$?"; tail -3 $O/gnn_alinet_100k.err; cat $O/gnn_alinet_100k_n1.json
timeout 900 python scripts/bench_gnn.py --config rdgcn --shape 100K > $O/gnn_rdgcn_100k_n1.json 2> $O/gnn_rdgcn_100k.err; echo "rdgcn 100k rc=$?"; tail -3 $O/gnn_rdgcn_100k.err; cat $O/gnn_rdgcn_100k_n1.json
