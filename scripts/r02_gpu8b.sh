#!/bin/bash
# Second (short) 8-GPU call: throughput of the accurate multi-GPU modes, Hits of delta-sum under weak scaling, the lifecycle tests.
O=gpurun_out/r02n8b; mkdir -p $O
run() { local n=$1 port=$2; shift 2; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
run 8 29701 scripts/bench_multi_modes.py --workload bootea_100k > $O/modes_100k_n8.json 2> $O/modes.err; echo "modes 100k n8 rc=$?"; tail -c 900 $O/modes_100k_n8.json
run 8 29702 scripts/bench_multi_modes.py --workload bootea_15k > $O/modes_15k_n8.json 2>> $O/modes.err; echo "modes 15k n8 rc=$?"; tail -c 900 $O/modes_15k_n8.json
run 2 29703 scripts/bench_multi_modes.py --workload bootea_100k > $O/modes_100k_n2.json 2>> $O/modes.err; echo "modes 100k n2 rc=$?"; tail -c 700 $O/modes_100k_n2.json
run 8 29704 scripts/hits_multigpu.py --mode delta --scaling weak --seeds 11 12 > $O/hits_delta_weak_n8.json 2> $O/hits.err; echo "hits delta weak rc=$?"; tail -c 700 $O/hits_delta_weak_n8.json
run 2 29705 scripts/hits_multigpu.py --mode delta --scaling weak --seeds 11 12 > $O/hits_delta_weak_n2.json 2>> $O/hits.err; echo "hits delta weak n2 rc=$?"; tail -c 700 $O/hits_delta_weak_n2.json
timeout 400 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -rfEX -k lifecycle > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -5 $O/tests.txt
tail -5 $O/modes.err
