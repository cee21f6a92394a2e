#!/bin/bash
# Multi-GPU call (gpurun --gpus N): NCCL / peer-memory tests, the bench at N ranks in both exchange transports, Hits@k of the
# stale-replica and exact-parity modes.   usage: bash scripts/r02_gpuN.sh N [tests]
N=${1:-2}; O=gpurun_out/r02n$N; mkdir -p $O
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt
  tail -15 $O/tests.txt
fi
run 29511 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_100k_p2p.json 2> $O/bench_100k_p2p.err; echo "bench 100k p2p rc=$?"; tail -2 $O/bench_100k_p2p.err
OEA_XCHG_MODE=nccl run 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-secondary > $O/bench_100k_nccl.json 2> $O/bench_100k_nccl.err; echo "bench 100k nccl rc=$?"
run 29513 bench.py --gpus $N --steps 50 --warmup 5 --workload bootea_15k --no-secondary > $O/bench_15k_p2p.json 2> $O/bench_15k_p2p.err; echo "bench 15k p2p rc=$?"
OEA_XCHG_MODE=nccl run 29514 bench.py --gpus $N --steps 50 --warmup 5 --workload bootea_15k --no-secondary > $O/bench_15k_nccl.json 2> $O/bench_15k_nccl.err; echo "bench 15k nccl rc=$?"
run 29515 scripts/hits_multigpu.py --mode stale --scaling strong > $O/hits_stale_strong.json 2> $O/hits.err; echo "hits stale strong rc=$?"
run 29516 scripts/hits_multigpu.py --mode stale --scaling weak > $O/hits_stale_weak.json 2>> $O/hits.err; echo "hits stale weak rc=$?"
run 29517 scripts/hits_multigpu.py --mode exact > $O/hits_exact.json 2>> $O/hits.err; echo "hits exact rc=$?"
for f in $O/*.json; do echo "== $f"; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    if 'at' in d: print(d)
    else: print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d.get('collective'), 'share', d['roofline']['kernel_share_of_step'], 'csls', (d.get('csls') or {}).get('value'))
except Exception as e: print('unreadable', e)
"; done
ls -la $O
