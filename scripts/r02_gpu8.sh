#!/bin/bash
# The 8-GPU call (gpurun --gpus 8): bench at N = 8 / 4 / 2 / 1 on the same box (both exchange transports at N = 8), Hits@k of the
# multi-GPU modes at N = 8, BASELINE configs 4 (AliNet 100K on 4 GPUs) and 5 (RDGCN 100K + 100000^2 CSLS on 8 GPUs), the NCCL tests.
O=gpurun_out/r02n8; mkdir -p $O
run() { local n=$1 port=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
nvidia-smi topo -m > $O/topo.txt 2>&1
run 8 29601 bench.py --gpus 8 --steps 50 --warmup 5 > $O/bench_100k_n8_p2p.json 2> $O/bench_100k_n8_p2p.err; echo "bench 100k n8 p2p rc=$?"; tail -2 $O/bench_100k_n8_p2p.err
run 8 29602 bench.py --gpus 8 --steps 50 --warmup 5 --workload bootea_15k --no-secondary > $O/bench_15k_n8_p2p.json 2> $O/bench_15k_n8_p2p.err; echo "bench 15k n8 p2p rc=$?"
OEA_XCHG_MODE=nccl run 8 29603 bench.py --gpus 8 --steps 50 --warmup 5 --no-secondary > $O/bench_100k_n8_nccl.json 2> $O/bench_100k_n8_nccl.err; echo "bench 100k n8 nccl rc=$?"
OEA_XCHG_MODE=nccl run 8 29604 bench.py --gpus 8 --steps 50 --warmup 5 --workload bootea_15k --no-secondary > $O/bench_15k_n8_nccl.json 2> $O/bench_15k_n8_nccl.err; echo "bench 15k n8 nccl rc=$?"
for n in 4 2; do
  run $n 2961$n bench.py --gpus $n --steps 50 --warmup 5 --no-secondary > $O/bench_100k_n$n.json 2> $O/bench_100k_n$n.err; echo "bench 100k n$n rc=$?"
  run $n 2962$n bench.py --gpus $n --steps 50 --warmup 5 --workload bootea_15k --no-secondary > $O/bench_15k_n$n.json 2> $O/bench_15k_n$n.err; echo "bench 15k n$n rc=$?"
done
timeout 200 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_100k_n1.json 2> $O/bench_100k_n1.err; echo "bench 100k n1 rc=$?"
timeout 200 python bench.py --steps 50 --warmup 5 --workload bootea_15k --no-secondary --no-cpu-baseline > $O/bench_15k_n1.json 2> $O/bench_15k_n1.err; echo "bench 15k n1 rc=$?"
run 8 29631 scripts/hits_multigpu.py --mode stale --scaling weak --seeds 11 12 > $O/hits_stale_weak.json 2> $O/hits.err; echo "hits stale weak rc=$?"
run 8 29632 scripts/hits_multigpu.py --mode stale --scaling strong --seeds 11 12 > $O/hits_stale_strong.json 2>> $O/hits.err; echo "hits stale strong rc=$?"
run 8 29633 scripts/hits_multigpu.py --mode exact --seeds 11 12 > $O/hits_exact.json 2>> $O/hits.err; echo "hits exact rc=$?"
run 8 29634 scripts/hits_multigpu.py --mode delta --scaling strong --seeds 11 12 > $O/hits_delta_epoch.json 2>> $O/hits.err; echo "hits delta rc=$?"
run 8 29641 scripts/bench_gnn.py --config rdgcn > $O/gnn_rdgcn_100k_n8.json 2> $O/gnn_rdgcn_n8.err; echo "rdgcn n8 rc=$?"; tail -2 $O/gnn_rdgcn_n8.err
run 4 29642 scripts/bench_gnn.py --config alinet > $O/gnn_alinet_100k_n4.json 2> $O/gnn_alinet_n4.err; echo "alinet n4 rc=$?"; tail -2 $O/gnn_alinet_n4.err
timeout 400 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -rfEX > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -5 $O/tests.txt
for f in $O/bench_*.json; do python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); c=d.get('collective') or {}
    print('$f'.split('/')[-1], 'value %.3e'%d['value'], 'us/step %.1f'%(d['ms_per_step']*1e3), 'share %.3f'%d['roofline']['kernel_share_of_step'], c.get('mode'), c.get('epoch_steps'), 'csls', (d.get('csls') or {}).get('value'))
except Exception as e: print('$f', 'unreadable', e)
"; done
for f in $O/hits_*.json $O/gnn_*.json; do echo "== $f"; tail -c 900 $f; echo; done
ls -la $O
