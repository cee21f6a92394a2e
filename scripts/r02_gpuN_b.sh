#!/bin/bash
# Multi-GPU follow-up: the BootEA lifecycle test and the delta-sum replica mode's Hits@k.   usage: bash scripts/r02_gpuN_b.sh N
N=${1:-2}; O=gpurun_out/r02n${N}b; mkdir -p $O
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multigpu.py -q -m gpu -p no:cacheprovider -rfEX -k "lifecycle or transports" > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt; tail -5 $O/tests.txt
fi
run 29521 scripts/hits_multigpu.py --mode delta > $O/hits_delta_epoch.json 2> $O/hits.err; echo "delta/epoch rc=$?"
run 29522 scripts/hits_multigpu.py --mode delta --sync-steps 4 > $O/hits_delta_4.json 2>> $O/hits.err; echo "delta/4 rc=$?"
run 29523 scripts/hits_multigpu.py --mode delta --scaling weak > $O/hits_delta_epoch_weak.json 2>> $O/hits.err; echo "delta/epoch weak rc=$?"
for f in $O/*.json; do echo "== $f"; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print({k:v for k,v in d.items() if k!='at'}); [print(' ',e,{k:round(x,3) for k,x in v.items()}) for e,v in d['at'].items()]"; done
tail -5 $O/hits.err
