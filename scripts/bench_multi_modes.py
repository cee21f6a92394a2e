#!/usr/bin/env python
"""Throughput of the ACCURATE multi-GPU modes of path (i) next to the seed-row exchange bench.py times:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_multi_modes.py \
        [--workload bootea_100k|bootea_15k] [--steps 40]

 exact : the step's batch of B positives sharded over the ranks (strong scaling: the global batch is what one GPU would
         draw), gradients / flags / loss all-reduced (NCCL), identical optimiser step everywhere.
 delta : every rank trains its head-owned triple shard with the full batch B (weak scaling, as the seed-row mode); once
         per global epoch the replicas are combined, x ← x_ref + Σ_g (x_g − x_ref) for weights and Adagrad slots
         (all-reduce of the tables).
Same timing rules as bench.py: CUDA events around every step, 512 MiB L2 flush between steps, max over ranks."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="bootea_100k")
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    import torch
    import torch.distributed as dist
    from openea_b200 import parallel as par
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = bench.WORKLOADS[args.workload]
    B, k = cfg["batch"], cfg["k"]
    flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    out = {"workload": cfg["name"], "n_gpus": world, "steps": args.steps}
    for mode in ("exact", "delta"):
        W = bench.build_workload(args.workload, rank, dev, world if mode == "delta" else 1)     # exact: unsharded triples
        tr, kg1, kg2, tset = W["trainer"], W["kg1"], W["kg2"], W["tset"]
        full_steps = max(1, W["steps_per_epoch"] - 1)
        exact = par.ExactReplicaStep(tr) if mode == "exact" else None
        delta = par.ReplicaDeltaSum([W["ent"], W["rel"]]) if (mode == "delta" and world > 1) else None
        epoch_steps = max(1, int(np.ceil(W["n_triples_global"] / float(world * B))))
        n_sync = [0]

        def step(i):
            if exact is not None:
                exact.step(kg1, kg2, tset, B, k, i % full_steps, 0xB007EA + i // full_steps)
            else:
                tr.step_sampled(kg1, kg2, tset, B, k, i % full_steps, 0xB007EA + 1000003 * (i // full_steps) + rank)
                if delta is not None and (i + 1) % epoch_steps == 0:
                    delta.sync(); n_sync[0] += 1
        for i in range(6):
            step(i)
        if delta is not None:
            delta.sync()
        n_sync[0] = 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(float(i))
            evs[i][0].record(); step(6 + i); evs[i][1].record()
        torch.cuda.synchronize()
        total = float(sum(a.elapsed_time(b) for a, b in evs))
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        pos_per_step = B if mode == "exact" else B * world
        out[mode] = {"ms_per_step": total / args.steps, "positives_per_s": pos_per_step * args.steps / (total * 1e-3),
                     "scaling": "strong" if mode == "exact" else "weak", "global_batch": pos_per_step,
                     "combinations_in_timed_region": n_sync[0], "epoch_steps": epoch_steps,
                     "bytes_per_combination": None if delta is None else delta.bytes_per_sync}
        tr.read_loss()
        del W, tr, kg1, kg2, tset, exact, delta
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
