#!/usr/bin/env python
"""Hits@k of path (i) trained on N GPUs vs one GPU, same synthetic 15K KG, same epochs.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/hits_multigpu.py \
        [--mode stale|exact] [--epochs 100] [--seeds 11 12 13] [--scaling strong|weak]

 stale  (north-star design, openea_b200/parallel.py): triples sharded by head-row owner, replicated tables, the owners'
        seed-pair rows exchanged once per global epoch over NVLink peer memory (oea_seed_push / _pull), the final table
        assembled from the owners' rows.  Statistical parity only.
 exact  (SURVEY 8e exact-parity mode, parallel.ExactReplicaStep): batch sharded, gradients all-reduced, identical updates.
 strong scaling: per-rank batch B/N, the same number of optimiser steps per epoch as one GPU (the comparison that makes
        sense for accuracy); weak: per-rank batch B (what bench.py times), N× fewer steps per epoch.
Rank 0 prints one JSON line with mean / sd of Hits@1, Hits@10, MRR (plain and CSLS) at the evaluated epochs.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="stale", choices=["stale", "exact", "delta"])
    ap.add_argument("--sync-steps", type=int, default=0, help="delta mode: steps between delta sums (0 = once per global epoch)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--seeds", type=int, nargs="+", default=[11, 12, 13])
    ap.add_argument("--eval-at", type=int, nargs="+", default=[20, 50, 100])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from openea_b200 import engine as eng, finding as F, parallel as par
    from openea_b200.synth import synth_id_arrays
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    arr = synth_id_arrays("15K")
    dim, B, k, eps, freq = 100, 5000, 10, 0.9, 10
    T_global = len(arr["triples1"]) + len(arr["triples2"])
    t1, t2 = arr["triples1"], arr["triples2"]
    if args.mode in ("stale", "delta") and world > 1:
        t1, t2 = par.shard_triples(t1, rank, world), par.shard_triples(t2, rank, world)
    b_rank = B if (args.scaling == "weak" or args.mode == "exact") else max(1, B // world)
    b_global = b_rank * (world if args.mode in ("stale", "delta") else 1)
    steps = -(-T_global // b_global)
    links = arr["test_links"]
    results = []
    for seed in args.seeds:
        g = torch.Generator().manual_seed(seed)
        std = dim ** -0.5

        def trunc(n):
            x = torch.randn(n, dim, generator=g)
            bad = x.abs() > 2
            while bad.any():
                x[bad] = torch.randn(int(bad.sum()), generator=g)
                bad = x.abs() > 2
            return x * std
        ent, rel = eng.EmbeddingTable(trunc(arr["n_ent"]), True, "Adagrad", dev), eng.EmbeddingTable(trunc(arr["n_rel"]), True, "Adagrad", dev)
        trn = eng.TripleTrainer(ent, rel, eng.loss_cfg("limited", "L2", 0.01, 2.0, 0.2), 0.01)
        kg1, kg2 = eng.DeviceKG(t1, arr["entities1"], arr["n_ent"], dev), eng.DeviceKG(t2, arr["entities2"], arr["n_ent"], dev)
        # membership filter over ALL known triples (a corrupted triple that exists on another shard is still a true triple)
        full1 = torch.as_tensor(arr["triples1"], device=dev)
        full2 = torch.as_tensor(arr["triples2"], device=dev)
        tset = eng.DeviceTripleSet([full1, full2], arr["n_ent"], arr["n_rel"], dev)
        xchg = exact = None
        if world > 1 and args.mode == "stale":
            seeds_rows = np.concatenate([arr["train_links"][:, 0], arr["train_links"][:, 1]])
            xchg = par.SeedRowSync(ent.weight, seeds_rows, rank, world)
        if args.mode == "exact":
            exact = par.ExactReplicaStep(trn)
        delta = par.ReplicaDeltaSum([ent, rel]) if (args.mode == "delta" and world > 1) else None
        gstep = 0
        curve = {}
        for epoch in range(1, args.epochs + 1):
            for step in range(steps):
                if exact is not None:
                    exact.step(kg1, kg2, tset, b_rank, k, step, 7919 * seed + epoch)
                else:
                    # stale mode: every rank draws from ITS shard (different permutation seed per rank)
                    trn.step_sampled(kg1, kg2, tset, b_rank, k, step, 7919 * seed + epoch + 104729 * rank)
                    gstep += 1
                    if delta is not None and args.sync_steps > 0 and gstep % args.sync_steps == 0:
                        delta.sync()
            trn.read_loss()
            if xchg is not None:
                xchg.sync()
            if delta is not None and (args.sync_steps == 0 or epoch in args.eval_at or epoch % freq == 0):
                delta.sync()
            need_full = epoch in args.eval_at or epoch % freq == 0
            if need_full and world > 1 and args.mode == "stale":
                par.assemble_owned_rows(ent.weight, rank, world)      # every row from its owner: replicas agree here
            if epoch in args.eval_at:
                e1, e2 = ent.lookup(links[:, 0]), ent.lookup(links[:, 1])
                _, _, hits, mr, mrr = F.eval_alignment(e1, e2, [1, 5, 10, 50], "inner", False, 0)
                _, _, chits, _, cmrr = F.eval_alignment(e1, e2, [1, 5, 10, 50], "inner", False, 10)
                curve[epoch] = {"hits": list(hits), "mrr": float(mrr), "csls_hits": list(chits), "csls_mrr": float(cmrr)}
            if epoch % freq == 0:
                for kg, ents in ((kg1, arr["entities1"]), (kg2, arr["entities2"])):
                    kg.set_candidates(F.find_neighbours_device(ent.lookup(ents), ents, int((1 - eps) * len(ents))), ents)
        if xchg is not None:
            st = xchg.status()
            xchg.close()
            assert st == 0, "seed exchange timed out"
        results.append(curve)
        del ent, rel, trn, kg1, kg2, tset
        torch.cuda.empty_cache()
    if rank == 0:
        out = {"n_gpus": world, "mode": args.mode if world > 1 else "single", "scaling": args.scaling, "epochs": args.epochs,
               "seeds": args.seeds, "sync_steps": args.sync_steps, "batch_per_rank": b_rank, "steps_per_epoch": steps, "at": {}}
        for ep in args.eval_at:
            h1 = np.array([c[ep]["hits"][0] for c in results]); h10 = np.array([c[ep]["hits"][2] for c in results])
            c1 = np.array([c[ep]["csls_hits"][0] for c in results]); mrr = np.array([c[ep]["mrr"] for c in results])
            sd = lambda a: float(a.std(ddof=1)) if len(a) > 1 else 0.0
            out["at"][ep] = {"hits1_mean": float(h1.mean()), "hits1_sd": sd(h1), "hits10_mean": float(h10.mean()),
                             "csls_hits1_mean": float(c1.mean()), "csls_hits1_sd": sd(c1), "mrr_mean": float(mrr.mean())}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
