"""Weighted margin scorer (oea_triple_score_margin_weighted: triples, and relation paths with the relation table as both
tables) and the pair-distance loss at the IPTransE / IMUSE 15K and 100K batch shapes: CUDA events, L2 flushed between
launches, against the measured copy peak's byte count (24·d B gathered + 24·d B reduced per pair; 8·d + 8·d for pairs)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openea_b200 import engine as eng  # noqa: E402

flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, after):
    times = []
    for it in range(13):
        flush.fill_(it & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
        after()
    return 1e3 * float(np.median(times))


for shape, n_ent, n_rel, n in (("15K", 30000, 450, 5000), ("100K", 200000, 600, 20000)):
    rng = np.random.default_rng(0)
    d = 100
    ent = (rng.standard_normal((n_ent, d)) / 10).astype(np.float32)
    rel = (rng.standard_normal((n_rel, d)) / 10).astype(np.float32)
    tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), eng.loss_cfg("margin-based", "L2", margin=1.5), 0.01)
    triple = lambda hi: torch.from_numpy(np.stack([rng.integers(0, hi, n), rng.integers(0, n_rel, n), rng.integers(0, hi, n)]).astype(np.int32)).cuda()
    pos, neg, ppos, pneg = triple(n_ent), triple(n_ent), triple(n_rel), triple(n_rel)
    w = torch.rand(n, device="cuda") + 0.5
    a, b = torch.randint(0, n_ent, (n,), dtype=torch.int32, device="cuda"), torch.randint(0, n_ent, (n,), dtype=torch.int32, device="cuda")
    row = {"shape": shape, "pairs": n, "dim": d}
    row["weighted_triples_us"] = timed(lambda: tr.score_margin_weighted(pos, neg, w), tr.apply)
    row["weighted_paths_us"] = timed(lambda: tr.score_margin_weighted(ppos, pneg, w, reciprocal=True, scale=0.1, paths=True), tr.apply)
    row["plain_margin_us"] = timed(lambda: tr.score_fed(pos, neg), tr.apply)
    row["pair_distance_us"] = timed(lambda: tr.score_pairs(a, b, w), tr.apply)
    row["weighted_triples_GBps"] = 48.0 * d * n / (row["weighted_triples_us"] * 1e-6) / 1e9
    row["pair_distance_GBps"] = 16.0 * d * n / (row["pair_distance_us"] * 1e-6) / 1e9
    print(json.dumps(row), flush=True)
