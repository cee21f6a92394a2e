"""Per-triple fed scorer (oea_triple_score_fed) vs grouped fed scorer (oea_triple_score_fed_grouped) at the BootEA 15K /
100K batch shapes, and the host-index step with and without OEA_FED_GROUPED: CUDA events, L2 flushed between launches."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openea_b200 import engine as eng  # noqa: E402

flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for shape, n_ent, n_rel, B, k in (("15K", 30000, 450, 5000, 10), ("100K", 200000, 600, 20000, 10)):
    rng = np.random.default_rng(0)
    d = 100
    ent = (rng.standard_normal((n_ent, d)) / 10).astype(np.float32)
    rel = (rng.standard_normal((n_rel, d)) / 10).astype(np.float32)
    pos = np.stack([rng.integers(0, n_ent, B), rng.integers(0, n_rel, B), rng.integers(0, n_ent, B)]).astype(np.int32)
    neg = np.repeat(pos, k, axis=1)
    side = rng.random(B * k) < 0.5
    neg[0, side] = rng.integers(0, n_ent, side.sum())
    neg[2, ~side] = rng.integers(0, n_ent, (~side).sum())
    for loss in ("limited", "logistic"):
        cfg = eng.loss_cfg(loss, "L2", 0.01, 2.0, 0.2)
        row = {"shape": shape, "loss": loss}
        for grouped in (False, True):
            tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), cfg, 0.01)
            dp, dn = torch.from_numpy(pos).cuda(), torch.from_numpy(neg).cuda()
            times = []
            for it in range(13):
                flush.fill_(it & 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); tr.score_fed(dp, dn, grouped=grouped); e1.record(); torch.cuda.synchronize()
                if it >= 3:
                    times.append(e0.elapsed_time(e1))
                tr.apply()
            row["grouped_us" if grouped else "per_triple_us"] = 1e3 * float(np.median(times))
        tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), cfg, 0.01)
        times = []
        for it in range(13):                 # grouped scoring + row optimiser as one cooperative launch (whole step)
            flush.fill_(it & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.step_fed_grouped(dp, dn); e1.record(); torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        row["one_launch_step_us"] = 1e3 * float(np.median(times))
        tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), cfg, 0.01)
        times = []
        for it in range(13):                 # the same step as two launches (per-triple scorer, then k_rowopt_pair)
            flush.fill_(it & 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.score_fed(dp, dn); tr.apply(); e1.record(); torch.cuda.synchronize()
            if it >= 3:
                times.append(e0.elapsed_time(e1))
        row["two_launch_step_us"] = 1e3 * float(np.median(times))
        hp, hn = torch.from_numpy(pos).pin_memory(), torch.from_numpy(neg).pin_memory()
        for flag in ("0", "1"):
            os.environ["OEA_FED_GROUPED"] = flag
            tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), cfg, 0.01)
            for _ in range(3):
                tr.step_fed_host(hp, hn)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(20):
                tr.step_fed_host(hp, hn)
            row["host_step_grouped_us" if flag == "1" else "host_step_us"] = 1e6 * (time.perf_counter() - t0) / 20
        os.environ["OEA_FED_GROUPED"] = "0"
        os.environ["OEA_FED_FUSED"] = "1"
        tr = eng.TripleTrainer(eng.EmbeddingTable(ent, True), eng.EmbeddingTable(rel, True), cfg, 0.01)
        for _ in range(3):
            tr.step_fed_host(hp, hn)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            tr.step_fed_host(hp, hn)
        row["host_step_one_launch_us"] = 1e6 * (time.perf_counter() - t0) / 20
        os.environ["OEA_FED_FUSED"] = "0"
        print(json.dumps(row), flush=True)
