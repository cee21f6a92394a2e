"""Secondary measurements for DESIGN.md (not the driver's bench): SpMM (K2) achieved bytes/s on the GCN-Align 15K
and AliNet-like 100K adjacency shapes, one GCN-Align SE-unit training step, the ε-truncated neighbour search, and
RDGCN's get_neg (L1 top-k).  Device-timed with CUDA events, L2 flushed between repetitions.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openea_b200 import finding as F, gnn  # noqa: E402
from openea_b200.engine import EmbeddingTable  # noqa: E402
from openea_b200.synth import synth_id_arrays  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(512 << 18, dtype=torch.float32, device=dev)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for i in range(reps):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {}
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for shape, d in (("15K", 100), ("100K", 400)):
    arr = synth_id_arrays(shape, swapping=False)
    n = arr["n_ent"]
    tri = np.concatenate([arr["triples1"], arr["triples2"]])
    A = gnn.DeviceCsr(gnn.preprocess_adj(gnn.weighted_adjacency(n, tri)))
    X = torch.randn(n, d, device=dev)
    Y = torch.empty(n, d, device=dev)
    ms = timed(lambda: gnn.spmm(A, X, out=Y))
    compulsory = A.nnz * 8 + (n + 1) * 4 + 2 * n * d * 4
    noreuse = A.nnz * (8 + 4 * d) + 4 * n * d
    out["spmm_" + shape] = {"n": n, "nnz": A.nnz, "d": d, "ms": ms, "long_rows": int(A.long_rows.numel()),
                            "compulsory_GBs": compulsory / ms / 1e6, "noreuse_GBs": noreuse / ms / 1e6,
                            "frac_of_hbm_peak_compulsory": compulsory / ms / 1e6 / peak,
                            "frac_of_hbm_peak_noreuse": noreuse / ms / 1e6 / peak}
    if shape == "15K":   # one SE-unit training step of GCN-Align (2 fwd + 2 bwd SpMM, loss, normalise bwd, SGD)
        from openea_b200.approaches.gcn_align import GCNAlignUnit
        tab = EmbeddingTable(torch.randn(n, d) / np.sqrt(n), True, "SGD")
        ill = arr["train_links"].astype(np.int64)
        unit = GCNAlignUnit(A, tab, None, ill, 3.0, 5, 8.0)
        t, k = len(ill), 5
        g = torch.Generator().manual_seed(0)
        negs = [torch.as_tensor(np.repeat(ill[:, 0], k), dtype=torch.int32, device=dev),
                torch.randint(0, n, (t * k,), generator=g, dtype=torch.int32).to(dev),
                torch.randint(0, n, (t * k,), generator=g, dtype=torch.int32).to(dev),
                torch.as_tensor(np.repeat(ill[:, 1], k), dtype=torch.int32, device=dev)]
        out["gcn_align_se_step_15K_ms"] = timed(lambda: unit.train_step(*negs))

# ε-truncated neighbour search (batch.py:145-165): 15 000 × 15 000, k = 1 500
e = torch.nn.functional.normalize(torch.randn(15000, 100, device=dev), dim=1)
ids = torch.arange(15000, dtype=torch.int32, device=dev)
ms = timed(lambda: F.find_neighbours_device(e, ids, 1500), reps=5)
out["neighbours_15K"] = {"n": 15000, "k": 1500, "ms": ms, "pairs_per_s": 15000.0 ** 2 / ms * 1e3}

# RDGCN get_neg (rdgcn.py:75-87): L1-nearest k of t seed rows against all E rows, d = 300
from openea_b200.approaches.rdgcn_ops import get_neg  # noqa: E402
for t, E in ((3000, 30000), (20000, 200000)):
    emb = torch.randn(E, 300, device=dev)
    ill = torch.randperm(E, device=dev)[:t].to(torch.int32)
    ms = timed(lambda: get_neg(ill, emb, 10), reps=3, warm=1)
    out["rdgcn_get_neg_%dx%d" % (t, E)] = {"ms": ms, "pairs_per_s": float(t) * E / ms * 1e3, "fp32_Tops": 3.0 * t * E * 300 / ms / 1e9}
print(json.dumps(out))
