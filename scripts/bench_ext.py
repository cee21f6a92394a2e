"""Device-timed launches of the score-family scorer (oea_model_score_fed) at the BootEA 15K / 100K batch shapes:
kernel time (CUDA events on the launching stream, L2 flushed between launches), algorithmic GB/s against the measured
HBM copy peak.  Algorithmic bytes per scored triple = (rows read + gradient rows written) · d · 4:
TransE/DistMult 6 rows, TransH 8, TransD/SimplE 12."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openea_b200 import engine as eng  # noqa: E402

ROWS = {"TransE": 6, "TransH": 8, "TransD": 12, "DistMult": 6, "SimplE": 12}
SLOTS = {"TransE": "er", "TransH": "er-R", "TransD": "erER", "DistMult": "er", "SimplE": "erER"}


def main():
    peak = 6577.7
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    res = []
    for shape, n_ent, n_rel, B, k in (("15K", 30000, 450, 5000, 10), ("100K", 200000, 600, 20000, 10)):
        rng = np.random.default_rng(0)
        d = 100
        for model in ("TransE", "TransH", "TransD", "DistMult", "SimplE"):
            tabs = []
            for c in SLOTS[model]:
                if c == "-":
                    tabs.append(None)
                    continue
                rows = n_ent if c in "eE" else n_rel
                tabs.append(eng.EmbeddingTable((rng.standard_normal((rows, d)) / 10).astype(np.float32), True))
            loss = eng.loss_cfg("logistic" if model in ("DistMult", "SimplE") else "limited", "L2", 0.01, 2.0, 0.2)
            tr = eng.ModelTrainer(model, tabs, loss, 0.01)
            pos = torch.from_numpy(np.stack([rng.integers(0, n_ent, B), rng.integers(0, n_rel, B),
                                             rng.integers(0, n_ent, B)]).astype(np.int32)).cuda()
            neg = pos.repeat_interleave(k, dim=1).clone()
            neg[0, ::2] = torch.from_numpy(rng.integers(0, n_ent, neg.shape[1] // 2 + neg.shape[1] % 2).astype(np.int32)).cuda()[:neg[0, ::2].numel()]
            neg[2, 1::2] = torch.from_numpy(rng.integers(0, n_ent, neg.shape[1] // 2).astype(np.int32)).cuda()
            times = []
            for it in range(13):
                flush.fill_(it & 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                tr.score_fed(pos, neg)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    times.append(e0.elapsed_time(e1))
                for t in tr.live:
                    t.grad.zero_(); t.touched.zero_()
            ms = float(np.median(times))
            nbytes = ROWS[model] * d * 4 * B * (1 + k)
            res.append(dict(shape=shape, model=model, ms=ms, scored_triples_per_s=B * (1 + k) / ms * 1e3,
                            algorithmic_gbs=nbytes / ms / 1e6, frac_of_hbm_copy_peak=nbytes / ms / 1e6 / peak))
            print(json.dumps(res[-1]), flush=True)
    return res


if __name__ == "__main__":
    main()
