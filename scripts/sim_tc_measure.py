#!/usr/bin/env python
"""Tensor-core similarity (oea_sim_matrix_tc) vs the FP32 kernel at the bench's CSLS size: time of the stored-matrix pass,
of the whole materialised CSLS evaluation, value error, and arg-max / rank disagreements.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openea_b200 import finding as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda")
g = torch.Generator().manual_seed(99)
e2 = torch.randn(n, d, generator=g)
e1 = e2 + 0.5 * torch.randn(n, d, generator=g)
d1, _ = F.to_device_rows(e1.to(dev), False)
d2, _ = F.to_device_rows(e2.to(dev), False)
ld = (n + 3) // 4 * 4


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1_.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1_))
    return float(np.median(ts))


out = {"n": n, "d": d}
buf = torch.empty(n, ld, dtype=torch.float32, device=dev)
out["store_fp32_ms"] = timed(lambda: F.sim_matrix(d1, d2, d, "inner", out=buf))
ref_rows = buf[:2048].clone()
out["store_tc_ms"] = timed(lambda: F.sim_matrix_tc(d1, d2, d, "inner", out=buf))
out["max_abs_diff_first_2048_rows"] = float((buf[:2048, :n] - ref_rows[:, :n]).abs().max())
out["max_abs_value"] = float(ref_rows.abs().max())
flops = 2.0 * n * n * d
out["store_fp32_tflops"] = flops / out["store_fp32_ms"] / 1e9
out["store_tc_tflops_fp32_equivalent"] = flops / out["store_tc_ms"] / 1e9
del buf, ref_rows
torch.cuda.empty_cache()
res = {}
for flag in ("0", "1"):
    os.environ["OEA_SIM_TC"] = flag
    ms = timed(lambda: F.eval_alignment(d1, d2, [1, 5, 10, 50], "inner", False, 10), reps=3)
    top1, rk, hits, mr, mrr = F.eval_alignment(d1, d2, [1, 5, 10, 50], "inner", False, 10)
    res[flag] = (ms, top1.clone(), rk.clone(), hits, mrr)
out["csls_eval_fp32_ms"], out["csls_eval_tc_ms"] = res["0"][0], res["1"][0]
out["csls_pairs_per_s_fp32"] = float(n) * n / res["0"][0] * 1e3
out["csls_pairs_per_s_tc"] = float(n) * n / res["1"][0] * 1e3
out["argmax_disagreements"] = int((res["0"][1] != res["1"][1]).sum())
out["rank_disagreements"] = int((res["0"][2] != res["1"][2]).sum())
out["hits_fp32"], out["hits_tc"] = res["0"][3], res["1"][3]
print(json.dumps(out))
