"""Driver for a per-kernel launch list of one CSLS evaluation (eval_alignment, inner, k = 10) at a given n."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openea_b200 import finding as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(0)
e2 = torch.randn(n, 100, generator=g).cuda()
e1 = e2 + 0.5 * torch.randn(n, 100, device="cuda")
d1, d = F.to_device_rows(e1, False); d2, _ = F.to_device_rows(e2, False)
for _ in range(reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    top1, rk, hits, mr, mrr = F.eval_alignment(d1, d2, [1, 5, 10, 50], "inner", False, 10)
    ev1.record(); torch.cuda.synchronize()
    print("n=%d eval %.3f ms hits1 %.1f" % (n, ev0.elapsed_time(ev1), hits[0]))
