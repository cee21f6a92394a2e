#!/bin/bash
# Round 2, second single-GPU call: the duo scorer on hardware (full GPU suite with it as the default + the Hits parity
# test), A/B against the octet kernel and build variants, the bench line, the approaches' epoch times, a PCIe probe.
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -rfEX -x > $O/tests.txt 2>&1; echo "pytest rc=$?" | tee -a $O/tests.txt
tail -12 $O/tests.txt
timeout 900 python scripts/ab_duo.py run 2>&1 | tee $O/ab_duo.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY' > $O/pcie.txt 2>&1
import torch, time
for mb in (0.66, 2.64, 16):
    n = int(mb * 1e6 / 4)
    h = torch.empty(n, dtype=torch.int32).pin_memory(); d = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("H2D pinned %.2f MB: %.1f us  %.1f GB/s" % (mb, ms * 1e3, mb / ms))
PY
cat $O/pcie.txt
timeout 900 python scripts/bench_approaches.py > $O/approaches.json 2> $O/approaches.err; echo "approaches rc=$?"
ls -la $O
