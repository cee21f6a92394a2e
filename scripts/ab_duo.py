#!/usr/bin/env python
"""A/B of the sampled scorers on the bench (device-timed, L2 flushed): octet (one positive per warp) vs duo (two), and
duo build variants (CTAs per SM / warps per CTA).   build (here, no GPU):  python scripts/ab_duo.py build
                                                      run (on the box):     python scripts/ab_duo.py run"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {"w4b5": ["OEA_DUO_WARPS=4", "OEA_DUO_MINB=5"], "w4b8": ["OEA_DUO_WARPS=4", "OEA_DUO_MINB=8"],
            "w2b12": ["OEA_DUO_WARPS=2", "OEA_DUO_MINB=12"]}
VDIR = os.path.join(ROOT, "openea_b200", "_lib", "variants")


def build():
    from openea_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        print(b.build_cuda(defines=defs, out=os.path.join(VDIR, "liboea_duo_%s.so" % name)))


def one(label, env, wl):
    e = dict(os.environ); e.update(env)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", "60", "--warmup", "8",
                          "--no-cpu-baseline", "--no-secondary"], capture_output=True, text=True, env=e, timeout=600)
    try:
        d = json.loads(res.stdout.strip().splitlines()[-1])
        r = d["roofline"]
        line = "%-12s %-10s step %.1f us  kernel %.1f us  score-alone %.1f us  value %.3e  frac %.2f  e2e %.3e (%.1f us)" % (
            wl, label, d["ms_per_step"] * 1e3, r["kernel_ms_median"] * 1e3, r["score_kernel_alone"]["ms_median"] * 1e3, d["value"],
            r["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"] * 1e3)
    except Exception as exc:
        line = "%-12s %-10s FAILED %r %s" % (wl, label, exc, res.stderr[-300:])
    print(line, flush=True)
    return line


def run():
    out = []
    for wl in ("bootea_15k", "bootea_100k"):
        out.append(one("oct", {"OEA_SCORE_DUO": "0"}, wl))
        out.append(one("duo w4b6", {"OEA_SCORE_DUO": "1"}, wl))
        for name in VARIANTS:
            lib = os.path.join(VDIR, "liboea_duo_%s.so" % name)
            if os.path.exists(lib):
                out.append(one("duo " + name, {"OEA_SCORE_DUO": "1", "OEA_LIB_PATH": lib}, wl))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_duo.txt"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    build() if sys.argv[1:] == ["build"] else run()
