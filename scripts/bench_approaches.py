"""End-to-end epoch times of the five BASELINE.json approaches on the synthetic 15K-shape dataset through the
reference lifecycle (set_args / set_kgs / init / run), for DESIGN.md.  Prints one JSON line."""
import contextlib
import io
import json
import os
import re
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openea_b200 import presets  # noqa: E402
from openea_b200.approaches import AliNet, BootEA, GCN_Align, MTransE, RDGCN  # noqa: E402
from openea_b200.modules.load.kgs import read_kgs_from_folder  # noqa: E402
from openea_b200.synth import write_dataset  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "15K"
folder = tempfile.mkdtemp() + "/"
t0 = time.time()
write_dataset(folder, shape)
out = {"shape": shape, "write_dataset_s": round(time.time() - t0, 2)}


def run(name, cls, args, mode, epochs):
    args.training_data, args.output = folder, folder + "out/"
    args.max_epoch = epochs
    buf = io.StringIO()
    t = time.time()
    with contextlib.redirect_stdout(buf):
        kgs = read_kgs_from_folder(folder, args.dataset_division, mode, args.ordered)
        t_load = time.time() - t
        m = cls(); m.set_args(args); m.set_kgs(kgs)
        t = time.time(); m.init(); torch.cuda.synchronize(); t_init = time.time() - t
        t = time.time(); m.run(); torch.cuda.synchronize(); t_run = time.time() - t
        t = time.time(); m.test(save=False); torch.cuda.synchronize(); t_test = time.time() - t
    text = buf.getvalue()
    ep = [float(x) for x in re.findall(r"cost time: ([0-9.]+)s", text)]
    h1 = re.findall(r"accurate results: hits@\[1, 5, 10, 50\] = \[\s*([0-9.]+)", text)
    out[name] = {"load_s": round(t_load, 2), "init_s": round(t_init, 2), "run_s": round(t_run, 3), "epochs": epochs,
                 "median_epoch_ms": round(1e3 * float(np.median(ep[2:])) if len(ep) > 2 else -1, 3),
                 "test_s": round(t_test, 3), "hits1": float(h1[-1]) if h1 else None}
    sys.stderr.write("%s %s\n" % (name, out[name]))


a = presets.bootea(shape); a.start_valid, a.sub_epoch = 10, 10
run("BootEA", BootEA, a, "swapping", 40)
a = presets.mtranse(shape, dim=75); a.start_valid = 20
run("MTransE_d75", MTransE, a, "mapping", 40)
a = presets.gcn_align(shape); a.start_valid = 20
np.random.seed(0)
run("GCN_Align", GCN_Align, a, "mapping", 40)
a = presets.alinet(shape); a.start_valid, a.eval_freq = 10, 10
run("AliNet", AliNet, a, "mapping", 20)
a = presets.rdgcn(shape); a.start_valid = 10
run("RDGCN", RDGCN, a, "mapping", 20)
print(json.dumps(out))
