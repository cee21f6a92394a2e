"""GPU end-to-end: the reference lifecycle (set_args / set_kgs / init / run / test / save) of the in-scope
approaches on a tiny synthetic dataset; learning must happen (Hits@1 far above chance) and the files and
stdout lines of the reference must be produced."""
import contextlib
import io
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_kgs(tmp_path_factory):
    from openea_b200.synth import write_dataset
    folder = str(tmp_path_factory.mktemp("tiny")) + "/"
    write_dataset(folder, "tiny")
    return folder


def _run(model_cls, args, folder, mode, tmp_path):
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    args.training_data = folder
    args.output = str(tmp_path) + "/out/"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kgs = read_kgs_from_folder(folder, args.dataset_division, mode, args.ordered)
        model = model_cls()
        model.set_args(args)
        model.set_kgs(kgs)
        model.init()
        model.run()
        model.test()
        model.save()
    return model, buf.getvalue()


def _hits1(out, tag):
    m = re.findall(re.escape(tag) + r".*?hits@\[1, 5, 10, 50\] = \[\s*([0-9.]+)", out)
    assert m, "no result line %r in stdout" % tag
    return float(m[-1])


def test_bootea_lifecycle(cuda_device, tiny_kgs, tmp_path):
    from openea_b200 import presets
    from openea_b200.approaches import BootEA
    args = presets.bootea("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.sub_epoch = 1000, 60, 20, 10
    args.truncated_epsilon, args.dim = 0.9, 32
    model, out = _run(BootEA, args, tiny_kgs, "swapping", tmp_path)
    assert "avg. triple loss" in out and "generating neighbors of" in out and "Training ends. Total time" in out
    assert "after mwgm" in out or "empty aligned pairs" in out
    h1 = _hits1(out, "accurate results:")
    h1_csls = _hits1(out, "accurate results with csls: csls=10,")
    assert h1 > 20.0 and h1_csls > 20.0, (h1, h1_csls)      # chance = 1/420 = 0.24 %
    for f in ("ent_embeds.npy", "rel_embeds.npy", "alignment_results_12", "kg1_ent_ids", "kg2_rel_ids", "kg1_ent_embeds_txt"):
        assert os.path.exists(model.out_folder + f), f
    ent = np.load(model.out_folder + "ent_embeds.npy")
    np.testing.assert_allclose(np.linalg.norm(ent, axis=1), 1.0, rtol=1e-4)   # save() writes the normalised table
    pairs = [l.split("\t") for l in open(model.out_folder + "alignment_results_12").read().strip().split("\n")]
    assert len(pairs) == len(model.kgs.test_links)


def test_mtranse_lifecycle(cuda_device, tiny_kgs, tmp_path):
    from openea_b200 import presets
    from openea_b200.approaches import MTransE
    args = presets.mtranse("15K", dim=75)          # BASELINE config 1 uses dim 75 (pitch 76 internally)
    args.batch_size, args.max_epoch, args.start_valid = 1000, 60, 30
    model, out = _run(MTransE, args, tiny_kgs, "mapping", tmp_path)
    assert "avg. mapping loss" in out and "quick results:" in out
    h1 = _hits1(out, "accurate results:")
    assert h1 > 5.0, h1
    assert os.path.exists(model.out_folder + "mapping_mat.npy")
    assert np.load(model.out_folder + "ent_embeds.npy").shape[1] == 75


def test_transe_margin_lifecycle(cuda_device, tiny_kgs, tmp_path):
    from openea_b200 import presets
    from openea_b200.models.trans import TransE
    args = presets.transe("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 40, 20, 32
    model, out = _run(TransE, args, tiny_kgs, "sharing", tmp_path)
    assert _hits1(out, "accurate results:") > 10.0
