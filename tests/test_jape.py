"""CPU: JAPE (openea_b200/approaches/jape.py).  The structure loss Σ s⁺ − α·Σ s⁻ is pinned to the reference's own graph
in tests/test_reference_graph_goldens.py; here: the candidate sampler / NCE pieces of the attribute auxiliary against
their closed forms, and the lifecycle on the CPU warp emulator."""
import math
import re

import numpy as np
import pytest
import torch

from openea_b200.approaches import jape as jp
from tests.test_emu_triple_core import cpu_engine      # noqa: F401  (fixture: the engine over the emulated library)


def test_log_uniform_sampler_and_nce_loss_closed_forms():
    rng = np.random.default_rng(0)
    counts = np.zeros(50)
    for _ in range(4000):
        ids, tries = jp.log_uniform_unique(50, 5, rng)
        assert len(set(ids.tolist())) == 5 and tries >= 5 and ids.min() >= 0 and ids.max() < 50
        counts[ids[0]] += 1                                    # the first draw follows P(c) = log((c+2)/(c+1)) / log(51)
    p = np.log((np.arange(50) + 2.0) / (np.arange(50) + 1.0)) / math.log(51.0)
    assert np.abs(counts / 4000 - p).max() < 0.03
    # nce_loss against the formula written out for one example
    torch.manual_seed(0)
    w, b, x = torch.randn(7, 4), torch.randn(7), torch.randn(1, 4)
    label, sampled, tries = torch.tensor([2]), torch.tensor([0, 5]), 3
    q = lambda c: -math.expm1(tries * math.log1p(-(math.log((c + 2) / (c + 1)) / math.log(8.0))))
    sp = lambda z: math.log1p(math.exp(z))
    t = float(x[0] @ w[2] + b[2]) - math.log(q(2))
    want = sp(-t) + sum(sp(float(x[0] @ w[c] + b[c]) - math.log(q(c))) for c in (0, 5))
    assert float(jp.nce_loss(w, b, label, x, sampled, tries, 7)[0]) == pytest.approx(want, rel=1e-5)


def test_jape_lifecycle_on_the_emulator(cpu_engine, monkeypatch, tmp_path, capsys):
    from openea_b200 import presets
    from openea_b200.approaches import JAPE
    from openea_b200.modules.base import initializers
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    monkeypatch.setattr(initializers, "_make", lambda values, norm, optimizer=None: cpu_engine.EmbeddingTable(
        values, bool(norm), optimizer or "Adagrad", "cpu"))
    monkeypatch.setattr(jp, "load_session", lambda: None)
    folder = write_dataset(str(tmp_path) + "/micro/", "micro")
    args = presets.jape("15K")
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.batch_size, args.max_epoch, args.start_valid, args.dim, args.cuda_graph = 64, 2, 1000, 16, False
    args.attr_max_epoch, args.sub_mat_size, args.attr_sim_mat_threshold = 2, 4, 0.5
    kgs = read_kgs_from_folder(folder, args.dataset_division, "sharing", args.ordered)
    model = JAPE()
    model.set_args(args)
    model.set_kgs(kgs)
    model.init()
    before = model.ent_embeds.weight.clone()
    model.run()
    out = capsys.readouterr().out
    triple = [float(x) for x in re.findall(r"avg\. triple loss: (-?[0-9.]+)", out)]
    sim = [float(x) for x in re.findall(r"sim loss: ([0-9.]+)", out)]
    assert len(triple) == 2 and len(sim) == 2 and all(np.isfinite(triple + sim))
    assert "Training attribute embeddings:" in out and "Joint training:" in out
    assert model.attr_sim_mat.shape == (len(model.ref_entities1), len(model.ref_entities2))
    assert not torch.equal(before, model.ent_embeds.weight) and torch.isfinite(model.ent_embeds.weight).all()
