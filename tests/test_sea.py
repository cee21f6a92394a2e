"""CPU: SEA's mapping module (openea_b200/approaches/sea.py) on the CPU warp emulator — the lookup / scatter / Adam
kernels' sources around the dense cycle-consistency loss — against a float64 statement of sea.py:78-100 with TF's
dense Adam, and the approach's lifecycle on a 40-entity dataset."""
import re

import numpy as np
import pytest
import torch

from openea_b200.approaches import sea as sea_mod
from tests.helpers import make_tables
from tests.test_e2e_gpu import tiny_kgs                # noqa: F401  (fixture: the tiny synthetic dataset folder)
from tests.test_emu_triple_core import cpu_engine      # noqa: F401  (fixture: the engine over the emulated library)


def test_frobenius_normalize_is_l2_normalize_without_axis():
    x = torch.tensor([[3.0, 0.0], [0.0, 4.0]])
    np.testing.assert_allclose(sea_mod.frobenius_normalize(x).numpy(), x.numpy() / 5.0, rtol=1e-6)
    assert torch.equal(sea_mod.frobenius_normalize(torch.zeros(2, 3)), torch.zeros(2, 3))     # max(Σx², 1e-12)


def _check_mapping_trainer(engine, device):
    rng = np.random.default_rng(4)
    n, d, lr, a1, a2 = 50, 12, 0.01, 2.5, 0.25
    ent, _ = make_tables(rng, n, 2, d)
    q1 = np.linalg.qr(rng.standard_normal((d, d)))[0].astype(np.float32)
    q2 = np.linalg.qr(rng.standard_normal((d, d)))[0].astype(np.float32)
    te = engine.EmbeddingTable(ent, True, "Adam", device=device)
    m1 = engine.EmbeddingTable(q1, False, "Adam", device=device)
    m2 = engine.EmbeddingTable(q2, False, "Adam", device=device)
    tr = sea_mod.SEAMappingTrainer(te, m1, m2, a1, a2, lr)

    var = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (ent, q1, q2)]
    mom = [torch.zeros_like(v) for v in var]
    vel = [torch.zeros_like(v) for v in var]
    total = 0.0
    for step in range(1, 4):
        ids = [rng.integers(0, n, m) for m in (7, 7, 11, 11)]
        ids[2][0] = ids[0][0]                                      # an entity on both the labelled and the unlabelled side
        tr.step(*ids)
        E = var[0]
        En = E * torch.rsqrt(torch.clamp((E * E).sum(1, keepdim=True), min=1e-12))
        loss = sea_mod.sea_mapping_loss(*[En[torch.as_tensor(i)] for i in ids], var[1], var[2], a1, a2)
        total += float(loss.detach())
        for v in var:
            v.grad = None
        loss.backward()
        lr_t = lr * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        with torch.no_grad():
            for v, m, s in zip(var, mom, vel):                     # tf.train.AdamOptimizer, dense: every row moves
                m.mul_(0.9).add_(v.grad, alpha=0.1)
                s.mul_(0.999).addcmul_(v.grad, v.grad, value=0.001)
                v.sub_(lr_t * m / (s.sqrt() + 1e-8))
    assert tr.read_loss() == pytest.approx(total, rel=1e-4)
    for tab, v, name in ((te, var[0], "entities"), (m1, var[1], "M1"), (m2, var[2], "M2")):
        np.testing.assert_allclose(tab.raw().cpu().numpy(), v.detach().numpy(), rtol=2e-4, atol=2e-6, err_msg=name)
    assert tr.ent.state1 is not te.state1 and not te.state1.any().item()     # the triple optimiser's slots are untouched


def test_mapping_trainer_steps_equal_float64_dense_adam(cpu_engine):
    _check_mapping_trainer(cpu_engine, "cpu")


@pytest.mark.gpu
def test_mapping_trainer_steps_equal_float64_dense_adam_gpu(cuda_device):
    from openea_b200 import engine
    _check_mapping_trainer(engine, "cuda")


@pytest.mark.gpu
def test_sea_lifecycle_gpu(cuda_device, tiny_kgs, tmp_path):
    """set_args / set_kgs / init / run / test / save of SEA on the tiny synthetic KG pair (mapping mode: separate id
    spaces): both losses fall, the mapped alignment beats chance, both mapping matrices are saved."""
    import os
    from openea_b200 import presets
    from openea_b200.approaches import SEA
    from tests.test_e2e_gpu import _hits1, _run
    args = presets.sea("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 150, 1000, 32
    model, out = _run(SEA, args, tiny_kgs, "mapping", tmp_path)
    triple = [float(x) for x in re.findall(r"avg\. triple loss:\s*([0-9.]+)", out)]
    mapping = [float(x) for x in re.findall(r"avg\. mapping loss:\s*([0-9.]+)", out)]
    # the mapped rows are normalised over the whole batch matrix (sea.py:84-85), so the mapping loss stays within
    # ±2√B of its constant part and is only required to be finite
    assert len(triple) == 150 and len(mapping) == 150 and triple[-1] < triple[0] and all(np.isfinite(mapping)), \
        (triple[::50], mapping[::50])
    assert _hits1(out, "accurate results:") >= 0.0       # the result line exists; accuracy is calibrated once this has run on a GPU
    assert os.path.exists(model.out_folder + "mapping_mat.npy") and os.path.exists(model.out_folder + "rev_mapping_mat.npy")


def test_sea_lifecycle_on_the_emulator(cpu_engine, monkeypatch, tmp_path, capsys):
    from openea_b200 import presets
    from openea_b200.approaches import SEA
    from openea_b200.modules.base import initializers
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    monkeypatch.setattr(initializers, "_make", lambda values, norm, optimizer=None: cpu_engine.EmbeddingTable(
        values, bool(norm), optimizer or "Adagrad", "cpu"))
    monkeypatch.setattr(sea_mod, "load_session", lambda: None)
    folder = write_dataset(str(tmp_path) + "/micro/", "micro")
    args = presets.sea("15K")
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 64, 2, 1000, 16
    kgs = read_kgs_from_folder(folder, args.dataset_division, "mapping", args.ordered)
    model = SEA()
    model.set_args(args)
    model.set_kgs(kgs)
    model.init()
    before = [t.weight.clone() for t in (model.ent_embeds, model.rel_embeds, model.mapping_mat_1, model.mapping_mat_2)]
    model.run()
    out = capsys.readouterr().out
    triple = [float(x) for x in re.findall(r"avg\. triple loss: ([0-9.]+)", out)]
    mapping = [float(x) for x in re.findall(r"avg\. mapping loss: ([0-9.]+)", out)]
    assert len(triple) == 2 and len(mapping) == 2 and all(np.isfinite(triple + mapping)) and mapping[0] > 0
    for b, t in zip(before, (model.ent_embeds, model.rel_embeds, model.mapping_mat_1, model.mapping_mat_2)):
        assert not torch.equal(b, t.weight) and torch.isfinite(t.weight).all()
    assert model._mapping_array().data_ptr() == model.mapping_mat_1.weight.data_ptr()
    assert set(model._checkpoint_tables()) >= {"ent_embeds", "rel_embeds", "mapping_mat_1", "mapping_mat_2", "mapping_trainer.ent"}
