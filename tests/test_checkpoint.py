"""CPU: checkpoint / resume of the engine state (the reference has none): table round trips with every
optimiser's slots, the model-level file with the trainers' separate slot views, refusal of mismatching checkpoints."""
import random

import numpy as np
import pytest
import torch

from openea_b200 import engine as eng
from openea_b200.models.basic_model import BasicModel


def _table(rows, d, opt, seed):
    g = torch.Generator().manual_seed(seed)
    t = eng.EmbeddingTable(torch.randn(rows, d, generator=g), True, opt, device="cpu")
    for name in ("state1", "state2"):
        if getattr(t, name) is not None:
            getattr(t, name).copy_(torch.rand(getattr(t, name).shape, generator=g))
    t.adam_t = 7 if opt == "Adam" else 0
    return t


@pytest.mark.parametrize("opt", ["Adagrad", "Adam", "SGD", "Adadelta"])
def test_table_state_round_trip(opt):
    a, b = _table(30, 10, opt, 1), _table(30, 10, opt, 2)
    ptr = b.weight.data_ptr()
    b.grad.fill_(3.0)
    b.touched.fill_(1)
    b.load_state_dict(a.state_dict())
    assert torch.equal(a.weight, b.weight) and b.weight.data_ptr() == ptr        # restored in place
    for name in ("state1", "state2"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None and y is None) or torch.equal(x, y)
    assert b.adam_t == a.adam_t and not b.grad.any() and not b.touched.any()
    with pytest.raises(ValueError):
        _table(31, 10, opt, 3).load_state_dict(a.state_dict())
    with pytest.raises(ValueError):
        _table(30, 10, "SGD" if opt != "SGD" else "Adam", 3).load_state_dict(a.state_dict())


def _model(seed):
    m = BasicModel()
    m.ent_embeds, m.rel_embeds = _table(40, 12, "Adagrad", seed), _table(6, 12, "Adagrad", seed + 1)
    m.mapping_mat = _table(12, 12, "Adagrad", seed + 2)
    # a second optimiser instance over the entity table, as MTransE's mapping trainer / BootEA's alignment trainer own
    m.alignment_trainer = type("T", (), {})()
    m.alignment_trainer.ent, m.alignment_trainer.rel = m.ent_embeds.new_slots(), m.rel_embeds.new_slots()
    m.alignment_trainer.ent.state1.fill_(0.25 + seed)
    m._epoch_seed, m.flag1, m.flag2 = 1000 + seed, 0.5, 0.25
    return m


def test_model_checkpoint_restores_tables_slot_views_seeds_and_rng(tmp_path):
    a, b = _model(1), _model(5)
    random.seed(3)
    np.random.seed(4)
    path = a.save_checkpoint(str(tmp_path) + "/ckpt/checkpoint.pt", epoch=17)
    want_py, want_np = random.random(), np.random.rand()
    assert b.load_checkpoint(path) == 18 and b._start_epoch == 18
    for name, tab in a._checkpoint_tables().items():
        other = b._checkpoint_tables()[name]
        assert torch.equal(tab.weight, other.weight) and torch.equal(tab.state1, other.state1), name
    assert sorted(a._checkpoint_tables()) == ["alignment_trainer.ent", "alignment_trainer.rel", "ent_embeds",
                                              "mapping_mat", "rel_embeds"]
    assert b.alignment_trainer.ent.weight.data_ptr() == b.ent_embeds.weight.data_ptr()    # still one variable, two slot sets
    assert float(b.alignment_trainer.ent.state1[0, 0]) == 1.25 and (b._epoch_seed, b.flag1, b.flag2) == (1001, 0.5, 0.25)
    assert (random.random(), np.random.rand()) == (want_py, want_np)                   # host RNG streams continue


def test_checkpoint_of_another_model_is_refused(tmp_path):
    a = _model(1)
    path = a.save_checkpoint(str(tmp_path) + "/checkpoint.pt", epoch=2)
    b = _model(2)
    del b.mapping_mat
    with pytest.raises(ValueError):
        b.load_checkpoint(path)

    class Other(BasicModel):
        pass
    c = Other()
    c.__dict__.update(_model(3).__dict__)
    with pytest.raises(ValueError):
        c.load_checkpoint(path)


def test_bootea_checkpoint_carries_the_bootstrapped_labels(tmp_path):
    from openea_b200.approaches.bootea import BootEA
    def model(seed):
        m = BootEA()
        m.__dict__.update(_model(seed).__dict__)
        m.ref_ent1, m.ref_ent2 = list(range(10)), list(range(10, 20))
        return m
    a, b = model(1), model(2)
    a._label = torch.tensor([3, -1, 0, -1, 9, -1, -1, 2, -1, -1])
    path = a.save_checkpoint(str(tmp_path) + "/checkpoint.pt", epoch=20)
    assert b.load_checkpoint(path) == 21
    assert torch.equal(b._label, a._label) and b._ref1.tolist() == a.ref_ent1 and b._ref2.tolist() == a.ref_ent2
    c = model(3)                                         # a checkpoint from before the first bootstrapping pass
    c.load_checkpoint(model(4).save_checkpoint(str(tmp_path) + "/early.pt", epoch=10))
    assert getattr(c, "_label", None) is None


def test_mapping_epoch_draws_distinct_seed_pairs_per_step_on_the_device(capsys):
    """BasicModel.launch_mapping_training_1epo (basic_model.py:238-250): every step trains on len(train_links) // steps
    DISTINCT seed pairs drawn independently of the other steps; the loss is read once per epoch."""
    from types import SimpleNamespace
    rng = np.random.default_rng(0)
    links = [(int(a), int(b)) for a, b in zip(rng.permutation(500)[:97], 500 + rng.permutation(500)[:97])]
    seen = []

    class Recorder:
        loss_dev = torch.zeros(1, dtype=torch.float64)

        def step(self, e1, e2, read_loss=True):
            assert read_loss is False and e1.dtype == torch.int32
            seen.append(list(zip(e1.tolist(), e2.tolist())))
            self.loss_dev += len(e1)

        def read_loss(self):
            v = float(self.loss_dev.item())
            self.loss_dev.zero_()
            return v
    m = BasicModel()
    m.kgs = SimpleNamespace(train_links=links)
    m.ent_embeds = SimpleNamespace(device=torch.device("cpu"))
    m.mapping_trainer = Recorder()
    loss = m.launch_mapping_training_1epo(3, 8)
    assert len(seen) == 8 and all(len(b) == 97 // 8 and len(set(b)) == len(b) and set(b) <= set(links) for b in seen)
    assert len({frozenset(b) for b in seen}) == 8                      # independent draws: the steps differ
    assert loss == 1.0 and "avg. mapping loss: 1.0000" in capsys.readouterr().out
    seen.clear()
    assert m.launch_mapping_training_1epo(4, 200) == 0.0 and not seen    # more steps than links: empty batches, as [] would be
