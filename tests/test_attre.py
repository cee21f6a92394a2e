"""CPU: AttrE's host layer (literal cleaning, character vocabulary, value → character-id table, n-gram weights, fixed-size
attribute batches) and its lifecycle on the CPU warp emulator.  The three losses themselves are pinned to the
reference's own graph in tests/test_reference_graph_goldens.py."""
import re
import types

import numpy as np
import torch

from openea_b200.approaches import attre as at
from tests.test_emu_triple_core import cpu_engine      # noqa: F401  (fixture: the engine over the emulated library)


def test_literal_cleaning_and_character_table():
    assert at.clean_literal('Ada_Lovelace (mathematician)') == 'Ada Lovelace'
    assert at.clean_literal('1815-12-10"^^xsd:date') == '1815 12 10'
    assert at.clean_literal('a.b,c') == 'abc'
    kg = lambda triples: types.SimpleNamespace(local_attribute_triples_list=triples)
    kgs = types.SimpleNamespace(kg1=kg([(0, 0, "abc"), (1, 1, "ab")]), kg2=kg([(5, 2, "b_c"), (6, 0, "abc")]))
    t1, t2, chars, n_rows = at.formatting_attr_triples(kgs, 4)
    assert t1.tolist() == [[0, 0, 0], [1, 1, 1]] and t2.tolist() == [[5, 2, 2], [6, 0, 3]]      # one value id per triple
    assert chars.shape == (4, 4) and n_rows == 5                       # a, b, c, ' ' + the padding row 0
    assert chars[0].tolist() == chars[3].tolist() and chars[0][3] == 0 and chars[1][2] == 0          # same literal, padding
    assert len({chars[0][0], chars[0][1], chars[0][2]}) == 3 and chars[2][1] != 0                   # 'b c': the blank is a character
    np.testing.assert_allclose(at.ngram_weights(5), [1 / 5 + 1 / 4 + 1 / 3 + 1 / 2 + 1, 1 / 5 + 1 / 4 + 1 / 3 + 1 / 2,
                                                     1 / 5 + 1 / 4 + 1 / 3, 1 / 5 + 1 / 4, 1 / 5], rtol=1e-6)


def test_attre_lifecycle_on_the_emulator(cpu_engine, monkeypatch, tmp_path, capsys):
    import ctypes as C
    from openea_b200 import presets
    from openea_b200.approaches import AttrE
    from openea_b200.modules.base import initializers, losses
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    monkeypatch.setattr(initializers, "_make", lambda values, norm, optimizer=None: cpu_engine.EmbeddingTable(
        values, bool(norm), optimizer or "Adagrad", "cpu"))
    monkeypatch.setattr(at, "load_session", lambda: None)
    monkeypatch.setattr(losses, "_stream_ptr", lambda: C.c_void_p(0))
    folder = write_dataset(str(tmp_path) + "/micro/", "micro")
    args = presets.attre("15K")
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.batch_size, args.max_epoch, args.start_valid, args.dim, args.cuda_graph = 64, 2, 1000, 16, False
    kgs = read_kgs_from_folder(folder, args.dataset_division, "sharing", args.ordered)
    model = AttrE()
    model.set_args(args)
    model.set_kgs(kgs)
    model.init()
    pos, neg = model._attribute_batch(0)
    assert pos.shape == neg.shape == (3, 64) and torch.equal(pos[1:], neg[1:]) and not bool((pos[0] == neg[0]).any())
    last = model._attribute_batch((len(model.attribute_triples_list1) + len(model.attribute_triples_list2)) // 64)
    assert last[0].shape == (3, 64)                                     # wrap-around keeps the batch size fixed
    before = [t.weight.clone() for t in (model.ent_embeds, model.ent_embeds_ce, model.attr_embeds, model.char_embeds)]
    model.run()
    out = capsys.readouterr().out
    for tag in (r"avg\. triple loss: ([0-9.]+)", r"CE, avg\. triple loss: ([0-9.]+)", r"joint learning loss: ([0-9.]+)"):
        vals = [float(x) for x in re.findall(tag, out)]
        assert len(vals) >= 2 and all(np.isfinite(vals)), tag
    for b, t in zip(before, (model.ent_embeds, model.ent_embeds_ce, model.attr_embeds, model.char_embeds)):
        assert not torch.equal(b, t.weight) and torch.isfinite(t.weight).all()
