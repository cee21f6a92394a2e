"""CPU: IMUSE (openea_b200/approaches/imuse.py) — the string matcher against known answers and a direct statement of
the reference's nested scan, and the approach's lifecycle on the CPU warp emulator (sampled TransE steps with SGD + the
pair-distance align loss)."""
import re

import numpy as np
import pytest
import torch

from openea_b200.approaches import imuse as im
from tests.test_emu_triple_core import cpu_engine      # noqa: F401  (fixture: the engine over the emulated library)


def lcs_dp(a, b):
    t = [[0] * (len(b) + 1) for _ in range(len(a) + 1)]
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            t[i + 1][j + 1] = t[i][j] + 1 if x == y else max(t[i][j + 1], t[i + 1][j])
    return t[len(a)][len(b)]


def test_levenshtein_ratio_known_answers_and_lcs_against_dp():
    assert im.levenshtein_ratio("Hello world!", "Holly grail!") == pytest.approx(0.583333, abs=1e-6)   # python-Levenshtein's doc example
    assert im.levenshtein_ratio("kitten", "sitting") == pytest.approx(8 / 13)
    assert im.levenshtein_ratio("", "") == 1.0 and im.levenshtein_ratio("abc", "") == 0.0
    assert im.levenshtein_ratio("same", "same") == 1.0
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = "".join(rng.choice(list("abcé "), rng.integers(0, 90)))
        b = "".join(rng.choice(list("abcé "), rng.integers(0, 90)))
        assert im.lcs_length(a, b) == lcs_dp(a, b)


def test_greedy_partner_scan_equals_the_nested_loops():
    rng = np.random.default_rng(1)
    for _ in range(30):
        n1, n2 = rng.integers(1, 9), rng.integers(1, 9)
        sim = np.round(rng.random((n1, n2)), 1)               # ties on purpose
        th = 0.4
        # imuse.py:47-65 with one worker: e2 loop, strict improvement, pair emitted inside the loop if not yet taken
        pairs, taken = set(), set()
        for i in range(n1):
            target, best = None, th
            for j in range(n2):
                if sim[i, j] > best:
                    target, best = j, sim[i, j]
                if target is not None and target not in taken:
                    pairs.add((i, target))
                    taken.add(target)
        want = {}
        for i, j in sorted(pairs, key=lambda p: (p[0], sim[p[0], p[1]])):    # several partners: keep the last (best) one
            want[i] = j
        cand = {i: [(j, sim[i, j]) for j in range(n2)] for i in range(n1)}
        assert im.greedy_partner_scan(cand, th) == want


class _KG:
    def __init__(self, attr_ids, triples):
        self.attributes_id_dict = attr_ids
        self.attributes_set = set(attr_ids.values())
        self.attribute_triples_set = set(triples)


def test_interactive_model_on_a_hand_made_pair_of_kgs():
    kg1 = _KG({"http://a/name": 0, "http://a/birthDate": 1, "http://a/zzz": 2},
              [(10, 0, "Ada Lovelace"), (11, 0, "Alan Turing"), (12, 0, "Grace Hopper"), (10, 1, "1815-12-10"),
               (11, 1, "1912-06-23"), (12, 2, "x")])
    kg2 = _KG({"http://b/names": 0, "http://b/birth_date": 1, "http://b/qqq": 2},
              [(20, 0, "Ada Lovelace"), (21, 0, "Alan M. Turing"), (22, 0, "Someone Else"), (20, 1, "1815-12-10"),
               (21, 1, "1912-06-23"), (22, 2, "y")])

    class KGs:
        pass
    kgs = KGs()
    kgs.kg1, kgs.kg2 = kg1, kg2
    attr_pairs = im.get_aligned_attr_pair_by_name_similarity(kgs, 0.6)
    assert attr_pairs == {(0, 0), (1, 1)}                    # name~names, birthDate~birth_date; zzz / qqq stay unpaired

    class Args:
        sim_thresholds_ent, sim_thresholds_attr, interactive_model_iter_num = 0.6, 0.6, 1
    pairs = im.interactive_model(kgs, Args)
    assert pairs == {(10, 20), (11, 21)}                     # Grace Hopper has no counterpart above the threshold
    Args.interactive_model_iter_num = 2
    assert im.interactive_model(kgs, Args) == {(10, 20), (11, 21)}


def test_imuse_lifecycle_on_the_emulator(cpu_engine, monkeypatch, tmp_path, capsys):
    from openea_b200 import presets
    from openea_b200.approaches import IMUSE
    from openea_b200.modules.base import initializers
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    monkeypatch.setattr(initializers, "_make", lambda values, norm, optimizer=None: cpu_engine.EmbeddingTable(
        values, bool(norm), optimizer or "Adagrad", "cpu"))
    monkeypatch.setattr(im, "load_session", lambda: None)
    folder = write_dataset(str(tmp_path) + "/micro/", "micro")
    args = presets.imuse("15K")
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 64, 2, 1000, 16
    args.cuda_graph = False                 # CUDA graphs need a device; the epochs run step by step on the emulator
    kgs = read_kgs_from_folder(folder, args.dataset_division, "sharing", args.ordered)
    model = IMUSE()
    model.set_args(args)
    model.set_kgs(kgs)
    model.init()
    if not model.aligned_ent_pair_set:      # the synthetic literals may not be similar enough: train on two seed pairs
        model.aligned_ent_pair_set = set(kgs.train_links[:2])
    before = model.ent_embeds.weight.clone()
    model.run()
    out = capsys.readouterr().out
    triple = [float(x) for x in re.findall(r"avg\. triple loss: ([0-9.]+)", out)]
    align = [float(x) for x in re.findall(r"align learning loss: ([0-9.]+)", out)]
    assert len(triple) == 2 and len(align) == 2 and all(np.isfinite(triple + align))
    assert "aligned_attr_pair_set:" in out and "align_entity_by_attributes..." in out
    assert not torch.equal(before, model.ent_embeds.weight) and torch.isfinite(model.ent_embeds.weight).all()
