"""CPU: AliNet's batch generator as tensor ops (approaches/alinet.py: NeighborTable, sample_negative_links,
generate_input_batch, generate_rel_batch on the relation index) — the invariants of alinet.py:983-1017 on CPU tensors
(the product runs the same code on CUDA tensors)."""
from types import SimpleNamespace

import numpy as np
import torch

from openea_b200.approaches import alinet


def _setup(seed=0, n=200, s=40, num=12):
    rng = np.random.default_rng(seed)
    sup1, sup2 = rng.permutation(100)[:s].tolist(), (100 + rng.permutation(100)[:s]).tolist()
    ref1 = [e for e in range(100) if e not in set(sup1)][:50]
    ref2 = [e for e in range(100, 200) if e not in set(sup2)][:50]
    ents1, ents2 = sup1 + ref1, sup2 + ref2
    nb1 = alinet.NeighborTable(ents1, torch.as_tensor(np.stack([rng.choice(ents2, num, replace=False) for _ in ents1])), n)
    nb2 = alinet.NeighborTable(ents2, torch.as_tensor(np.stack([rng.choice(ents1, num, replace=False) for _ in ents2])), n)
    return rng, n, sup1, sup2, ref1, ref2, nb1, nb2


def test_truncated_negatives_come_from_the_neighbour_lists_without_replacement():
    rng, n, sup1, sup2, ref1, ref2, nb1, nb2 = _setup()
    k = 5
    pos = torch.as_tensor(np.stack([sup1[:30], sup2[:30]], 1))
    known = {(a, b) for a, b in zip(sup1, sup2)} | {(sup1[0], nb1[sup1[0]][0])}        # one candidate is a known link
    keys = torch.as_tensor(np.sort(np.array([a * n + b for a, b in known], dtype=np.int64)))
    gen = torch.Generator().manual_seed(1)
    neg = alinet.sample_negative_links(pos, nb1, nb2, None, k, keys, n, gen)
    pairs = [tuple(x) for x in neg.tolist()]
    assert pairs == sorted(set(pairs))                                    # a sorted set, like np.array(sorted(neg))
    assert not set(pairs) & known
    by_left, by_right = {}, {}
    for a, b in pairs:
        by_left.setdefault(a, set()).add(b)
        by_right.setdefault(b, set()).add(a)
    for e1, e2 in pos.tolist():
        got1 = by_left[e1] & set(nb1[e1])                                 # (e1, c): c from e1's list
        got2 = by_right[e2] & set(nb2[e2])                                # (c, e2): c from e2's list
        assert len(got1) >= k - 1 and len(got2) >= k - 1                  # k distinct draws (one may be a known link)
    allowed = {(e1, c) for e1, _ in pos.tolist() for c in nb1[e1]} | {(c, e2) for _, e2 in pos.tolist() for c in nb2[e2]}
    assert set(pairs) <= allowed
    assert len(pairs) >= 2 * k * 30 - 35                                  # few collisions between the two sides
    again = alinet.sample_negative_links(pos, nb1, nb2, None, k, keys, n, torch.Generator().manual_seed(1))
    assert torch.equal(again, neg)
    other = alinet.sample_negative_links(pos, nb1, nb2, None, k, keys, n, torch.Generator().manual_seed(2))
    assert not torch.equal(other, neg)


def test_uniform_negatives_zip_distinct_pool_draws():
    rng, n, sup1, sup2, ref1, ref2, _, _ = _setup(3)
    pools = (torch.as_tensor(sup1 + ref1), torch.as_tensor(sup2 + ref2))
    pos = torch.as_tensor(np.stack([sup1[:25], sup2[:25]], 1))
    neg = alinet.sample_negative_links(pos, None, None, pools, 4, torch.zeros(0, dtype=torch.long), n,
                                       torch.Generator().manual_seed(0))
    assert neg.shape[1] == 2 and 90 <= neg.shape[0] <= 100                # 4 rounds × 25 zipped pairs, minus collisions
    assert set(neg[:, 0].tolist()) <= set(sup1 + ref1) and set(neg[:, 1].tolist()) <= set(sup2 + ref2)


def test_generate_input_batch_and_rel_batch_through_the_model_object():
    rng, n, sup1, sup2, ref1, ref2, nb1, nb2 = _setup(5)
    m = object.__new__(alinet.AliNet)
    m.args = SimpleNamespace(neg_triple_num=3, seed=0)
    m.session = SimpleNamespace(device=torch.device("cpu"))
    m.kgs = SimpleNamespace(entities_num=n)
    m.sup_ent1, m.sup_ent2, m.ref_ent1, m.ref_ent2 = sup1, sup2, ref1, ref2
    m.sup_links = np.stack([np.array(sup1), np.array(sup2)], 1)
    m.sup_links_set, m.new_sup_links_set = set(zip(sup1, sup2)), set()
    np.random.seed(0)
    pos, neg = m.generate_input_batch(16, nb1, nb2)
    assert pos.shape == (16, 2) and {tuple(x) for x in pos.tolist()} <= m.sup_links_set
    assert neg.dtype == torch.long and not {tuple(x) for x in neg.tolist()} & m.sup_links_set
    pos_u, neg_u = m.generate_input_batch(16, None, None)
    assert pos_u.shape == (16, 2) and neg_u.shape[0] > 30
    m.new_sup_links_set = {tuple(neg[0].tolist())}                          # the augmentation labels a pair: excluded next time
    pos2, neg2 = m.generate_input_batch(1000, nb1, nb2)                      # batch_size is capped at the seed count
    assert pos2.shape[0] == len(sup1) and tuple(neg[0].tolist()) not in {tuple(x) for x in neg2.tolist()}
    # relation batches: rel_win_size pairs per relation, every pair one of that relation's (h, t)
    tri = np.unique(np.stack([rng.integers(0, n, 150), rng.integers(0, 6, 150), rng.integers(0, n, 150)], 1), axis=0)
    from openea_b200.approaches import alinet_graph as ag
    m.rel_index, m.rel_win_size = ag.relation_index(tri), 7
    hs, rs, ts = m.generate_rel_batch()
    assert len(hs) == len(rs) == len(ts) == 7 * len(m.rel_index[0])
    known = {tuple(x) for x in tri.tolist()}
    assert all((int(h), int(r), int(t)) in known for h, r, t in zip(hs, rs, ts))
    assert np.array_equal(np.asarray(rs).reshape(-1, 7)[:, 0], m.rel_index[0])
