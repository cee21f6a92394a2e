"""Synthetic KG pairs of the OpenEA 15K / 100K shapes: the real datasets are not available
offline, so benchmarks and end-to-end parity runs use seeded graphs with the same sizes and id layout.

Two forms:
  synth_id_arrays(...)  → integer arrays already in the reference's id layout (KG1 rank i → id 2i, KG2 rank
                          i → id 2i+1, modules/load/read.py:64-92) for kernel-level benchmarks;
  write_dataset(...)    → the on-disk layout of the reference README (rel_triples_{1,2}, attr_triples_{1,2},
                          ent_links, 721_5fold/<fold>/{train,valid,test}_links) for the drop-in CLI.
"""
import os

import numpy as np

SHAPES = {
    # entities per KG, relations (kg1, kg2), relation triples per KG, links train/valid/test, attributes
    "15K": dict(n_ent=15000, n_rel=(250, 200), n_tri=45000, links=(3000, 1500, 10500), n_attr=400),
    "100K": dict(n_ent=100000, n_rel=(300, 300), n_tri=300000, links=(20000, 10000, 70000), n_attr=400),
    "tiny": dict(n_ent=600, n_rel=(12, 10), n_tri=2400, links=(120, 60, 420), n_attr=20),
    "micro": dict(n_ent=40, n_rel=(4, 4), n_tri=120, links=(8, 4, 28), n_attr=5),   # CPU emulator runs (tests/emu)
}


def _zipf_probs(n, s):
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    return p / p.sum()


def _gen_graph(rng, n_ent, n_rel, n_tri, s=0.8):
    """Zipf-degree multigraph over entity ranks 0..n_ent-1 (rank = popularity order); every entity occurs."""
    pe, pr = _zipf_probs(n_ent, s), _zipf_probs(n_rel, 1.0)
    base_h = np.arange(n_ent)
    base_t = rng.choice(n_ent, size=n_ent, p=pe)
    extra = max(0, int(n_tri * 1.15) - n_ent)
    h = np.concatenate([base_h, rng.choice(n_ent, size=extra, p=pe)])
    t = np.concatenate([base_t, rng.choice(n_ent, size=extra, p=pe)])
    r = rng.choice(n_rel, size=h.size, p=pr)
    keep = h != t
    tri = np.unique(np.stack([h[keep], r[keep], t[keep]], 1), axis=0)
    if tri.shape[0] > n_tri:
        # keep the spanning part (first occurrence of every head) and a random subset of the rest
        first = np.zeros(tri.shape[0], dtype=bool)
        _, idx = np.unique(tri[:, 0], return_index=True)
        first[idx] = True
        rest = np.flatnonzero(~first)
        take = rng.choice(rest, size=max(0, n_tri - int(first.sum())), replace=False)
        sel = np.concatenate([np.flatnonzero(first), take])
        tri = tri[np.sort(sel)]
    return tri.astype(np.int64)


def _perturb(rng, tri, n_ent, n_rel2, n_rel1, rewire=0.2):
    """KG2 = copy of KG1 with relabelled relations and a fraction of rewired tails."""
    t2 = tri.copy()
    t2[:, 1] = (t2[:, 1] * 7 + 3) % n_rel2 if n_rel2 < n_rel1 else t2[:, 1] % n_rel2
    m = rng.random(t2.shape[0]) < rewire
    t2[m, 2] = rng.integers(0, n_ent, size=int(m.sum()))
    t2 = t2[t2[:, 0] != t2[:, 2]]
    return np.unique(t2, axis=0)


def _rank_ids(tri, n_ent):
    """Frequency-descending rank of every entity (ties by index, as a stand-in for the URI tie-break)."""
    freq = np.bincount(np.concatenate([tri[:, 0], tri[:, 2]]), minlength=n_ent)
    order = np.lexsort((-np.arange(n_ent), -freq))  # most frequent first
    rank = np.empty(n_ent, dtype=np.int64)
    rank[order] = np.arange(n_ent)
    return rank


def synth_id_arrays(shape="15K", seed=20200901, swapping=True, fold=1):
    """Returns a dict of int32 arrays in the reference id layout.

    keys: n_ent (total rows), n_rel, triples1/2 [T,3] (incl. swap triples when `swapping`, kgs.py:45-50),
    entities1/2, train/valid/test links [n,2].
    """
    cfg = SHAPES[shape]
    rng = np.random.default_rng(seed)
    n = cfg["n_ent"]
    g1 = _gen_graph(rng, n, cfg["n_rel"][0], cfg["n_tri"])
    g2 = _perturb(rng, g1, n, cfg["n_rel"][1], cfg["n_rel"][0])
    rank1, rank2 = _rank_ids(g1, n), _rank_ids(g2, n)
    id1, id2 = 2 * rank1, 2 * rank2 + 1            # read.py:69-79 interleaving (n1 == n2)
    # relation ids: interleaved the same way, overflow of the larger side appended
    r1n, r2n = cfg["n_rel"]
    rid1 = np.where(np.arange(r1n) < r2n, 2 * np.arange(r1n), 2 * r2n + (np.arange(r1n) - r2n))
    rid2 = np.where(np.arange(r2n) < r1n, 2 * np.arange(r2n) + 1, 2 * r1n + (np.arange(r2n) - r1n))
    t1 = np.stack([id1[g1[:, 0]], rid1[g1[:, 1]], id1[g1[:, 2]]], 1)
    t2 = np.stack([id2[g2[:, 0]], rid2[g2[:, 1]], id2[g2[:, 2]]], 1)
    # links: entity k of KG1 ↔ entity k of KG2 ; split 20/10/70 with a fold-dependent shuffle
    lrng = np.random.default_rng(seed + fold)
    perm = lrng.permutation(n)
    ntr, nva, nte = cfg["links"]
    links = np.stack([id1, id2], 1)
    train, valid, test = links[perm[:ntr]], links[perm[ntr:ntr + nva]], links[perm[ntr + nva:ntr + nva + nte]]
    if swapping:  # read.py:136-151: for every train link swap the linked entity into the other KG's triples
        a2b = -np.ones(2 * n, dtype=np.int64)
        a2b[train[:, 0]] = train[:, 1]
        b2a = -np.ones(2 * n, dtype=np.int64)
        b2a[train[:, 1]] = train[:, 0]

        def swap(tri, mp):
            hh, tt = mp[tri[:, 0]], mp[tri[:, 2]]
            s_h = tri[hh >= 0].copy(); s_h[:, 0] = hh[hh >= 0]
            s_t = tri[tt >= 0].copy(); s_t[:, 2] = tt[tt >= 0]
            return np.unique(np.concatenate([tri, s_h, s_t]), axis=0)
        t1, t2 = swap(t1, a2b), swap(t2, b2a)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return dict(shape=shape, n_ent=2 * n, n_rel=r1n + r2n, triples1=i32(t1), triples2=i32(t2),
                entities1=i32(np.sort(id1)), entities2=i32(np.sort(id2)),
                train_links=i32(train), valid_links=i32(valid), test_links=i32(test))


def write_dataset(folder, shape="15K", seed=20200901, folds=(1,), attr_per_ent=3):
    """Write a synthetic dataset in the reference's on-disk layout (README 'Dataset description')."""
    cfg = SHAPES[shape]
    rng = np.random.default_rng(seed)
    n = cfg["n_ent"]
    g1 = _gen_graph(rng, n, cfg["n_rel"][0], cfg["n_tri"])
    g2 = _perturb(rng, g1, n, cfg["n_rel"][1], cfg["n_rel"][0])
    os.makedirs(folder, exist_ok=True)
    e1 = lambda i: "http://kg1.synth/resource/E%07d" % i
    e2 = lambda i: "http://kg2.synth/entity/Q%07d" % i
    with open(os.path.join(folder, "rel_triples_1"), "w", encoding="utf8") as f:
        for h, r, t in g1.tolist():
            f.write("%s\thttp://kg1.synth/ontology/p%d\t%s\n" % (e1(h), r, e1(t)))
    with open(os.path.join(folder, "rel_triples_2"), "w", encoding="utf8") as f:
        for h, r, t in g2.tolist():
            f.write("%s\thttp://kg2.synth/prop/P%d\t%s\n" % (e2(h), r, e2(t)))
    pa = _zipf_probs(cfg["n_attr"], 1.0)
    for side, ename in ((1, e1), (2, e2)):
        with open(os.path.join(folder, "attr_triples_%d" % side), "w", encoding="utf8") as f:
            ents = np.repeat(np.arange(n), attr_per_ent)
            attrs = rng.choice(cfg["n_attr"], size=ents.size, p=pa)
            for e, a in zip(ents.tolist(), attrs.tolist()):
                f.write('%s\thttp://kg%d.synth/attr/a%d\t"v%d"\n' % (ename(e), side, a, (e * 31 + a) % 997))
    with open(os.path.join(folder, "ent_links"), "w", encoding="utf8") as f:
        for i in range(n):
            f.write("%s\t%s\n" % (e1(i), e2(i)))
    ntr, nva, nte = cfg["links"]
    for fold in folds:
        sub = os.path.join(folder, "721_5fold", str(fold))
        os.makedirs(sub, exist_ok=True)
        perm = np.random.default_rng(seed + fold).permutation(n)
        parts = {"train_links": perm[:ntr], "valid_links": perm[ntr:ntr + nva], "test_links": perm[ntr + nva:ntr + nva + nte]}
        for name, idx in parts.items():
            with open(os.path.join(sub, name), "w", encoding="utf8") as f:
                for i in idx.tolist():
                    f.write("%s\t%s\n" % (e1(i), e2(i)))
    return folder
