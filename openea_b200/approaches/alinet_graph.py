"""AliNet's graph builders on arrays (the adjacency builders of approaches/alinet.py:155-204,250-287,399-436).

The reference joins the triple list with itself through pandas and walks the result row by row in Python
(`generate_2hop_triples`, alinet.py:250-287); the line-by-line mirror in approaches/alinet.py does the same walk with
dicts.  On the synthetic 15K shape that is 5.6 million two-hop paths and ≈ 30 s of host time before the first training
step (minutes at 100K).  Here the same sets are computed with sorted-key joins on int64 arrays:

  two-hop paths   for every triple (h, r_x, m) all triples (m, r_y, t): a CSR-by-head expansion with np.repeat, in chunks
  "not already a neighbour"   (h, t) ∉ {(h', t') of the KG}: one searchsorted against the sorted packed keys
                  (`t not in out[h]` and `h not in in[t]` of alinet.py:266-267 are the same condition)
  relation patterns   counts per (r_x, r_y) over the surviving joined rows (alinet.py:268 counts rows, not distinct paths);
                  all but the 5 most frequent patterns are kept (alinet.py:277-279)
  outputs         AliNet only uses the (h, t) ends of the selected paths plus a self loop per path head
                  (`no_weighted_adj(…, list(two_hop_triples))`, alinet.py:1068-1072), so the result is a pair array

Ties among the 5 most frequent patterns are ordered by dict insertion order in the reference, i.e. by the hash order of
a Python set — unspecified; here they are ordered by (count desc, r_x, r_y).
tests/test_gnn_host.py checks these functions against the set / dict versions of approaches/alinet.py (which are
pinned against the reference's own source) on random graphs.
"""
import numpy as np
import scipy.sparse as sp

from openea_b200 import gnn


def pack(a, b, n):
    return a.astype(np.int64) * np.int64(n) + b.astype(np.int64)


def unique_sorted(keys):
    """Distinct values of an int64 key array, ascending (sort + neighbour compare: several times faster than np.unique's
    hash path on tens of millions of keys)."""
    if len(keys) == 0:
        return keys
    keys = np.sort(keys)
    keep = np.empty(len(keys), dtype=bool)
    keep[0] = True
    np.not_equal(keys[1:], keys[:-1], out=keep[1:])
    return keys[keep]


def linked_only(tri, linked_ids, n):
    """remove_unlinked_triples (alinet.py:206-216): triples whose head AND tail are linked entities."""
    mask = np.zeros(n, dtype=bool)
    mask[np.asarray(list(linked_ids), dtype=np.int64)] = True
    return tri[mask[tri[:, 0]] & mask[tri[:, 2]]]


def enhance(tri1, tri2, ents1, ents2, n):
    """enhance_triples (alinet.py:399-436): every triple between two seed-linked entities of one KG, projected into the
    other KG, unless the projected (head, tail) edge already exists there.  Returns (new1, new2) [m, 3] arrays."""
    def project(src, dst, frm, to):
        m = np.full(n, -1, dtype=np.int64)
        m[np.asarray(frm, dtype=np.int64)] = np.asarray(to, dtype=np.int64)
        h, t = m[src[:, 0]], m[src[:, 2]]
        ok = (h >= 0) & (t >= 0)
        cand = np.stack([h[ok], src[ok, 1].astype(np.int64), t[ok]], 1)
        have = unique_sorted(pack(dst[:, 0], dst[:, 2], n))
        keys = pack(cand[:, 0], cand[:, 2], n)
        pos = np.searchsorted(have, keys)
        present = (pos < len(have)) & (have[np.minimum(pos, len(have) - 1)] == keys) if len(have) else np.zeros(len(keys), bool)
        out = cand[~present]
        return np.unique(out, axis=0) if len(out) else out.reshape(0, 3)
    new2 = project(tri1, tri2, ents1, ents2)
    new1 = project(tri2, tri1, ents2, ents1)
    return new1, new2


def two_hop_pairs(tri, linked_ids, n, chunk=2_000_000):
    """generate_2hop_triples (alinet.py:250-287) reduced to what AliNet consumes: the distinct (h, t) ends of the
    selected two-hop paths and one (h, h) self loop per selected path head.  `tri` [T, 3] int array of ALL triples of the
    KG (the "already a neighbour" test uses all of them, the join only the linked ones).  Returns [m, 2] int64."""
    ht_all = unique_sorted(pack(tri[:, 0], tri[:, 2], n))
    t = linked_only(tri, linked_ids, n) if linked_ids is not None else tri
    if len(t) == 0:
        return np.zeros((0, 2), dtype=np.int64)
    order = np.argsort(t[:, 0], kind="stable")
    by_head = t[order]
    start = np.searchsorted(by_head[:, 0], np.arange(n + 1))
    deg = np.diff(start)
    n_rel = int(t[:, 1].max()) + 1
    n_pat = np.int64(n_rel) * n_rel
    assert float(n) * n * float(n_pat) < 2.0 ** 62, "(h, t, pattern) does not fit one int64 key"
    fan = deg[t[:, 2]]                                   # second hops of every first hop
    bounds = np.concatenate([[0], np.cumsum(fan)])
    pat_count = np.zeros(n_rel * n_rel, dtype=np.int64)
    kept = []                                            # (h·n + t, pattern) of surviving rows, de-duplicated per chunk
    lo = 0
    while lo < len(t):
        hi = int(np.searchsorted(bounds, bounds[lo] + chunk, side="right")) - 1
        hi = max(hi, lo + 1)
        f = fan[lo:hi]
        total = int(f.sum())
        if total:
            first = np.repeat(np.arange(lo, hi), f)
            within = np.arange(total) - np.repeat(bounds[lo:hi] - bounds[lo], f)
            second = start[t[first, 2]] + within
            h, rx = t[first, 0].astype(np.int64), t[first, 1].astype(np.int64)
            ry, tt = by_head[second, 1].astype(np.int64), by_head[second, 2].astype(np.int64)
            key = h * np.int64(n) + tt
            pos = np.searchsorted(ht_all, key)
            new = ~((pos < len(ht_all)) & (ht_all[np.minimum(pos, len(ht_all) - 1)] == key))
            pat = rx[new] * n_rel + ry[new]
            pat_count += np.bincount(pat, minlength=n_rel * n_rel)          # per joined row (alinet.py:268)
            kept.append(unique_sorted(key[new] * n_pat + pat))                  # one int64 per (h, t, r_x, r_y)
        lo = hi
    if not kept:
        return np.zeros((0, 2), dtype=np.int64)
    rows = unique_sorted(np.concatenate(kept))               # distinct (h, t, r_x, r_y): the reference's set of quadruples
    print("total 2-hop neighbors:", len(rows))
    seen = np.flatnonzero(pat_count)
    print("total 2-hop relation patterns:", len(seen))
    ranked = seen[np.lexsort((seen, -pat_count[seen]))]  # count descending, then (r_x, r_y)
    selected = np.zeros(n_rel * n_rel, dtype=bool)
    selected[ranked[5:]] = True
    print("selected relation patterns:", int(selected.sum()))
    ok = unique_sorted(rows[selected[rows % n_pat]] // n_pat)          # (h, t) keys of the selected paths
    h = unique_sorted(ok // n)
    keys = unique_sorted(np.concatenate([ok, h * np.int64(n) + h]))    # plus one self loop per path head
    return np.stack([keys // n, keys % n], 1)


def unweighted_adj(n, pairs):
    """no_weighted_adj (alinet.py:155-178) from an [m, 2] (head, tail) array: symmetric 0/1 neighbour matrix (multi-edges
    count once) with self loops, then D^-½ (A + I)ᵀ D^-½."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    rows = np.concatenate([pairs[:, 0], pairs[:, 1]])
    cols = np.concatenate([pairs[:, 1], pairs[:, 0]])
    m = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
    m.data[:] = 1.0
    return gnn.normalize_adj(m + sp.eye(n))


def relation_index(tri):
    """generate_rel_ht (alinet.py:728-740) as arrays: (relation ids, row pointer, (h, t) pairs grouped by relation)."""
    order = np.argsort(tri[:, 1], kind="stable")
    by_rel = tri[order]
    rels, first = np.unique(by_rel[:, 1], return_index=True)
    ptr = np.concatenate([first, [len(by_rel)]])
    return rels, ptr, by_rel[:, [0, 2]]


def build(kg1_tri, kg2_tri, sup1, sup2, linked_ids, n):
    """Everything AliNet.init derives from the two KGs (alinet.py:1041-1079): (one-hop adjacency, two-hop adjacency,
    the linked + enhanced triple array that feeds generate_rel_ht)."""
    new1, new2 = enhance(kg1_tri, kg2_tri, sup1, sup2, n)
    print("after enhanced:", len(new1), len(new2))
    every = np.concatenate([kg1_tri.astype(np.int64), kg2_tri.astype(np.int64), new1, new2])
    n_rel = int(every[:, 1].max()) + 1 if len(every) else 1
    keys = unique_sorted((every[:, 0] * n_rel + every[:, 1]) * np.int64(n) + every[:, 2])       # distinct (h, r, t)
    every = np.stack([keys // n // n_rel, keys // n % n_rel, keys % n], 1)
    tri = linked_only(every, linked_ids, n)
    one_adj = unweighted_adj(n, tri[:, [0, 2]])
    two = np.concatenate([two_hop_pairs(kg1_tri, linked_ids, n), two_hop_pairs(kg2_tri, linked_ids, n)])
    two_adj = unweighted_adj(n, two)
    return one_adj, two_adj, tri
