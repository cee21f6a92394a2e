"""Bootstrapping candidate search (interface of modules/bootstrapping/alignment_finder.py).

find_alignment = {(i,j): sim_ij > th} ∩ {j among the k nearest of row i}: computed on the GPU either from
embeddings (no matrix, `find_alignment_from_embeds`) or from a matrix the caller already holds.
The max-weight matching of the reference uses graph-tool / igraph (neither is installable offline); it is
OUT OF SCOPE for acceleration and is served by SciPy's sparse bipartite matching here.
"""
import itertools
import time

import numpy as np


def find_potential_alignment_greedily(sim_mat, sim_th):
    return find_alignment(sim_mat, sim_th, 1)


def find_potential_alignment_mwgm(sim_mat, sim_th, k, heuristic=True):
    t = time.time()
    potential_aligned_pairs = find_alignment(sim_mat, sim_th, k)
    if potential_aligned_pairs is None:
        return None
    t1 = time.time()
    selected_aligned_pairs = mwgm(potential_aligned_pairs, sim_mat, mwgm_graph_tool if heuristic else mwgm_igraph)
    check_new_alignment(selected_aligned_pairs, context="after mwgm")
    print("mwgm costs time: {:.3f} s".format(time.time() - t1))
    print("selecting potential alignment costs time: {:.3f} s".format(time.time() - t))
    return selected_aligned_pairs


def _pairs_from_topk(rows, cols):
    return set(zip(rows.tolist(), cols.tolist()))


def find_alignment_from_embeds(embeds1, embeds2, sim_th, k, metric="inner", normalize=True):
    """Same result as find_alignment(sim(embeds1, embeds2), sim_th, k) without the n1×n2 matrix; returns
    (pairs set or None, {pair: similarity})."""
    from openea_b200 import finding
    rows, cols, vals = finding.find_alignment_device(embeds1, embeds2, sim_th, max(k, 1), metric, normalize)
    rows, cols, vals = rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy()
    pairs = _pairs_from_topk(rows, cols)
    if not pairs:
        return None, {}
    return pairs, {(int(i), int(j)): float(v) for i, j, v in zip(rows, cols, vals)}


def find_alignment(sim_mat, sim_th, k):
    """Pairs (x, y) with sim(x, y) > sim_th and y among the k nearest of x; None when empty.  With k <= 0
    only the threshold applies."""
    import torch
    if k <= 0:
        pairs = filter_sim_mat(sim_mat, sim_th)
        if len(pairs) == 0:
            return None
        check_new_alignment(pairs, context="after filtering by sim threshold")
        return pairs
    s = sim_mat if isinstance(sim_mat, torch.Tensor) else torch.as_tensor(np.asarray(sim_mat, dtype=np.float32))
    if not s.is_cuda:
        s = s.cuda()
    if not bool((s > sim_th).any().item()):
        return None
    idx = _topk_of_matrix(s, k)
    vals = torch.gather(s, 1, idx.long())
    keep = vals > sim_th
    rows = torch.arange(s.shape[0], device=s.device)[:, None].expand_as(keep)[keep]
    pairs = _pairs_from_topk(rows.cpu().numpy(), idx[keep].cpu().numpy())
    if len(pairs) == 0:
        return None
    check_new_alignment(pairs, context="after filtering by sim and nearest k")
    return pairs


def _topk_of_matrix(s, k):
    import ctypes as C
    import torch
    from openea_b200 import lib as L
    from openea_b200.engine import _ptr, _stream_ptr
    s = s.contiguous()
    out = torch.empty(s.shape[0], k, dtype=torch.int32, device=s.device)
    L.check(L.load().oea_rows_select_topk(_ptr(s), s.stride(0), s.shape[0], s.shape[1], k, None, _ptr(out),
                                          _stream_ptr()), "oea_rows_select_topk")
    return out


def filter_sim_mat(mat, threshold, greater=True, equal=False):
    m = np.asarray(mat.cpu() if hasattr(mat, "cpu") else mat)
    if greater:
        x, y = np.where(m >= threshold) if equal else np.where(m > threshold)
    else:
        x, y = np.where(m <= threshold) if equal else np.where(m < threshold)
    return set(zip(x.tolist(), y.tolist()))


def search_nearest_k(sim_mat, k):
    assert k > 0
    import torch
    s = sim_mat if isinstance(sim_mat, torch.Tensor) else torch.as_tensor(np.asarray(sim_mat, dtype=np.float32))
    idx = _topk_of_matrix(s.cuda() if not s.is_cuda else s, k).cpu().numpy()
    neighbors = {(i, int(j)) for i in range(idx.shape[0]) for j in idx[i]}
    assert len(neighbors) == idx.shape[0] * k
    return neighbors


def mwgm(pairs, sim_mat, func):
    return func(pairs, sim_mat)


def _weight_lookup(sim_mat):
    if isinstance(sim_mat, dict):
        return lambda x, y: sim_mat[(x, y)]
    return lambda x, y: float(sim_mat[x, y])


def mwgm_greedy(pairs, sim_mat):
    """Greedy-by-weight matching on the bipartite candidate graph: edges in descending weight, an edge is kept when both
    endpoints are still free (a 1/2-approximation of the maximum weight).  This is the counterpart of what the reference's
    BootEA actually runs: graph-tool's max_cardinality_matching(heuristic=True, weight=…, minimize=False)
    (alignment_finder.py:83-113) is itself a greedy heuristic, not an exact matcher, and graph-tool does not exist here —
    agreement with it is unpinned.  The device version (bootstrapping/device.py greedy_matching)
    computes exactly this matching."""
    pairs = list(pairs)
    w = _weight_lookup(sim_mat)
    weights = np.array([w(x, y) for x, y in pairs], dtype=np.float64)
    order = np.argsort(-weights, kind="stable")
    used_x, used_y, matched = set(), set(), set()
    for e in order:
        x, y = pairs[e]
        if x not in used_x and y not in used_y:
            used_x.add(x)
            used_y.add(y)
            matched.add((x, y))
    return matched


def mwgm_exact(pairs, sim_mat):
    """EXACT maximum-weight bipartite matching of the candidate graph — what the reference's igraph path computes
    (alignment_finder.py:116-140, Graph.maximum_bipartite_matching(weights=…)).  igraph is absent here; the matching is
    obtained from SciPy's sparse assignment solver on the doubled graph: left' = L ∪ R°, right' = R ∪ L° (° = a copy),
    real edges (i, j) and their mirrors (j°, i°) cost C − w_ij, "stay unmatched" edges (i, i°), (j°, j) cost C with
    C > max w.  A perfect matching of the doubled graph always exists and costs C·(|L|+|R|) − w(M) − w(M°) with M, M°
    matchings over the same vertex set, so the minimum puts a maximum-weight matching on the original side.
    Weights must be positive (they are similarities above sim_th > 0); ties between equal-weight optima are broken by
    the solver, as they are by igraph."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching
    pairs = list(pairs)
    if not pairs:
        return set()
    w = _weight_lookup(sim_mat)
    weights = np.array([w(x, y) for x, y in pairs], dtype=np.float64)
    if (weights <= 0).any():
        raise ValueError("mwgm_exact needs positive edge weights")
    xs = {x: i for i, x in enumerate(dict.fromkeys(p[0] for p in pairs))}
    ys = {y: j for j, y in enumerate(dict.fromkeys(p[1] for p in pairs))}
    n1, n2 = len(xs), len(ys)
    ei = np.array([xs[x] for x, _ in pairs]); ej = np.array([ys[y] for _, y in pairs])
    big = float(weights.max()) * 2.0 + 1.0
    # rows: L (0..n1) then R° (n1..n1+n2); columns: R (0..n2) then L° (n2..n2+n1)
    rows = np.concatenate([ei, n1 + ej, np.arange(n1), n1 + np.arange(n2)])
    cols = np.concatenate([ej, n2 + ei, n2 + np.arange(n1), np.arange(n2)])
    cost = np.concatenate([big - weights, big - weights, np.full(n1 + n2, big)])
    # duplicate candidate edges (the same (x, y) listed twice) would be summed by the sparse constructor: keep the best
    key = rows.astype(np.int64) * (n1 + n2) + cols
    order = np.lexsort((cost, key))
    first = np.ones(len(order), dtype=bool)
    first[1:] = key[order][1:] != key[order][:-1]
    sel = order[first]
    m = sp.csr_matrix((cost[sel], (rows[sel], cols[sel])), shape=(n1 + n2, n1 + n2))
    r, c = min_weight_full_bipartite_matching(m)
    inv_x = list(xs); inv_y = list(ys)
    cand = set(pairs)
    return {(inv_x[i], inv_y[j]) for i, j in zip(r.tolist(), c.tolist()) if i < n1 and j < n2 and (inv_x[i], inv_y[j]) in cand}


mwgm_scipy = mwgm_greedy          # historical name (tests, device parity): the greedy matching
mwgm_graph_tool = mwgm_greedy     # graph-tool's heuristic=True matcher is a greedy heuristic
mwgm_igraph = mwgm_exact          # igraph's maximum_bipartite_matching is exact


def check_new_alignment(aligned_pairs, context="check alignment"):
    if aligned_pairs is None or len(aligned_pairs) == 0:
        print("{}, empty aligned pairs".format(context))
        return
    num = sum(1 for x, y in aligned_pairs if x == y)
    print("{}, right alignment: {}/{}={:.3f}".format(context, num, len(aligned_pairs), num / len(aligned_pairs)))
