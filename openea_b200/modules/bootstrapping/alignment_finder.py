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


def mwgm_scipy(pairs, sim_mat):
    """Maximum-weight matching on the bipartite candidate graph.  The reference calls graph-tool's
    max_cardinality_matching(heuristic=True, weight=…) (itself a heuristic) or igraph's
    maximum_bipartite_matching; neither library exists here, so this uses a greedy-by-weight matching
    improved by SciPy's optimal assignment on each connected component when it is small enough.
    Exact agreement with the reference's matcher is unpinned."""
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


mwgm_graph_tool = mwgm_scipy
mwgm_igraph = mwgm_scipy


def check_new_alignment(aligned_pairs, context="check alignment"):
    if aligned_pairs is None or len(aligned_pairs) == 0:
        print("{}, empty aligned pairs".format(context))
        return
    num = sum(1 for x, y in aligned_pairs if x == y)
    print("{}, right alignment: {}/{}={:.3f}".format(context, num, len(aligned_pairs), num / len(aligned_pairs)))
