"""Alignment search with the reference's signatures (modules/finding/alignment.py): greedy_alignment :13,
stable_alignment :87, calculate_rank :146, galeshapley :171, arg_sort :136, retrieve_topk_alignment :227.
greedy search / ranking run on the GPU without materialising or sorting the similarity matrix."""
import itertools
import time

import numpy as np

from openea_b200 import finding as _f
from openea_b200.modules.finding.similarity import sim
from openea_b200.modules.utils.util import task_divide, merge_dic  # noqa: F401  (re-exported like the reference)


def greedy_alignment(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate):
    """Greedy (arg-max) alignment of row i of embed1 against embed2; the gold match of row i is column i.
    Returns (set of (i, argmax_i), hits@1 in %, MR, MRR) and prints the reference's result line."""
    return _f.greedy_alignment(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate)


def calculate_rank(idx, sim_mat, top_k, accurate, total_num):
    """Rank statistics of a block of similarity rows whose gold columns are `idx` (host helper kept for
    callers that already hold a matrix): returns (mr, mrr, hits list, {(gold, argmax)})."""
    assert 1 in top_k
    s = np.asarray(sim_mat)
    gold = np.asarray(list(idx), dtype=np.int64)
    g = s[np.arange(len(gold)), gold]
    cols = np.arange(s.shape[1])[None, :]
    rank = (s > g[:, None]).sum(1) + ((s == g[:, None]) & (cols < gold[:, None])).sum(1)
    best = s.argmax(1)
    hits = [int((rank < k).sum()) for k in top_k]
    mr = float((rank + 1).sum()) / total_num
    mrr = float((1.0 / (rank + 1)).sum()) / total_num
    return mr, mrr, hits, {(int(a), int(b)) for a, b in zip(gold, best)}


def arg_sort(idx, sim_mat, prefix1, prefix2):
    order = np.argsort(-np.asarray(sim_mat), axis=1, kind="stable")
    return {prefix1 + str(idx[i]): [prefix2 + str(r) for r in order[i]] for i in range(len(idx))}


def stable_alignment(embed1, embed2, metric, normalize, csls_k, nums_threads, cut=100, sim_mat=None):
    """Stable (Gale–Shapley) alignment, reference signature (alignment.py:87).  Without a precomputed `sim_mat` the whole
    search runs on the device: K3's top-`cut` lists are the preference lists (only the first `cut` entries of a list can be
    reached in `cut` proposal rounds) and oea_gale_shapley plays the rounds; `nums_threads` is accepted and ignored.  With a
    host `sim_mat` the reference's own route (full argsort lists + the Python loop below) is kept."""
    t = time.time()
    if sim_mat is None:
        match, rounds = _f.stable_matching(embed1, embed2, metric, normalize, csls_k, cut)
        print("generating candidate lists costs time {:.3f} s ".format(time.time() - t))
        t = time.time()
        m = match.cpu().numpy()
        held = m >= 0
        n = int((m[held] == np.arange(len(m))[held]).sum())
        cost = time.time() - t
        print("stable alignment precision = {:.3f}%, time = {:.3f} s ".format(n / max(1, int(held.sum())) * 100, cost))
        return
    n1, n2 = sim_mat.shape
    kg1_candidates = arg_sort(list(range(n1)), sim_mat, 'x_', 'y_')
    kg2_candidates = arg_sort(list(range(n2)), sim_mat.T, 'y_', 'x_')
    print("generating candidate lists costs time {:.3f} s ".format(time.time() - t))
    t = time.time()
    matching = galeshapley(kg1_candidates, kg2_candidates, cut)
    n = sum(1 for i, j in matching.items() if int(i.split('_')[-1]) == int(j.split('_')[-1]))
    cost = time.time() - t
    print("stable alignment precision = {:.3f}%, time = {:.3f} s ".format(n / len(matching) * 100, cost))


def galeshapley(suitor_pref_dict, reviewer_pref_dict, max_iteration):
    """Suitor-proposing deferred acceptance, at most `max_iteration` proposal rounds.  A suitor proposes to
    the head of its list; a taken reviewer switches only if it ranks the new suitor higher, otherwise the
    suitor strikes that reviewer off."""
    rank_of = {r: {s: pos for pos, s in enumerate(prefs)} for r, prefs in reviewer_pref_dict.items()}
    matching, holder = {}, {}
    free = list(suitor_pref_dict.keys())
    for _ in range(max_iteration):
        if not free:
            break
        for s in free:
            r = suitor_pref_dict[s][0]
            if r not in holder:
                matching[s] = r
                holder[r] = s
            else:
                cur = holder[r]
                if rank_of[r][s] < rank_of[r][cur]:
                    del matching[cur]
                    matching[s] = r
                    holder[r] = s
                else:
                    suitor_pref_dict[s].remove(r)
        free = [s for s in suitor_pref_dict if s not in matching]
    return matching


def retrieve_topk_alignment(kg1_source_ents, kg1_embeddings, kg2_candidates, kg2_embeddings, session=None, k=1,
                            metric='inner', normalize=False, csls_k=0, output_path=None):
    """Top-k candidates of every source entity with their similarity: [(ent1, ent2, sim)].  The embedding
    arguments are device tables (or arrays); `session` is accepted for signature compatibility."""
    lookup = lambda emb, ids: emb.lookup(ids) if hasattr(emb, "lookup") else np.asarray(emb)[ids]
    e1, e2, d = _f._prep_pair(lookup(kg1_embeddings, kg1_source_ents), lookup(kg2_embeddings, kg2_candidates),
                              metric, normalize)
    r = c = None
    if csls_k > 0:
        r, c = _f.csls_offsets(e1, e2, d, metric, csls_k)
    res = _f.topk(e1, e2, d, metric, k, r, c, want=("val", "idx"))
    idx, val = res["idx"].cpu().numpy(), res["val"].cpu().numpy()
    out = [(kg1_source_ents[i], kg2_candidates[int(j)], float(v)) for i in range(idx.shape[0])
           for j, v in zip(idx[i], val[i])]
    if output_path is not None:
        with open(output_path, 'w', encoding='utf8') as fh:
            fh.writelines("%s\t%s\t%s\n" % t for t in out)
        print(output_path, "saved")
    return out
