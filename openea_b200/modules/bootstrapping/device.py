"""BootEA's bootstrapping step with the working set kept on the device.

The reference edits a Python set of labelled pairs with dict lookups per pair (approaches/bootea.py:19-103), matches
candidates with graph-tool / igraph (modules/bootstrapping/alignment_finder.py:83-140) and expands every new pair into
swap triples through `rt_dict` / `hr_dict` loops (bootea.py:105-121).  At the 100K shape those host loops cost 1–2 s
per iteration while the 20 training epochs between them take ≈ 0.1 s on the GPU.  Here every step is index arithmetic
on tensors that stay where the embeddings are (CUDA in the product; the same code runs on CPU tensors in the tests):

  candidates   find_alignment_device (K3 top-k kernel + threshold)                       → rows, cols, vals
  matching     `greedy_matching`: the greedy-by-weight matching of `alignment_finder.mwgm_scipy`, computed by rounds of
               locally dominant edges (an edge that is the best remaining edge of both its row and its column is in
               the greedy matching; remove its row and column; repeat) — identical result, a handful of rounds
  editing      `edit_labels_x` / `edit_labels_y`: update_labeled_alignment_x / _y on a label vector
               label[i] = j or −1 with gathered pair similarities
  swap triples `swap_triples`: generate_supervised_triples as two masked gathers over the [T, 3] triple tensor

The printed diagnostics keep the reference's wording.  The exact matcher of the reference (graph-tool's heuristic
max-cardinality matching) is not available offline, so — as for the host version — agreement with it is unpinned; the
device functions are pinned against the host functions of approaches/bootea.py (tests/test_bootstrapping_device.py).
"""
import torch


def pair_sim(e1, e2, i, j):
    """Similarity of selected pairs without the n1×n2 matrix (what `sim_mat[i, j]` reads in bootea.py:35-78)."""
    return (e1.index_select(0, i) * e2.index_select(0, j)).sum(1)


def _report(context, i, j):
    n = int(i.numel())
    if n == 0:
        print("{}, empty aligned pairs".format(context))
        return
    num = int((i == j).sum())
    print("{}, right alignment: {}/{}={:.3f}".format(context, num, n, num / n))


def greedy_matching(rows, cols, vals, n_rows, n_cols):
    """Mask of the edges the sequential greedy matching picks when edges are visited by descending weight
    (ties: earlier edge first, as a stable sort does)."""
    e = rows.numel()
    dev = rows.device
    chosen = torch.zeros(e, dtype=torch.bool, device=dev)
    if e == 0:
        return chosen
    rows, cols = rows.long(), cols.long()
    rank = torch.empty(e, dtype=torch.long, device=dev)
    rank[torch.argsort(-vals, stable=True)] = torch.arange(e, device=dev)     # 0 = the edge greedy visits first
    alive = torch.ones(e, dtype=torch.bool, device=dev)
    big = e
    while True:
        r = torch.where(alive, rank, torch.full_like(rank, big))
        row_best = torch.full((n_rows,), big, dtype=torch.long, device=dev).scatter_reduce_(0, rows, r, "amin")
        col_best = torch.full((n_cols,), big, dtype=torch.long, device=dev).scatter_reduce_(0, cols, r, "amin")
        dominant = alive & (r == row_best[rows]) & (r == col_best[cols])
        if not bool(dominant.any()):
            return chosen
        chosen |= dominant
        row_taken = torch.zeros(n_rows, dtype=torch.bool, device=dev).index_fill_(0, rows[dominant], True)
        col_taken = torch.zeros(n_cols, dtype=torch.bool, device=dev).index_fill_(0, cols[dominant], True)
        alive &= ~(row_taken[rows] | col_taken[cols])


def edit_labels_x(label, mi, mj, e1, e2):
    """update_labeled_alignment_x (bootea.py:35-55): a newly matched pair (i, j) replaces i's previous partner unless
    the previous one is more similar.  `label` [n1] int64 (−1 = unlabelled) is edited in place and returned."""
    pre = label[mi]
    has = pre >= 0
    s_new = pair_sim(e1, e2, mi, mj)
    s_pre = pair_sim(e1, e2, mi, pre.clamp(min=0))
    take = ~has | (s_new >= s_pre)
    n2 = int(((pre == mi) & (mj != mi)).sum())
    n1 = int((has & take & (pre == mi) & (mj != mi)).sum())
    print("update wrongly: ", n1, "greedy update wrongly: ", n2)
    label[mi[take]] = mj[take]
    i = torch.nonzero(label >= 0).flatten()
    _report("after editing (<-)", i, label[i])
    return label


def edit_labels_y(label, e1, e2):
    """update_labeled_alignment_y (bootea.py:57-78): a KG2 index claimed by several KG1 indices stays with the most
    similar claimant; the others lose their label."""
    i = torch.nonzero(label >= 0).flatten()
    if i.numel():
        j = label[i]
        s = pair_sim(e1, e2, i, j)
        by_sim = torch.argsort(-s, stable=True)
        order = by_sim[torch.argsort(j[by_sim], stable=True)]       # grouped by j, most similar claimant first
        js = j[order]
        first = torch.ones_like(js, dtype=torch.bool)
        first[1:] = js[1:] != js[:-1]
        label[i[order[~first]]] = -1
    i = torch.nonzero(label >= 0).flatten()
    _report("after editing (->)", i, label[i])
    return label


def bootstrap_labels(e1, e2, label, candidates):
    """One bootstrapping pass (bootea.py:19-32) on device tensors.  e1 / e2: row-normalised reference-entity embeddings
    [n, d]; label: [n1] int64 state (−1 = unlabelled), edited in place; candidates: (rows, cols, vals) from
    find_alignment_device, or None when nothing passed the threshold.  Returns (label, i, j) with i / j the labelled
    pairs as index tensors into the reference-entity lists."""
    import time
    if candidates is not None and candidates[0].numel():
        rows, cols, vals = candidates
        _report("after filtering by sim and nearest k", rows, cols)
        t = time.time()
        sel = greedy_matching(rows, cols, vals, e1.shape[0], e2.shape[0])
        mi, mj = rows[sel].long(), cols[sel].long()
        _report("after mwgm", mi, mj)
        print("mwgm costs time: {:.3f} s".format(time.time() - t))
        edit_labels_x(label, mi, mj, e1, e2)
        edit_labels_y(label, e1, e2)
    i = torch.nonzero(label >= 0).flatten()
    return label, i, label[i]


def swap_triples(triples, src, dst, n_ent):
    """generate_supervised_triples for one KG (bootea.py:105-121): for every aligned pair (src[p] → dst[p]) each triple
    with head src[p] yields (dst[p], r, t) and each triple with tail src[p] yields (h, r, dst[p]).
    triples [T, 3] int32 (the KG's LOCAL relation triples), src / dst int64 index tensors → [M, 3] int32."""
    to = torch.full((n_ent,), -1, dtype=torch.int32, device=triples.device)
    to[src] = dst.to(torch.int32)
    h_to, t_to = to[triples[:, 0].long()], to[triples[:, 2].long()]
    a = triples[h_to >= 0].clone()
    a[:, 0] = h_to[h_to >= 0]
    b = triples[t_to >= 0].clone()
    b[:, 2] = t_to[t_to >= 0]
    return torch.cat([a, b], 0)


def pos_batch(tri1, tri2, step, batch_size):
    """generate_pos_batch (bootea.py:124-131) on tensors → [3, n] int32 (h | r | t rows) for the fed scorer."""
    n1, n2 = tri1.shape[0], tri2.shape[0]
    num1 = int(n1 / (n1 + n2) * batch_size)
    num2 = batch_size - num1
    b = torch.cat([tri1[step * num1:min(step * num1 + num1, n1)], tri2[step * num2:min(step * num2 + num2, n2)]], 0)
    return b.t().contiguous()
