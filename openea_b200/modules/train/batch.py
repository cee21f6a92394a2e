"""Batch slicing, negative sampling and ε-truncated neighbour search with the reference's names
(modules/train/batch.py).  The training loop of BasicModel samples ON DEVICE (oea_triple_score_sampled);
the host functions here keep the reference's call signatures and sampling rules for callers that feed
index batches themselves (and for parity tests).  Neighbour search runs on the GPU (K3 + radix select).
"""
import random

import numpy as np

from openea_b200.modules.utils.util import merge_dic


# ---- positive batches --------------------------------------------------------------------------------
def _split_sizes(n1, n2, batch_size):
    size1 = int(n1 / (n1 + n2) * batch_size)          # float division, multiply, truncate
    return size1, batch_size - size1


def generate_pos_triples(triples, batch_size, step, is_fixed_size=False):
    """The step-th contiguous slice of `triples`, clipped at the end; optionally topped up from the front."""
    lo = step * batch_size
    hi = min(lo + batch_size, len(triples))
    batch = triples[lo:hi]
    if is_fixed_size and len(batch) < batch_size:
        batch += triples[:batch_size - len(batch)]
    return batch


def generate_pos_batch(triple_list1, triple_list2, batch_size, step):
    size1, size2 = _split_sizes(len(triple_list1), len(triple_list2), batch_size)
    return generate_pos_triples(triple_list1, size1, step) + generate_pos_triples(triple_list2, size2, step)


def generate_pos_batch_queue(triple_list1, triple_list2, batch_size, steps, out_queue):
    for step in steps:
        out_queue.put(generate_pos_batch(triple_list1, triple_list2, batch_size, step))


# ---- negatives ---------------------------------------------------------------------------------------
def generate_neg_triples_fast(pos_batch, all_triples_set, entities_list, neg_triples_num, neighbor=None, max_try=10):
    """Per positive: up to `max_try` rounds; a round flips one coin (head or tail) for ALL still-missing
    negatives, draws that many distinct candidates (the corrupted entity's ε-neighbours, else the KG's entity
    list), and keeps those that are not known triples — the last round keeps everything."""
    neighbor = neighbor or {}
    neg_batch = []
    for head, relation, tail in pos_batch:
        found = []
        missing = neg_triples_num
        head_pool = neighbor.get(head, entities_list)
        tail_pool = neighbor.get(tail, entities_list)
        for attempt in range(max_try):
            if np.random.binomial(1, 0.5):
                drawn = {(h2, relation, tail) for h2 in random.sample(head_pool, missing)}
            else:
                drawn = {(head, relation, t2) for t2 in random.sample(tail_pool, missing)}
            if attempt == max_try - 1:
                found += list(drawn)
                break
            found += list(drawn - all_triples_set)
            if len(found) == neg_triples_num:
                break
            missing = neg_triples_num - len(found)
        assert len(found) == neg_triples_num
        neg_batch.extend(found)
    assert len(neg_batch) == neg_triples_num * len(pos_batch)
    return neg_batch


def generate_neg_triples(pos_batch, all_triples_set, entities_list, neg_triples_num, neighbor=None, max_try=10):
    """One negative at a time, each with its own head/tail coin; after `max_try` rejected draws the tail is
    replaced by a uniform entity."""
    neighbor = neighbor or {}
    neg_batch = []
    for head, relation, tail in pos_batch:
        head_pool = neighbor.get(head, entities_list)
        tail_pool = neighbor.get(tail, entities_list)
        for _ in range(neg_triples_num):
            for attempt in range(1, max_try + 1):
                if np.random.binomial(1, 0.5):
                    cand = (random.choice(head_pool), relation, tail)
                else:
                    cand = (head, relation, random.choice(tail_pool))
                if cand not in all_triples_set:
                    neg_batch.append(cand)
                    break
                if attempt == max_try:
                    neg_batch.append((head, relation, random.choice(entities_list)))
    assert len(neg_batch) == neg_triples_num * len(pos_batch)
    return neg_batch


def generate_relation_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2,
                                   entity_list1, entity_list2, batch_size,
                                   step, neighbor1, neighbor2, neg_triples_num):
    size1, size2 = _split_sizes(len(triple_list1), len(triple_list2), batch_size)
    pos1 = generate_pos_triples(triple_list1, size1, step)
    pos2 = generate_pos_triples(triple_list2, size2, step)
    neg1 = generate_neg_triples_fast(pos1, triple_set1, entity_list1, neg_triples_num, neighbor=neighbor1)
    neg2 = generate_neg_triples_fast(pos2, triple_set2, entity_list2, neg_triples_num, neighbor=neighbor2)
    return pos1 + pos2, neg1 + neg2


def generate_relation_triple_batch_queue(triple_list1, triple_list2, triple_set1, triple_set2,
                                         entity_list1, entity_list2, batch_size,
                                         steps, out_queue, neighbor1, neighbor2, neg_triples_num):
    for step in steps:
        out_queue.put(generate_relation_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2,
                                                     entity_list1, entity_list2, batch_size,
                                                     step, neighbor1, neighbor2, neg_triples_num))


def generate_triple_label_batch(triple_list1, triple_list2, triple_set1, triple_set2, entity_list1, entity_list2,
                                batch_size, steps, out_queue, neighbor1, neighbor2, neg_triples_num):
    size1, size2 = _split_sizes(len(triple_list1), len(triple_list2), batch_size)
    for step in steps:
        pos1 = generate_pos_triples(triple_list1, size1, step)
        pos2 = generate_pos_triples(triple_list2, size2, step)
        neg = (generate_neg_triples(pos1, triple_set1, entity_list1, neg_triples_num, neighbor=neighbor1)
               + generate_neg_triples(pos2, triple_set2, entity_list2, neg_triples_num, neighbor=neighbor2))
        pos = pos1 + pos2
        out_queue.put((pos + neg, [1] * len(pos) + [-1] * len(neg)))


# ---- attribute triples -------------------------------------------------------------------------------
def generate_neg_attribute_triples(pos_batch, all_triples_set, entity_list, neg_triples_num, neighbor=None):
    neighbor = neighbor or {}
    neg_batch = []
    for head, attribute, value in pos_batch:
        pool = neighbor.get(head, entity_list)
        for _ in range(neg_triples_num):
            neg_head = random.choice(pool)
            while (neg_head, attribute, value) in all_triples_set:
                neg_head = random.choice(pool)
            neg_batch.append((neg_head, attribute, value))
    assert len(neg_batch) == neg_triples_num * len(pos_batch)
    return neg_batch


def generate_attribute_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2,
                                    entity_list1, entity_list2, batch_size,
                                    step, neighbor1, neighbor2, neg_triples_num, is_fixed_size):
    size1, size2 = _split_sizes(len(triple_list1), len(triple_list2), batch_size)
    pos1 = generate_pos_triples(triple_list1, size1, step, is_fixed_size=is_fixed_size)
    pos2 = generate_pos_triples(triple_list2, size2, step, is_fixed_size=is_fixed_size)
    neg1 = generate_neg_attribute_triples(pos1, triple_set1, entity_list1, neg_triples_num, neighbor=neighbor1)
    neg2 = generate_neg_attribute_triples(pos2, triple_set2, entity_list2, neg_triples_num, neighbor=neighbor2)
    return pos1 + pos2, neg1 + neg2


def generate_attribute_triple_batch_queue(triple_list1, triple_list2, triple_set1, triple_set2,
                                          entity_list1, entity_list2, batch_size,
                                          steps, out_queue, neighbor1, neighbor2, neg_triples_num, is_fixed_size):
    for step in steps:
        out_queue.put(generate_attribute_triple_batch(triple_list1, triple_list2, triple_set1, triple_set2,
                                                      entity_list1, entity_list2, batch_size, step,
                                                      neighbor1, neighbor2, neg_triples_num, is_fixed_size))


# ---- ε-truncated neighbour search (GPU) ----------------------------------------------------------------
def neighbours_device(entity_embeds, entity_list, neighbors_num):
    """[n, k] int32 CUDA tensor of the k most similar entity ids per row (the device form the fused sampler
    consumes; see DeviceKG.set_candidates)."""
    from openea_b200 import finding
    return finding.find_neighbours_device(entity_embeds, entity_list, neighbors_num)


def find_neighbours(frags, entity_list, sub_embed, embed, k):
    """dict frag-entity → list of its k nearest entities (inner product), computed on the GPU."""
    import torch
    from openea_b200 import finding, lib as L
    import ctypes as C
    from openea_b200.engine import _ptr, _stream_ptr
    sub, d = finding.to_device_rows(sub_embed)
    full, _ = finding.to_device_rows(embed)
    n_sub, n = sub.shape[0], full.shape[0]
    ids = torch.as_tensor(np.asarray(entity_list, dtype=np.int32), device=sub.device)
    ld = (n + 3) // 4 * 4
    buf = torch.empty(n_sub, ld, dtype=torch.float32, device=sub.device)
    out = torch.empty(n_sub, k, dtype=torch.int32, device=sub.device)
    lib = L.load()
    cfg = finding._cfg("inner", sub, full, d)
    L.check(lib.oea_sim_matrix(C.byref(cfg), _ptr(sub), _ptr(full), None, None, _ptr(buf), ld, _stream_ptr()), "oea_sim_matrix")
    L.check(lib.oea_rows_select_topk(_ptr(buf), ld, n_sub, n, k, _ptr(ids), _ptr(out), _stream_ptr()), "oea_rows_select_topk")
    host = out.cpu().numpy()
    return {frags[i]: host[i].tolist() for i in range(n_sub)}


def generate_neighbours_single_thread(entity_embeds, entity_list, neighbors_num, threads_num):
    cand = neighbours_device(entity_embeds, entity_list, neighbors_num).cpu().numpy()
    return {int(e): cand[i].tolist() for i, e in enumerate(entity_list)}


def generate_neighbours(entity_embeds, entity_list, neighbors_num, threads_num):
    return generate_neighbours_single_thread(entity_embeds, entity_list, neighbors_num, threads_num)
