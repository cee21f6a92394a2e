"""TSV readers, id assignment and result writers.

Behavioural mirror of the reference's modules/load/read.py (id layout matters for parity: ordered ids
interleave the two KGs by descending frequency, read.py:64-92).  Written from the behaviour, not the code.
"""
import os
from collections import Counter

import numpy as np


# ---- id assignment ---------------------------------------------------------------------------------
def sort_elements(triples, elements_set):
    """Order `elements_set` by (occurrence count in `triples`, name) descending (read.py:12-29).
    Returns (ordered list, count dict); elements that never occur count 0."""
    counts = Counter()
    for s, p, o in triples:
        if s in elements_set:
            counts[s] += 1
        if p in elements_set:
            counts[p] += 1
        if o in elements_set:
            counts[o] += 1
    freq = {e: counts.get(e, 0) for e in elements_set}
    ordered = sorted(freq, key=lambda e: (freq[e], e), reverse=True)
    assert len(freq) == len(elements_set)
    return ordered, freq


def generate_mapping_id(kg1_triples, kg1_elements, kg2_triples, kg2_elements, ordered=True):
    """Disjoint ids for the two KGs.  ordered: rank i of KG1 → 2i, rank i of KG2 → 2i+1 while both have a
    rank-i element; the longer list's overflow continues contiguously after 2·min(n1,n2) (read.py:64-92)."""
    ids1, ids2 = {}, {}
    if ordered:
        order1, _ = sort_elements(kg1_triples, kg1_elements)
        order2, _ = sort_elements(kg2_triples, kg2_elements)
        n1, n2 = len(order1), len(order2)
        both = min(n1, n2)
        for rank in range(both):
            ids1[order1[rank]] = 2 * rank
            ids2[order2[rank]] = 2 * rank + 1
        for rank in range(both, n1):
            ids1[order1[rank]] = 2 * n2 + (rank - n2)
        for rank in range(both, n2):
            ids2[order2[rank]] = 2 * n1 + (rank - n1)
    else:
        nxt = 0
        for ele in kg1_elements:
            if ele not in ids1:
                ids1[ele] = nxt
                nxt += 1
        for ele in kg2_elements:
            if ele not in ids2:
                ids2[ele] = nxt
                nxt += 1
    assert len(ids1) == len(set(kg1_elements))
    assert len(ids2) == len(set(kg2_elements))
    return ids1, ids2


def generate_sharing_id(train_links, kg1_triples, kg1_elements, kg2_triples, kg2_elements, ordered=True):
    """Seed-linked KG2 elements reuse their KG1 counterpart's id (read.py:32-61)."""
    ids1, ids2 = {}, {}
    if ordered:
        counterpart = {y: x for x, y in train_links}
        linked2 = [y for _, y in train_links]
        unlinked2 = set(kg2_elements) - set(linked2)
        ids1, ids2 = generate_mapping_id(kg1_triples, kg1_elements, kg2_triples, unlinked2, ordered=True)
        for ele in linked2:
            ids2[ele] = ids1[counterpart[ele]]
    else:
        nxt = 0
        for e1, e2 in train_links:
            assert e1 in kg1_elements and e2 in kg2_elements
            ids1[e1] = ids2[e2] = nxt
            nxt += 1
        for ele in kg1_elements:
            if ele not in ids1:
                ids1[ele] = nxt
                nxt += 1
        for ele in kg2_elements:
            if ele not in ids2:
                ids2[ele] = nxt
                nxt += 1
    assert len(ids1) == len(set(kg1_elements))
    assert len(ids2) == len(set(kg2_elements))
    return ids1, ids2


# ---- uri → id conversion -----------------------------------------------------------------------------
def uris_list_2ids(uris, ids):
    out = [ids[u] for u in uris]
    assert len(out) == len(set(uris))
    return out


def uris_pair_2ids(uris, ids1, ids2):
    return [(ids1[a], ids2[b]) for a, b in uris if a in ids1 and b in ids2]


def uris_relation_triple_2ids(uris, ent_ids, rel_ids):
    out = [(ent_ids[h], rel_ids[r], ent_ids[t]) for h, r, t in uris]
    assert len(out) == len(set(uris))
    return out


def uris_attribute_triple_2ids(uris, ent_ids, attr_ids):
    out = [(ent_ids[h], attr_ids[a], v) for h, a, v in uris]
    assert len(out) == len(set(uris))
    return out


# ---- swap ("supervised") triples for alignment_module == 'swapping' -------------------------------------
def generate_sup_relation_triples_one_link(e1, e2, rt_dict, hr_dict):
    swapped = {(e2, r, t) for r, t in rt_dict.get(e1, ())}
    swapped.update((h, r, e2) for h, r in hr_dict.get(e1, ()))
    return swapped


def generate_sup_relation_triples(sup_links, rt_dict1, hr_dict1, rt_dict2, hr_dict2):
    new1, new2 = set(), set()
    for ent1, ent2 in sup_links:
        new1 |= generate_sup_relation_triples_one_link(ent1, ent2, rt_dict1, hr_dict1)
        new2 |= generate_sup_relation_triples_one_link(ent2, ent1, rt_dict2, hr_dict2)
    print("supervised relation triples: {}, {}".format(len(new1), len(new2)))
    return new1, new2


def generate_sup_attribute_triples_one_link(e1, e2, av_dict):
    return {(e2, a, v) for a, v in av_dict.get(e1, ())}


def generate_sup_attribute_triples(sup_links, av_dict1, av_dict2):
    new1, new2 = set(), set()
    for ent1, ent2 in sup_links:
        new1 |= generate_sup_attribute_triples_one_link(ent1, ent2, av_dict1)
        new2 |= generate_sup_attribute_triples_one_link(ent2, ent1, av_dict2)
    print("supervised attribute triples: {}, {}".format(len(new1), len(new2)))
    return new1, new2


# ---- file readers ------------------------------------------------------------------------------------
def _fields(line):
    return line.rstrip("\n").split("\t")


def read_relation_triples(file_path):
    print("read relation triples:", file_path)
    triples, entities, relations = set(), set(), set()
    if file_path is None:
        return triples, entities, relations
    with open(file_path, "r", encoding="utf8") as fh:
        for line in fh:
            parts = _fields(line)
            assert len(parts) == 3
            h, r, t = (p.strip() for p in parts)
            triples.add((h, r, t))
            entities.update((h, t))
            relations.add(r)
    return triples, entities, relations


def read_attribute_triples(file_path):
    print("read attribute triples:", file_path)
    triples, entities, attributes = set(), set(), set()
    if file_path is None:
        return triples, entities, attributes
    with open(file_path, "r", encoding="utf8") as fh:
        for line in fh:
            parts = line.strip().strip("\n").split("\t")
            if len(parts) < 3:
                continue
            head, attr = parts[0].strip(), parts[1].strip()
            value = " ".join(p.strip() for p in parts[2:])      # extra tab-separated pieces join the value
            value = value.strip().rstrip(".").strip()
            entities.add(head)
            attributes.add(attr)
            triples.add((head, attr, value))
    return triples, entities, attributes


def read_links(file_path):
    print("read links:", file_path)
    links = []
    with open(file_path, "r", encoding="utf8") as fh:
        for line in fh:
            parts = _fields(line)
            assert len(parts) == 2
            links.append((parts[0].strip(), parts[1].strip()))
    return links


def read_dict(file_path):
    ids = {}
    with open(file_path, "r", encoding="utf8") as fh:
        for line in fh:
            parts = _fields(line)
            assert len(parts) == 2
            ids[parts[0]] = int(parts[1])
    return ids


def read_pair_ids(file_path):
    pairs = []
    with open(file_path, "r", encoding="utf8") as fh:
        for line in fh:
            parts = _fields(line)
            assert len(parts) == 2
            pairs.append((int(parts[0]), int(parts[1])))
    return pairs


# ---- writers -----------------------------------------------------------------------------------------
def pair2file(file, pairs):
    if pairs is None:
        return
    with open(file, "w", encoding="utf8") as fh:
        fh.writelines("%s\t%s\n" % (i, j) for i, j in pairs)


def dict2file(file, dic):
    if dic is None:
        return
    with open(file, "w", encoding="utf8") as fh:
        fh.writelines("%s\t%s\n" % (k, v) for k, v in dic.items())
    print(file, "saved.")


def line2file(file, lines):
    if lines is None:
        return
    with open(file, "w", encoding="utf8") as fh:
        fh.writelines(line + "\n" for line in lines)
    print(file, "saved.")


def radio_2file(radio, folder):
    path = folder + str(radio).replace(".", "_")
    os.makedirs(path, exist_ok=True)
    return path + "/"


def load_embeddings(file_name):
    return np.load(file_name) if os.path.exists(file_name) else None


def save_results(folder, rest_12):
    os.makedirs(folder, exist_ok=True)
    pair2file(folder + "alignment_results_12", rest_12)
    print("Results saved!")


def embed2file(results_folder, file_name, embedding, kg1_id_dict, kg2_id_dict, seperate=True):
    if embedding is None or kg1_id_dict is None or kg2_id_dict is None:
        return

    def dump(fh, id_dict):
        for uri, index in id_dict.items():
            fh.write(str(uri) + " " + " ".join(map(str, embedding[index])) + "\n")
    if seperate:
        with open(results_folder + "kg1_" + file_name, "w", encoding="utf8") as fh:
            dump(fh, kg1_id_dict)
        with open(results_folder + "kg2_" + file_name, "w", encoding="utf8") as fh:
            dump(fh, kg2_id_dict)
    else:
        with open(results_folder + "combined_" + file_name, "w", encoding="utf8") as fh:
            dump(fh, kg1_id_dict)
            dump(fh, kg2_id_dict)


def save_embeddings(folder, kgs, ent_embeds, rel_embeds, attr_embeds, mapping_mat=None, rev_mapping_mat=None):
    """Files of read.py:325-349: *.npy, kg{1,2}_{ent,rel,attr}_ids, kg{1,2}_*_embeds_txt."""
    os.makedirs(folder, exist_ok=True)
    for name, arr in (("ent_embeds", ent_embeds), ("rel_embeds", rel_embeds), ("attr_embeds", attr_embeds),
                      ("mapping_mat", mapping_mat), ("rev_mapping_mat", rev_mapping_mat)):
        if arr is not None:
            np.save(folder + name + ".npy", arr)
    dict2file(folder + "kg1_ent_ids", kgs.kg1.entities_id_dict)
    dict2file(folder + "kg2_ent_ids", kgs.kg2.entities_id_dict)
    dict2file(folder + "kg1_rel_ids", kgs.kg1.relations_id_dict)
    dict2file(folder + "kg2_rel_ids", kgs.kg2.relations_id_dict)
    dict2file(folder + "kg1_attr_ids", kgs.kg1.attributes_id_dict)
    dict2file(folder + "kg2_attr_ids", kgs.kg2.attributes_id_dict)
    embed2file(folder, "ent_embeds_txt", ent_embeds, kgs.kg1.entities_id_dict, kgs.kg2.entities_id_dict)
    embed2file(folder, "rel_embeds_txt", rel_embeds, kgs.kg1.relations_id_dict, kgs.kg2.relations_id_dict)
    embed2file(folder, "attr_embeds_txt", attr_embeds, kgs.kg1.attributes_id_dict, kgs.kg2.attributes_id_dict)
    print("Embeddings saved!")
