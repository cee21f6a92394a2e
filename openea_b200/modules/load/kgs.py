"""Pair of KGs in id space + the dataset readers (interface of modules/load/kgs.py:5-99 of the reference)."""
import os

from openea_b200.modules.load.kg import KG
from openea_b200.modules.load.read import *  # noqa: F401,F403  (the reference re-exports read.* here)
from openea_b200.modules.load import read as rd


class KGs:
    def __init__(self, kg1: KG, kg2: KG, train_links, test_links, valid_links=None, mode='mapping', ordered=True):
        id_fn = rd.generate_sharing_id if mode == "sharing" else None

        def make_ids(t1, e1, t2, e2, links):
            if id_fn is not None:
                return id_fn(links, t1, e1, t2, e2, ordered=ordered)
            return rd.generate_mapping_id(t1, e1, t2, e2, ordered=ordered)
        ent_ids1, ent_ids2 = make_ids(kg1.relation_triples_set, kg1.entities_set,
                                      kg2.relation_triples_set, kg2.entities_set, train_links)
        rel_ids1, rel_ids2 = make_ids(kg1.relation_triples_set, kg1.relations_set,
                                      kg2.relation_triples_set, kg2.relations_set, [])
        attr_ids1, attr_ids2 = make_ids(kg1.attribute_triples_set, kg1.attributes_set,
                                        kg2.attribute_triples_set, kg2.attributes_set, [])
        self.uri_kg1, self.uri_kg2 = kg1, kg2
        id_kg1 = KG(rd.uris_relation_triple_2ids(kg1.relation_triples_set, ent_ids1, rel_ids1),
                    rd.uris_attribute_triple_2ids(kg1.attribute_triples_set, ent_ids1, attr_ids1))
        id_kg2 = KG(rd.uris_relation_triple_2ids(kg2.relation_triples_set, ent_ids2, rel_ids2),
                    rd.uris_attribute_triple_2ids(kg2.attribute_triples_set, ent_ids2, attr_ids2))
        id_kg1.set_id_dict(ent_ids1, rel_ids1, attr_ids1)
        id_kg2.set_id_dict(ent_ids2, rel_ids2, attr_ids2)

        self.uri_train_links, self.uri_test_links = train_links, test_links
        self.train_links = rd.uris_pair_2ids(train_links, ent_ids1, ent_ids2)
        self.test_links = rd.uris_pair_2ids(test_links, ent_ids1, ent_ids2)
        self.train_entities1 = [a for a, _ in self.train_links]
        self.train_entities2 = [b for _, b in self.train_links]
        self.test_entities1 = [a for a, _ in self.test_links]
        self.test_entities2 = [b for _, b in self.test_links]

        if mode == 'swapping':
            sup1, sup2 = rd.generate_sup_relation_triples(self.train_links, id_kg1.rt_dict, id_kg1.hr_dict,
                                                          id_kg2.rt_dict, id_kg2.hr_dict)
            id_kg1.add_sup_relation_triples(sup1)
            id_kg2.add_sup_relation_triples(sup2)
            sup1, sup2 = rd.generate_sup_attribute_triples(self.train_links, id_kg1.av_dict, id_kg2.av_dict)
            id_kg1.add_sup_attribute_triples(sup1)
            id_kg2.add_sup_attribute_triples(sup2)
        self.kg1, self.kg2 = id_kg1, id_kg2

        self.valid_links, self.valid_entities1, self.valid_entities2 = [], [], []
        if valid_links is not None:
            self.uri_valid_links = valid_links
            self.valid_links = rd.uris_pair_2ids(valid_links, ent_ids1, ent_ids2)
            self.valid_entities1 = [a for a, _ in self.valid_links]
            self.valid_entities2 = [b for _, b in self.valid_links]

        self.useful_entities_list1 = self.kg1.entities_list
        self.useful_entities_list2 = self.kg2.entities_list
        self.entities_num = len(self.kg1.entities_set | self.kg2.entities_set)
        self.relations_num = len(self.kg1.relations_set | self.kg2.relations_set)
        self.attributes_num = len(self.kg1.attributes_set | self.kg2.attributes_set)


def _load_folder(folder, division, swap_sides):
    a, b = ("2", "1") if swap_sides else ("1", "2")
    rel1, _, _ = rd.read_relation_triples(folder + 'rel_triples_' + a)
    rel2, _, _ = rd.read_relation_triples(folder + 'rel_triples_' + b)
    attr1, _, _ = rd.read_attribute_triples(folder + 'attr_triples_' + a)
    attr2, _, _ = rd.read_attribute_triples(folder + 'attr_triples_' + b)
    links = [rd.read_links(folder + division + name) for name in ('train_links', 'valid_links', 'test_links')]
    if swap_sides:
        links = [[(j, i) for i, j in part] for part in links]
    return rel1, rel2, attr1, attr2, links


def _build(rel1, rel2, attr1, attr2, links, mode, ordered, remove_unlinked):
    train_links, valid_links, test_links = links
    if remove_unlinked:
        every = train_links + valid_links + test_links
        rel1 = remove_unlinked_triples(rel1, every)
        rel2 = remove_unlinked_triples(rel2, every)
    return KGs(KG(rel1, attr1), KG(rel2, attr2), train_links, test_links, valid_links=valid_links, mode=mode,
               ordered=ordered)


def read_kgs_from_folder(training_data_folder, division, mode, ordered, remove_unlinked=False):
    """kgs.py:102-118 of the reference.  Datasets in the standard layout load through the array-backed layer
    (modules/load/fast.py: C parser, vectorised id assignment, lazy Python containers, binary cache); what that layer
    does not cover, or OEA_LOADER=containers, uses the container-based loader below."""
    lowered = training_data_folder.lower()
    if 'dbp15k' in lowered or 'dwy100k' in lowered:
        return read_kgs_from_dbp_dwy(training_data_folder, division, mode, ordered, remove_unlinked=remove_unlinked)
    if os.environ.get("OEA_LOADER") != "containers":
        from openea_b200.modules.load import fast
        try:
            return fast.load(training_data_folder, division, mode, ordered, remove_unlinked=remove_unlinked)
        except fast.Unsupported as why:
            print("array-backed loader not applicable (%s); using the container-based loader" % why)
    return _build(*_load_folder(training_data_folder, division, False), mode, ordered, remove_unlinked)


def read_reversed_kgs_from_folder(training_data_folder, division, mode, ordered, remove_unlinked=False):
    return _build(*_load_folder(training_data_folder, division, True), mode, ordered, remove_unlinked)


def read_kgs_from_files(kg1_relation_triples, kg2_relation_triples, kg1_attribute_triples, kg2_attribute_triples,
                        train_links, valid_links, test_links, mode):
    return KGs(KG(kg1_relation_triples, kg1_attribute_triples), KG(kg2_relation_triples, kg2_attribute_triples),
               train_links, test_links, valid_links=valid_links, mode=mode)


def read_kgs_from_dbp_dwy(folder, division, mode, ordered, remove_unlinked=False):
    folder = folder + division
    rel1, _, _ = rd.read_relation_triples(folder + 'triples_1')
    rel2, _, _ = rd.read_relation_triples(folder + 'triples_2')
    train_links = rd.read_links(folder + ('sup_pairs' if os.path.exists(folder + 'sup_pairs') else 'sup_ent_ids'))
    test_links = rd.read_links(folder + ('ref_pairs' if os.path.exists(folder + 'ref_pairs') else 'ref_ent_ids'))
    print()
    if remove_unlinked:
        for i in range(10000):          # iterate link/triple pruning to a fixed point
            print("removing times:", i)
            rel1 = remove_unlinked_triples(rel1, train_links + test_links)
            rel2 = remove_unlinked_triples(rel2, train_links + test_links)
            before = (len(rel1), len(rel2))
            train_links, test_links = remove_no_triples_link(rel1, rel2, train_links, test_links)
            rel1 = remove_unlinked_triples(rel1, train_links + test_links)
            rel2 = remove_unlinked_triples(rel2, train_links + test_links)
            if before == (len(rel1), len(rel2)):
                break
            print()
    return KGs(KG(rel1, list()), KG(rel2, list()), train_links, test_links, mode=mode, ordered=ordered)


def remove_no_triples_link(kg1_relation_triples, kg2_relation_triples, train_links, test_links):
    ents1 = {e for h, _, t in kg1_relation_triples for e in (h, t)}
    ents2 = {e for h, _, t in kg2_relation_triples for e in (h, t)}
    print("before removing links with no triples:", len(train_links), len(test_links))
    keep = lambda links: list({(i, j) for i, j in links if i in ents1 and j in ents2})
    new_train, new_test = keep(train_links), keep(test_links)
    print("after removing links with no triples:", len(new_train), len(new_test))
    return new_train, new_test


def remove_unlinked_triples(triples, links):
    print("before removing unlinked triples:", len(triples))
    linked = {e for pair in links for e in pair}
    kept = {(h, r, t) for h, r, t in triples if h in linked and t in linked}
    print("after removing unlinked triples:", len(kept))
    return kept
