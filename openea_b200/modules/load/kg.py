"""KG container with the attribute names the reference's approaches read (modules/load/kg.py:10-141)."""


def parse_triples(triples):
    subjects, predicates, objects = set(), set(), set()
    for s, p, o in triples:
        subjects.add(s)
        predicates.add(p)
        objects.add(o)
    return subjects, predicates, objects


def _group(pairs):
    """{key: set(values)} from an iterable of (key, value)."""
    out = {}
    for key, val in pairs:
        out.setdefault(key, set()).add(val)
    return out


class KG:
    def __init__(self, relation_triples, attribute_triples):
        self.entities_id_dict = self.relations_id_dict = self.attributes_id_dict = None
        self.sup_relation_triples_set = self.sup_relation_triples_list = None
        self.sup_attribute_triples_set = self.sup_attribute_triples_list = None
        self.set_relations(relation_triples)
        self.set_attributes(attribute_triples)
        print()
        print("KG statistics:")
        print("Number of entities:", self.entities_num)
        print("Number of relations:", self.relations_num)
        print("Number of attributes:", self.attributes_num)
        print("Number of relation triples:", self.relation_triples_num)
        print("Number of attribute triples:", self.attribute_triples_num)
        print("Number of local relation triples:", self.local_relation_triples_num)
        print("Number of local attribute triples:", self.local_attribute_triples_num)
        print()

    def set_relations(self, relation_triples):
        self.relation_triples_set = set(relation_triples)
        self.relation_triples_list = list(self.relation_triples_set)
        self.local_relation_triples_set = self.relation_triples_set
        self.local_relation_triples_list = self.relation_triples_list
        heads, relations, tails = parse_triples(self.relation_triples_set)
        self.entities_set = heads | tails
        self.relations_set = relations
        self.entities_list = list(self.entities_set)
        self.relations_list = list(self.relations_set)
        self.entities_num = len(self.entities_set)
        self.relations_num = len(self.relations_set)
        self.relation_triples_num = len(self.relation_triples_set)
        self.local_relation_triples_num = len(self.local_relation_triples_set)
        self.generate_relation_triple_dict()
        self.parse_relations()

    def set_attributes(self, attribute_triples):
        self.attribute_triples_set = set(attribute_triples)
        self.attribute_triples_list = list(self.attribute_triples_set)
        self.local_attribute_triples_set = self.attribute_triples_set
        self.local_attribute_triples_list = self.attribute_triples_list
        entities, attributes, _ = parse_triples(self.attribute_triples_set)
        self.attributes_set = attributes
        self.attributes_list = list(attributes)
        self.attributes_num = len(attributes)
        self.entities_set |= entities            # entities that only have attribute triples
        self.entities_list = list(self.entities_set)
        self.entities_num = len(self.entities_set)
        self.attribute_triples_num = len(self.attribute_triples_set)
        self.local_attribute_triples_num = len(self.local_attribute_triples_set)
        self.generate_attribute_triple_dict()
        self.parse_attributes()

    def generate_relation_triple_dict(self):
        self.rt_dict = _group((h, (r, t)) for h, r, t in self.local_relation_triples_list)
        self.hr_dict = _group((t, (h, r)) for h, r, t in self.local_relation_triples_list)
        print("Number of rt_dict:", len(self.rt_dict))
        print("Number of hr_dict:", len(self.hr_dict))

    def generate_attribute_triple_dict(self):
        self.av_dict = _group((h, (a, v)) for h, a, v in self.local_attribute_triples_list)
        print("Number of av_dict:", len(self.av_dict))

    def parse_relations(self):
        self.entity_relations_dict = _group((h, r) for h, r, _ in self.local_relation_triples_set)
        print("entity relations dict:", len(self.entity_relations_dict))

    def parse_attributes(self):
        self.entity_attributes_dict = _group((h, a) for h, a, _ in self.local_attribute_triples_set)
        print("entity attributes dict:", len(self.entity_attributes_dict))

    def set_id_dict(self, entities_id_dict, relations_id_dict, attributes_id_dict):
        self.entities_id_dict = entities_id_dict
        self.relations_id_dict = relations_id_dict
        self.attributes_id_dict = attributes_id_dict

    def add_sup_relation_triples(self, sup_triples):
        self.sup_relation_triples_set = set(sup_triples)
        self.sup_relation_triples_list = list(self.sup_relation_triples_set)
        self.relation_triples_set |= sup_triples
        self.relation_triples_list = list(self.relation_triples_set)
        self.relation_triples_num = len(self.relation_triples_list)

    def add_sup_attribute_triples(self, sup_triples):
        self.sup_attribute_triples_set = set(sup_triples)
        self.sup_attribute_triples_list = list(self.sup_attribute_triples_set)
        self.attribute_triples_set |= sup_triples
        self.attribute_triples_list = list(self.attribute_triples_set)
        self.attribute_triples_num = len(self.attribute_triples_list)
