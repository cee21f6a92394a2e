"""KG container exposing the attribute names the reference's approaches read (modules/load/kg.py:10-141).

The attribute surface is the contract (`<kind>_triples_set/_list/_num`, `local_<kind>_…`, `entities_*`,
`rt_dict / hr_dict / av_dict`, `entity_relations_dict / entity_attributes_dict`, `sup_<kind>_triples_*`);
everything here is derived from the two triple sets by `_install`, one family of names per call.
"""


def parse_triples(triples):
    """(first column, second column, third column) as three sets."""
    cols = (set(), set(), set())
    for triple in triples:
        for col, item in zip(cols, triple):
            col.add(item)
    return cols


def _group(pairs):
    """{key: set(values)} from an iterable of (key, value)."""
    out = {}
    for key, val in pairs:
        out.setdefault(key, set()).add(val)
    return out


class KG:
    _STATS = (("entities", "entities_num"), ("relations", "relations_num"), ("attributes", "attributes_num"),
              ("relation triples", "relation_triples_num"), ("attribute triples", "attribute_triples_num"),
              ("local relation triples", "local_relation_triples_num"),
              ("local attribute triples", "local_attribute_triples_num"))

    def __init__(self, relation_triples, attribute_triples):
        for name in ("entities_id_dict", "relations_id_dict", "attributes_id_dict"):
            setattr(self, name, None)
        for kind in ("relation", "attribute"):
            setattr(self, "sup_%s_triples_set" % kind, None)
            setattr(self, "sup_%s_triples_list" % kind, None)
        self.entities_set = set()
        self.set_relations(relation_triples)
        self.set_attributes(attribute_triples)
        print("\nKG statistics:")
        for label, attr in self._STATS:
            print("Number of %s:" % label, getattr(self, attr))
        print()

    # -- shared plumbing ----------------------------------------------------------------------------------
    def _install(self, kind, triples):
        """Set `<kind>_triples_{set,list,num}` and their `local_` twins (same objects, as in the reference)."""
        tset = set(triples)
        tlist = list(tset)
        for prefix in ("", "local_"):
            setattr(self, "%s%s_triples_set" % (prefix, kind), tset)
            setattr(self, "%s%s_triples_list" % (prefix, kind), tlist)
            setattr(self, "%s%s_triples_num" % (prefix, kind), len(tset))
        return tset

    def _set_vocab(self, name, items):
        setattr(self, name + "_set", items)
        setattr(self, name + "_list", list(items))
        setattr(self, name + "_num", len(items))

    def _add_sup(self, kind, sup_triples):
        sup = set(sup_triples)
        setattr(self, "sup_%s_triples_set" % kind, sup)
        setattr(self, "sup_%s_triples_list" % kind, list(sup))
        merged = getattr(self, "%s_triples_set" % kind)
        merged |= sup_triples
        setattr(self, "%s_triples_list" % kind, list(merged))
        setattr(self, "%s_triples_num" % kind, len(merged))

    # -- reference API -------------------------------------------------------------------------------------
    def set_relations(self, relation_triples):
        heads, relations, tails = parse_triples(self._install("relation", relation_triples))
        self._set_vocab("entities", heads | tails)
        self._set_vocab("relations", relations)
        self.generate_relation_triple_dict()
        self.parse_relations()

    def set_attributes(self, attribute_triples):
        subjects, attributes, _ = parse_triples(self._install("attribute", attribute_triples))
        self._set_vocab("attributes", attributes)
        self._set_vocab("entities", self.entities_set | subjects)   # entities that only carry attribute triples
        self.generate_attribute_triple_dict()
        self.parse_attributes()

    def generate_relation_triple_dict(self):
        triples = self.local_relation_triples_list
        self.rt_dict = _group((h, (r, t)) for h, r, t in triples)
        self.hr_dict = _group((t, (h, r)) for h, r, t in triples)
        print("Number of rt_dict:", len(self.rt_dict))
        print("Number of hr_dict:", len(self.hr_dict))

    def generate_attribute_triple_dict(self):
        self.av_dict = _group((e, (a, v)) for e, a, v in self.local_attribute_triples_list)
        print("Number of av_dict:", len(self.av_dict))

    def parse_relations(self):
        self.entity_relations_dict = _group(t[:2] for t in self.local_relation_triples_set)
        print("entity relations dict:", len(self.entity_relations_dict))

    def parse_attributes(self):
        self.entity_attributes_dict = _group(t[:2] for t in self.local_attribute_triples_set)
        print("entity attributes dict:", len(self.entity_attributes_dict))

    def set_id_dict(self, entities_id_dict, relations_id_dict, attributes_id_dict):
        self.entities_id_dict, self.relations_id_dict, self.attributes_id_dict = \
            entities_id_dict, relations_id_dict, attributes_id_dict

    def add_sup_relation_triples(self, sup_triples):
        self._add_sup("relation", sup_triples)

    def add_sup_attribute_triples(self, sup_triples):
        self._add_sup("attribute", sup_triples)
