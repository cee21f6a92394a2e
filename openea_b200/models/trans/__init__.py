from openea_b200.models.trans.transe import TransE
from openea_b200.models._stubs import out_of_scope

TransH = out_of_scope("TransH", "hyperplane projection score: listed as a 'next' variant of kernel K1 (SURVEY §8f)")
TransR = out_of_scope("TransR", "per-relation projection matrices, not in BASELINE configs")
TransD = out_of_scope("TransD", "dynamic projection score: listed as a 'next' variant of kernel K1 (SURVEY §8f)")
