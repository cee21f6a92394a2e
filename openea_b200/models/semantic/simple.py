"""SimplE (models/semantic/simple.py of the reference): head- and tail-role entity tables, two relation tables, score
(Σ l2n(h_H∘r₁)∘t_T + Σ l2n(t_H∘r₂)∘h_T)/2, loss Σ softplus(−score⁺) + Σ softplus(score⁻) — the OEA_MODEL_SIMPLE
instance of oea_model_score_fed.  Evaluation and save use head + tail embeddings (simple.py:88-115)."""
import torch

import openea_b200.modules.load.read as rd
from openea_b200 import engine as eng
from openea_b200 import finding as F
from openea_b200.models.basic_model import BasicModel
from openea_b200.modules.base.initializers import init_embeddings
from openea_b200.modules.utils.util import load_session


class SimplE(BasicModel):

    def __init__(self):
        super().__init__()
        self.head_ent_embeds = self.tail_ent_embeds = self.rel_embeds1 = self.rel_embeds2 = None

    def init(self):
        self._define_variables()
        self._define_embed_graph()
        self.session = load_session()
        a = self.args
        assert a.init == 'xavier'
        assert a.alignment_module == 'sharing'
        assert a.neg_sampling == 'uniform'
        assert a.optimizer == 'Adagrad'
        assert a.eval_metric == 'inner'
        assert a.ent_l2_norm is True
        assert a.rel_l2_norm is True

    def _define_variables(self):
        a = self.args
        mk = lambda rows, name, norm: init_embeddings([rows, a.dim], name, a.init, norm, optimizer=a.optimizer)
        self.head_ent_embeds = mk(self.kgs.entities_num, 'head_ent_embeds', a.ent_l2_norm)
        self.tail_ent_embeds = mk(self.kgs.entities_num, 'tail_ent_embeds', a.ent_l2_norm)
        self.rel_embeds1 = mk(self.kgs.relations_num, 'rel_embeds1', a.rel_l2_norm)
        self.rel_embeds2 = mk(self.kgs.relations_num, 'rel_embeds2', a.rel_l2_norm)
        self.ent_embeds = self.head_ent_embeds        # device / shape queries of the base class

    def _define_embed_graph(self):
        self.triple_trainer = eng.ModelTrainer("SimplE", (self.head_ent_embeds, self.rel_embeds1, self.tail_ent_embeds,
                                                          self.rel_embeds2),
                                               eng.loss_cfg("logistic", "L2"), self.args.learning_rate)
        self.neg_per_pos = self.args.neg_triple_num
        self.triple_loss = self.triple_optimizer = self.triple_trainer

    def _both_roles(self, ids=None):
        return self.head_ent_embeds.lookup(ids) + self.tail_ent_embeds.lookup(ids)   # device row add (plumbing)

    def _eval_valid_embeddings(self):
        print("valid")
        return (self._both_roles(self.kgs.valid_entities1),
                self._both_roles(self.kgs.valid_entities2 + self.kgs.test_entities2), self._mapping_array())

    def _eval_test_embeddings(self):
        print("test")
        return self._both_roles(self.kgs.test_entities1), self._both_roles(self.kgs.test_entities2), self._mapping_array()

    def save(self):
        ent, _ = F.to_device_rows(self._both_roles(), normalize=True)     # preprocessing.normalize (simple.py:112)
        ent_embeds = ent[:, :self.args.dim].cpu().numpy()
        rel_embeds = (self.rel_embeds1.lookup() + self.rel_embeds2.lookup()).cpu().numpy()
        mapping_mat = self.mapping_mat.raw().cpu().numpy() if self.mapping_mat is not None else None
        rd.save_embeddings(self.out_folder, self.kgs, ent_embeds, rel_embeds, None, mapping_mat=mapping_mat)
