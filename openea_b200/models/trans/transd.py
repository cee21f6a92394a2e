"""TransD (models/trans/transd.py of the reference): e⊥ = l2_normalize(e + <e, e_p>·r_p) with entity and relation
transfer vectors, loss chosen by get_loss_func(args) — the OEA_MODEL_TRANSD instance of oea_model_score_fed."""
from openea_b200 import engine as eng
from openea_b200.models.basic_model import _loss_from_args
from openea_b200.models.trans.transe import TransE
from openea_b200.modules.base.initializers import init_embeddings


class TransD(TransE):

    def __init__(self):
        super().__init__()
        self.ent_transfer = None
        self.rel_transfer = None

    def _define_variables(self):
        super()._define_variables()
        a = self.args
        self.ent_transfer = init_embeddings([self.kgs.entities_num, a.dim], 'ent_transfer', a.init, a.ent_l2_norm,
                                            optimizer=a.optimizer)                                   # transd.py:20-23
        self.rel_transfer = init_embeddings([self.kgs.relations_num, a.dim], 'rel_transfer', a.init, a.rel_l2_norm,
                                            optimizer=a.optimizer)

    def _define_embed_graph(self):
        self.triple_trainer = eng.ModelTrainer("TransD", (self.ent_embeds, self.rel_embeds, self.ent_transfer,
                                                          self.rel_transfer),
                                               _loss_from_args(self.args), self.args.learning_rate)
        self.neg_per_pos = self.args.neg_triple_num
        self.triple_loss = self.triple_optimizer = self.triple_trainer
