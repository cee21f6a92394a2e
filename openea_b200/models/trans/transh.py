"""TransH (models/trans/transh.py of the reference): TransE on hyperplane projections e⊥ = e − <e, n̂>n̂ with a
per-relation normal vector, margin-based loss, one uniformly corrupted negative per positive.  The whole graph
(lookups, projection, margin_loss, gradients) is the OEA_MODEL_TRANSH instance of oea_model_score_fed."""
from openea_b200 import engine as eng
from openea_b200.models.trans.transe import TransE
from openea_b200.modules.base.initializers import init_embeddings


class TransH(TransE):

    def __init__(self):
        super().__init__()
        self.normal_vector = None

    def _define_variables(self):
        super()._define_variables()
        self.normal_vector = init_embeddings([self.kgs.relations_num, self.args.dim], 'normal_vector',
                                             self.args.init, True, optimizer=self.args.optimizer)   # transh.py:21-22

    def _define_embed_graph(self):
        # margin_loss(phs, prs, pts, nhs, nrs, nts, margin, loss_norm) whatever args.loss says (transh.py:45)
        self.triple_trainer = eng.ModelTrainer("TransH", (self.ent_embeds, self.rel_embeds, None, self.normal_vector),
                                               eng.loss_cfg("margin-based", self.args.loss_norm, margin=self.args.margin),
                                               self.args.learning_rate)
        self.neg_per_pos = self.args.neg_triple_num
        self.triple_loss = self.triple_optimizer = self.triple_trainer
