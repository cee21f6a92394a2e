"""BootEA_TransH (approaches/bootea_transh.py of the reference): BootEA whose triple graph scores hyperplane
projections (TransH `_calc`, :57-59,83-91) under the limited loss with ε-truncated negatives; the alignment loss
(:97-105), bootstrapping, validation and the iteration loop are BootEA's."""
from openea_b200 import engine as eng
from openea_b200.approaches.bootea import BootEA
from openea_b200.models.basic_model import _loss_from_args
from openea_b200.modules.base.initializers import init_embeddings


class BootEA_TransH(BootEA):

    def __init__(self):
        super().__init__()
        self.normal_vector = None

    def _define_variables(self):
        super()._define_variables()
        self.normal_vector = init_embeddings([self.kgs.relations_num, self.args.dim], 'normal_vector',
                                             self.args.init, True, optimizer=self.args.optimizer)   # :67-68

    def _define_embed_graph(self):
        self.triple_trainer = eng.ModelTrainer("TransH", (self.ent_embeds, self.rel_embeds, None, self.normal_vector),
                                               _loss_from_args(self.args, 'limited'), self.args.learning_rate)
        self.neg_per_pos = self.args.neg_triple_num
        self.triple_loss = self.triple_optimizer = self.triple_trainer
