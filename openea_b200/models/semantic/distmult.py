"""DistMult (models/semantic/distmult.py of the reference): similarity Σ h∘r∘t, loss reduce_mean softplus(−label·score)
over the batch of positives (label +1) and negatives (label −1) made by generate_triple_label_batch
(modules/train/batch.py:168-184, the with-replacement sampler generate_neg_triples), Adagrad regardless of
args.optimizer (distmult.py:59) — the OEA_MODEL_DISTMULT instance of oea_model_score_fed with loss_scale = 1/batch."""
import time

from openea_b200 import engine as eng
from openea_b200.models.basic_model import BasicModel
from openea_b200.modules.base.initializers import init_embeddings
from openea_b200.modules.utils.util import load_session


class DistMult(BasicModel):

    def __init__(self):
        super().__init__()
        self.metric = 'inner'

    def init(self):
        self._define_variables()
        self._define_mapping_variables()
        self._define_embed_graph()
        self._define_mapping_graph()
        self.session = load_session()

    def _define_variables(self):
        a = self.args
        self.ent_embeds = init_embeddings([self.kgs.entities_num, a.dim], 'ent_embeds', a.init, a.ent_l2_norm,
                                          optimizer='Adagrad')
        self.rel_embeds = init_embeddings([self.kgs.relations_num, a.dim], 'rel_embeds', a.init, a.rel_l2_norm,
                                          optimizer='Adagrad')

    def _define_embed_graph(self):
        self.triple_trainer = eng.ModelTrainer("DistMult", (self.ent_embeds, self.rel_embeds),
                                               eng.loss_cfg("logistic", "L2"), self.args.learning_rate,
                                               mean_loss=True, sampler="independent")
        self.neg_per_pos = self.args.neg_triple_num
        self.triple_loss = self.triple_optimizer = self.triple_trainer

    def launch_triple_training_1epo(self, epoch, triple_steps, steps_tasks, batch_queue, neighbors1, neighbors2):
        """distmult.py:61-88: the epoch loss is the SUM of the per-batch mean losses (not divided again)."""
        start = time.time()
        kg1, kg2, tset = self._device_kgs()
        self._install_candidates(kg1, kg2, neighbors1, neighbors2)
        self._epoch_seed = (self._epoch_seed * 6364136223846793005 + 1442695040888963407) & ((1 << 63) - 1)
        for step in range(triple_steps):
            self.triple_trainer.step_sampled(kg1, kg2, tset, self.args.batch_size, self.neg_per_pos, step,
                                             self._epoch_seed)
        epoch_loss = self.triple_trainer.read_loss()
        print('epoch {}, triple loss: {:.4f}, cost time: {:.4f}s'.format(epoch, epoch_loss, time.time() - start))
        return epoch_loss
