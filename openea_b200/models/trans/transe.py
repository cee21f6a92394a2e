"""TransE (models/trans/transe.py of the reference): BasicModel with the margin-based loss, one uniformly
corrupted negative per positive (args_hander.py:19-21) — the margin mode of the fused K1 kernel."""
from openea_b200.models.basic_model import BasicModel
from openea_b200.modules.utils.util import load_session


class TransE(BasicModel):

    def __init__(self):
        super().__init__()

    def init(self):
        self.session = load_session()
        self._define_variables()
        self._define_embed_graph()
        if self.args.alignment_module == 'mapping':
            self._define_mapping_variables()
            self._define_mapping_graph()
        if self.args.loss == 'margin-based':
            assert self.args.neg_triple_num == 1
