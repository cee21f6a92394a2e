"""Mapping module of the 'mapping' alignment mode (interface of modules/base/mapping.py:9-25)."""
from openea_b200.engine import MappingTrainer
from openea_b200.modules.base.initializers import orthogonal_init


def add_mapping_variables(model):
    model.mapping_mat = orthogonal_init([model.args.dim, model.args.dim], 'mapping_matrix',
                                        optimizer=model.args.optimizer)
    model.eye_mat = None          # the identity is implicit in oea_mapping_fwd_bwd


def add_mapping_module(model):
    model.mapping_trainer = MappingTrainer(model.ent_embeds, model.mapping_mat, model.args.alpha,
                                           model.args.learning_rate)
    model.mapping_loss = model.mapping_trainer         # handles kept under the reference's attribute names
    model.mapping_optimizer = model.mapping_trainer
