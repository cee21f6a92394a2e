"""Optimisers (interface of modules/base/optimizers.py:4-20): TF1 update rules on device tables.

`get_optimizer(opt, lr)` → an object with `.apply(table)`; `generate_optimizer(loss, lr, var_list, opt)` →
a callable train op for eager graphs built from `embedding_lookup` + the losses of this package: calling it
back-propagates `loss` into the tables' gradient buffers and applies the update (compute_gradients +
apply_gradients).  Each call owns its slot variables, as every tf.train.*Optimizer instance does.
"""
import ctypes as C

import torch

from openea_b200 import lib as L
from openea_b200.engine import EmbeddingTable, _ptr, _stream_ptr, opt_cfg


class RowOptimizer:
    def __init__(self, opt, learning_rate):
        self.kind = opt if opt in ('Adagrad', 'Adam', 'Adadelta') else 'SGD'
        self.lr = float(learning_rate)
        self._slots = {}

    def slots_for(self, table):
        """This optimiser instance's own view (own accumulators) of `table`."""
        key = id(table)
        if key not in self._slots:
            view = object.__new__(EmbeddingTable)
            view.__dict__.update(table.__dict__)
            view.optimizer = self.kind
            view._struct = None
            view.state1 = view.state2 = None
            view.adam_t = 0
            view._make_state()
            self._slots[key] = view
        return self._slots[key]

    def apply(self, table):
        lib = L.load()
        view = self.slots_for(table)
        if view.optimizer == "Adam":
            view.adam_t += 1
        cfg = opt_cfg(view, self.lr)
        L.check(lib.oea_rowopt_apply(C.byref(view.c_struct()), C.byref(cfg), _stream_ptr()), "oea_rowopt_apply")


def get_optimizer(opt, learning_rate):
    return RowOptimizer(opt, learning_rate)


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids, anchor):
        ctx.table = table
        ctx.ids = ids
        return table.lookup(ids)

    @staticmethod
    def backward(ctx, gout):
        lib = L.load()
        g = gout.contiguous().to(torch.float32)
        L.check(lib.oea_table_scatter_grad(C.byref(ctx.table.c_struct()), _ptr(ctx.ids), ctx.ids.numel(), _ptr(g),
                                           g.shape[1], _stream_ptr()), "oea_table_scatter_grad")
        return None, None, None


_ANCHOR = None


def embedding_lookup(table, ids):
    """tf.nn.embedding_lookup(table, ids) for eager graphs: normalised rows [n, dim] with a backward that
    scatter-adds through the normalisation into table.grad."""
    global _ANCHOR
    ids_t = torch.as_tensor(ids, dtype=torch.int32, device=table.device).contiguous()
    if _ANCHOR is None or _ANCHOR.device != table.device:
        _ANCHOR = torch.zeros(1, device=table.device, requires_grad=True)
    return _LookupFn.apply(table, ids_t, _ANCHOR)


def generate_optimizer(loss, learning_rate, var_list=None, opt='SGD'):
    """Eager equivalent of compute_gradients + apply_gradients: returns train_op(); train_op() → float loss."""
    optimizer = get_optimizer(opt, learning_rate)

    def train_op(loss_tensor=None, tables=None):
        lt = loss if loss_tensor is None else loss_tensor
        lt.backward()
        for tab in (tables if tables is not None else (var_list or [])):
            optimizer.apply(tab)
        return float(lt.detach().item())
    train_op.optimizer = optimizer
    return train_op
