"""ctypes binding of the C oracle for path (i) (oracle/oea_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboea_oracle.so")
_SRC = os.path.join(_HERE, "oea_oracle.c")

SCORE = {"L1": 0, "L2": 1}
LOSS = {"margin-based": 0, "limited": 1, "logistic": 2, "positive": 3, "logsigmoid": 4}
OPT = {"SGD": 0, "Adagrad": 1, "Adam": 2}

_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-std=c11",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        lib.orc_triple_fwd_bwd.restype = C.c_double
        lib.orc_triple_fwd_bwd.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           i32p, i32p, i32p, C.c_int, i32p, i32p, i32p, C.c_int,
                                           C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p]
        lib.orc_opt_dense.restype = None
        lib.orc_opt_dense.argtypes = [C.c_int, f32p, f32p, f32p, f32p, C.c_size_t, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_int]
        lib.orc_triple_step.restype = C.c_double
        lib.orc_triple_step.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        i32p, i32p, i32p, C.c_int, i32p, i32p, i32p, C.c_int,
                                        C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                        C.c_int, C.c_float, f32p, f32p, f32p, f32p, C.c_int]
        lib.orc_l2_normalize_rows.restype = None
        lib.orc_l2_normalize_rows.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p]
        lib.orc_num_threads.restype = C.c_int
        lib.orc_set_num_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _idx(hrt):
    if hrt is None:
        z = np.zeros(0, dtype=np.int32)
        return z, z, z, 0
    a = np.ascontiguousarray(np.asarray(hrt, dtype=np.int32))
    assert a.ndim == 2 and a.shape[0] == 3
    return a[0].copy(), a[1].copy(), a[2].copy(), a.shape[1]


def num_threads():
    return load().orc_num_threads()


def set_num_threads(n):
    load().orc_set_num_threads(int(n))


def l2_normalize(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    load().orc_l2_normalize_rows(_f(x), x.shape[0], x.shape[1], _f(y), None, None)
    return y


def fwd_bwd(ent, rel, pos, neg, loss, loss_norm, ent_norm, rel_norm, margin=0.0, neg_margin=0.0, balance=1.0):
    """Returns (loss, g_ent [N,d], g_rel [R,d], scores [n_pos+n_neg])."""
    lib = load()
    ent = np.ascontiguousarray(ent, dtype=np.float32)
    rel = np.ascontiguousarray(rel, dtype=np.float32)
    ph, pr, pt, n_pos = _idx(pos)
    nh, nr, nt, n_neg = _idx(neg)
    g_ent = np.zeros_like(ent)
    g_rel = np.zeros_like(rel)
    scores = np.zeros(n_pos + n_neg, dtype=np.float32)
    val = lib.orc_triple_fwd_bwd(_f(ent), ent.shape[0], _f(rel), rel.shape[0], ent.shape[1], int(ent_norm),
                                 int(rel_norm), _i(ph), _i(pr), _i(pt), n_pos, _i(nh), _i(nr), _i(nt), n_neg,
                                 SCORE["L1" if loss_norm == "L1" else "L2"], LOSS[loss], margin, neg_margin, balance,
                                 _f(g_ent), _f(g_rel), _f(scores))
    return float(val), g_ent, g_rel, scores


class DenseState:
    """The TF variables + optimiser slots as dense NumPy arrays (what the reference keeps on its device)."""

    def __init__(self, ent, rel, optimizer="Adagrad"):
        self.ent = np.ascontiguousarray(ent, dtype=np.float32).copy()
        self.rel = np.ascontiguousarray(rel, dtype=np.float32).copy()
        self.optimizer = optimizer
        self.t = 0
        if optimizer == "Adagrad":
            self.ent_s1 = np.full_like(self.ent, 0.1)  # initial_accumulator_value
            self.rel_s1 = np.full_like(self.rel, 0.1)
            self.ent_s2 = self.rel_s2 = None
        elif optimizer == "Adam":
            self.ent_s1, self.ent_s2 = np.zeros_like(self.ent), np.zeros_like(self.ent)
            self.rel_s1, self.rel_s2 = np.zeros_like(self.rel), np.zeros_like(self.rel)
        else:
            self.ent_s1 = self.ent_s2 = self.rel_s1 = self.rel_s2 = None


def step(state, pos, neg, loss, loss_norm, ent_norm, rel_norm, lr, margin=0.0, neg_margin=0.0, balance=1.0):
    """One session.run([triple_loss, triple_optimizer]) (basic_model.py:224-230). Returns the batch loss."""
    lib = load()
    ph, pr, pt, n_pos = _idx(pos)
    nh, nr, nt, n_neg = _idx(neg)
    state.t += 1
    val = lib.orc_triple_step(_f(state.ent), state.ent.shape[0], _f(state.rel), state.rel.shape[0],
                              state.ent.shape[1], int(ent_norm), int(rel_norm),
                              _i(ph), _i(pr), _i(pt), n_pos, _i(nh), _i(nr), _i(nt), n_neg,
                              SCORE["L1" if loss_norm == "L1" else "L2"], LOSS[loss], margin, neg_margin, balance,
                              OPT[state.optimizer], lr, _f(state.ent_s1), _f(state.ent_s2), _f(state.rel_s1),
                              _f(state.rel_s2), state.t)
    return float(val)
