"""CPU ORACLE for path (ii) (GCN-Align aggregation + L1 alignment loss) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the TF part: the graph of approaches/gcn_align.py runs in TensorFlow 1.x (absent here).
This restates it with torch-CPU autograd at the reference's call sites:
  GraphConvolution._call  gcn_align.py:239-267   (featureless first layer, weight-less second layer)
  trunc_normal + l2_normalize(·, 1)  gcn_align.py:52-56
  align_loss  gcn_align.py:298-320
  GradientDescentOptimizer  gcn_align.py:511
The adjacency builders ARE pinned: `reference_gcn_utils()` executes the reference's own GCN_Utils / load_attr
source (pure NumPy/SciPy, extracted from the file because the module itself imports TensorFlow).
"""
import ast
import os

import numpy as np
import scipy.sparse as sp
import torch

REF_FILE = "/root/reference/src/openea/approaches/gcn_align.py"


def reference_gcn_utils():
    """(GCN_Utils class, load_attr function) compiled from the reference source; None if it is not present."""
    if not os.path.exists(REF_FILE):
        return None
    with open(REF_FILE) as f:
        tree = ast.parse(f.read())
    keep = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name == "GCN_Utils")
            or (isinstance(n, ast.FunctionDef) and n.name == "load_attr")]
    ns = {"np": np, "sp": sp, "merge_dic": lambda a, b: {**a, **b}, "eigsh": None, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF_FILE, "exec"), ns)
    return ns["GCN_Utils"], ns["load_attr"]


def _sparse(mat, dtype):
    m = sp.coo_matrix(mat)
    idx = torch.tensor(np.vstack([m.row, m.col]), dtype=torch.long)
    return torch.sparse_coo_tensor(idx, torch.tensor(m.data, dtype=dtype), m.shape).coalesce()


def l2n(x):
    return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12))


def unit_forward(support, W, features=None, dtype=torch.float32):
    """OUT = A·relu(A·pre), pre = l2n(W) (featureless) or X_attr·l2n(W)."""
    A = _sparse(support, dtype)
    Wn = l2n(W)
    pre = Wn if features is None else torch.sparse.mm(_sparse(features, dtype), Wn)
    h1 = torch.relu(torch.sparse.mm(A, pre))
    return torch.sparse.mm(A, h1)


def align_loss(out, ill, gamma, k, neg_left, neg_right, neg2_left, neg2_right):
    left, right = torch.as_tensor(ill[:, 0], dtype=torch.long), torch.as_tensor(ill[:, 1], dtype=torch.long)
    t = len(ill)
    a = (out[left] - out[right]).abs().sum(1)
    d = (a + gamma).reshape(t, 1)
    tot = 0.0
    for nl, nr in ((neg_left, neg_right), (neg2_left, neg2_right)):
        b = (out[torch.as_tensor(nl, dtype=torch.long)] - out[torch.as_tensor(nr, dtype=torch.long)]).abs().sum(1)
        tot = tot + torch.relu(-b.reshape(t, k) + d).sum()
    return tot / (2.0 * k * t)


def unit_train_step(support, W0, features, ill, gamma, k, negs, lr, dtype=torch.float32):
    """One session.run([loss, opt_op]) of a GCN_Align_Unit: returns (loss, updated W, outputs)."""
    W = torch.tensor(np.asarray(W0), dtype=dtype, requires_grad=True)
    out = unit_forward(support, W, features, dtype)
    loss = align_loss(out, np.asarray(ill), gamma, k, *negs)
    loss.backward()
    with torch.no_grad():
        W_new = W - lr * W.grad
    return float(loss.detach()), W_new.numpy(), out.detach().numpy()
