"""CPU ORACLE (test infrastructure only) for the projected / bilinear score family of path (i), SURVEY §8f-2.

float64 torch restatement, TF op by TF op, of the graphs the reference builds for
  TransH    models/trans/transh.py:14-51      (margin_loss on hyperplane-projected h, t)
  TransD    models/trans/transd.py:14-65      (get_loss_func on l2_normalize(e + <e, e_p>·r_p))
  DistMult  models/semantic/distmult.py:36-59 (reduce_mean softplus(−label·Σ h∘r∘t))
  SimplE    models/semantic/simple.py:41-86   (softplus of ∓ the averaged two-direction score)
  BootEA_TransH  approaches/bootea_transh.py:57-95 (limited_loss on the TransH projection)
with tf.nn.embedding_lookup → indexing, tf.nn.l2_normalize(x, 1) → x·rsqrt(max(Σx², 1e-12)), reduce_sum → sum,
and `init_embeddings(..., is_l2_norm)` returning the normalised variable (modules/base/initializers.py:26-50),
so gradients flow through the normalisation into the raw variable.  The losses are those of
modules/base/losses.py:15-73.

PARITY: pinned to the reference's own graph code executed on oracle/tf1_shim.py (tests/test_reference_graph_goldens.py::
test_score_family_oracle_reproduces_the_reference_graph, 1e-9); TF's op / optimiser semantics, restated in that
interpreter, stay unpinned — TensorFlow 1.x is not in /root/reference and cannot be installed here, and the reference
ships no golden vectors for these graphs.  tests/test_oracle_triple_ext.py
checks this restatement against hand-computed known answers and against an independent closed-form statement
of the gradients (the formulas the CUDA kernels implement).
"""
import numpy as np
import torch

EPS = 1e-12
MODELS = ("TransE", "TransH", "TransD", "DistMult", "SimplE")
# table slots of every model, in the order of the C-ABI struct oea_model: (ent, rel, ent_aux, rel_aux)
SLOTS = {
    "TransE": ("ent", "rel"),
    "TransH": ("ent", "rel", None, "normal"),              # rel_aux = normal_vector
    "TransD": ("ent", "rel", "ent_transfer", "rel_transfer"),
    "DistMult": ("ent", "rel"),
    "SimplE": ("head_ent", "rel1", "tail_ent", "rel2"),
}


def l2n(x):
    """tf.nn.l2_normalize(x, 1)."""
    return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=EPS))


def _var(tab, norm):
    """What init_embeddings returns: the variable itself or l2_normalize(variable, 1)."""
    return l2n(tab) if norm else tab


def _score(u, loss_norm):
    """losses.py: 'L1' → Σ|u|, anything else → Σu² (squared, no sqrt)."""
    return u.abs().sum(1) if loss_norm == "L1" else (u * u).sum(1)


def energies(model, tabs, norms, hrt, loss_norm="L2"):
    """Per-triple 'energy' (low = plausible) of triples hrt [3, n]: the translation distance for the Trans* models,
    MINUS the similarity score for DistMult / SimplE (so that one set of loss formulas serves both)."""
    h, r, t = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in hrt)
    if model == "TransE":
        E, R = _var(tabs["ent"], norms["ent"]), _var(tabs["rel"], norms["rel"])
        return _score(E[h] + R[r] - E[t], loss_norm)
    if model == "TransH":
        E, R = _var(tabs["ent"], norms["ent"]), _var(tabs["rel"], norms["rel"])
        N = _var(tabs["normal"], norms["normal"])                         # transh.py:21-22: is_l2_norm=True
        n = l2n(N[r])                                                      # transh.py:50 (normalised again in _calc)
        calc = lambda e: e - (e * n).sum(1, keepdim=True) * n              # transh.py:51
        return _score(calc(E[h]) + R[r] - calc(E[t]), loss_norm)
    if model == "TransD":
        E, R = _var(tabs["ent"], norms["ent"]), _var(tabs["rel"], norms["rel"])
        Et, Rt = _var(tabs["ent_transfer"], norms["ent_transfer"]), _var(tabs["rel_transfer"], norms["rel_transfer"])
        calc = lambda e, t_, r_: l2n(e + (e * t_).sum(1, keepdim=True) * r_)   # transd.py:64-65
        return _score(calc(E[h], Et[h], Rt[r]) + R[r] - calc(E[t], Et[t], Rt[r]), loss_norm)
    if model == "DistMult":
        E, R = _var(tabs["ent"], norms["ent"]), _var(tabs["rel"], norms["rel"])
        return -(E[h] * R[r] * E[t]).sum(1)                                # distmult.py:43-44,57
    if model == "SimplE":
        H, T = _var(tabs["head_ent"], norms["head_ent"]), _var(tabs["tail_ent"], norms["tail_ent"])
        R1, R2 = _var(tabs["rel1"], norms["rel1"]), _var(tabs["rel2"], norms["rel2"])
        calc = lambda hs, rs, ts: (l2n(hs * rs) * ts).sum(1)               # simple.py:50-54
        return -(calc(H[h], R1[r], T[t]) + calc(H[t], R2[r], T[h])) / 2    # simple.py:57
    raise ValueError(model)


def loss_value(e_pos, e_neg, loss, margin=0.0, neg_margin=0.0, balance=1.0, scale=1.0):
    """modules/base/losses.py on energies.  `scale` multiplies the summed loss (DistMult's reduce_mean =
    1/(n_pos+n_neg), distmult.py:58)."""
    sp = torch.nn.functional.softplus
    if loss == "margin-based":          # losses.py:15-27
        assert e_neg is not None and e_pos.shape == e_neg.shape
        return scale * torch.relu(margin + e_pos - e_neg).sum()
    if loss == "limited":               # losses.py:42-56
        return scale * (torch.relu(e_pos - margin).sum() + balance * torch.relu(neg_margin - e_neg).sum())
    if loss == "logistic":              # losses.py:59-73 ; DistMult / SimplE: softplus(−score⁺) + softplus(score⁻)
        return scale * (sp(e_pos).sum() + (sp(-e_neg).sum() if e_neg is not None else 0.0))
    if loss == "positive":
        return scale * e_pos.sum()
    raise ValueError(loss)


def fwd_bwd(model, tabs, norms, pos, neg, loss, loss_norm="L2", margin=0.0, neg_margin=0.0, balance=1.0, scale=1.0):
    """tabs: {slot name: float array [rows, d]} raw variables; norms: {slot name: bool}.
    Returns (loss, {slot: d loss / d raw variable as float64 arrays}, energies [n_pos + n_neg])."""
    tv = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in tabs.items()}
    e_pos = energies(model, tv, norms, pos, loss_norm)
    e_neg = energies(model, tv, norms, neg, loss_norm) if neg is not None and np.asarray(neg).shape[1] > 0 else None
    val = loss_value(e_pos, e_neg, loss, margin, neg_margin, balance, scale)
    val.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in tv.items()}
    en = torch.cat([e_pos] + ([e_neg] if e_neg is not None else [])).detach().numpy()
    return float(val.detach()), grads, en


class DenseState:
    """All variables of one model + dense TF-1 optimiser slots (modules/base/optimizers.py:10-20: Adagrad with
    initial_accumulator_value 0.1 and no epsilon, SGD, TF-form Adam)."""

    def __init__(self, tabs, optimizer="Adagrad"):
        self.w = {k: np.asarray(v, dtype=np.float64).copy() for k, v in tabs.items()}
        self.optimizer, self.t = optimizer, 0
        if optimizer == "Adagrad":
            self.s1 = {k: np.full_like(v, 0.1) for k, v in self.w.items()}
        elif optimizer in ("Adam", "Adadelta"):
            self.s1 = {k: np.zeros_like(v) for k, v in self.w.items()}
            self.s2 = {k: np.zeros_like(v) for k, v in self.w.items()}

    def apply(self, grads, lr):
        self.t += 1
        for k, g in grads.items():
            if self.optimizer == "Adagrad":
                self.s1[k] += g * g
                self.w[k] -= lr * g / np.sqrt(self.s1[k])
            elif self.optimizer == "Adam":
                b1, b2, eps = 0.9, 0.999, 1e-8
                self.s1[k] = b1 * self.s1[k] + (1 - b1) * g
                self.s2[k] = b2 * self.s2[k] + (1 - b2) * g * g
                lr_t = lr * np.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
                self.w[k] -= lr_t * self.s1[k] / (np.sqrt(self.s2[k]) + eps)
            elif self.optimizer == "Adadelta":     # TF1 ApplyAdadelta, rho = 0.95, epsilon = 1e-8 (optimizers.py:13-15)
                rho, eps = 0.95, 1e-8
                self.s1[k] = rho * self.s1[k] + (1 - rho) * g * g
                upd = np.sqrt(self.s2[k] + eps) / np.sqrt(self.s1[k] + eps) * g
                self.w[k] -= lr * upd
                self.s2[k] = rho * self.s2[k] + (1 - rho) * upd * upd
            else:
                self.w[k] -= lr * g


def step(state, model, norms, pos, neg, loss, lr, **kw):
    """One session.run([triple_loss, triple_optimizer]); returns the batch loss."""
    val, grads, _ = fwd_bwd(model, state.w, norms, pos, neg, loss, **kw)
    state.apply(grads, lr)
    return val
