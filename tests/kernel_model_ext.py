"""Closed-form statement (NumPy float64, vectorised over triples) of exactly the forward/backward arithmetic the
CUDA kernels of openea_b200/csrc/oea_triple_ext.cu perform per triple — same intermediate quantities, same order
of the chain rule.  tests/test_oracle_triple_ext.py checks it against the torch-autograd oracle on the CPU, so an
error in the hand-derived gradients is caught here, before a GPU sees the kernel.  Test infrastructure only."""
import numpy as np

EPS = 1e-12


def norm_fwd(x, on):
    """x̂ = x·rsqrt(max(Σx², eps)) when `on` (inv_norm of oea_rowmath.cuh)."""
    ss = (x * x).sum(1)
    inv = 1.0 / np.sqrt(np.maximum(ss, EPS)) if on else np.ones_like(ss)
    return x * inv[:, None], inv, ss


def through_norm(ghat, xhat, inv, ss, on):
    """d loss / d raw row from d loss / d normalised row (through_norm of oea_rowmath.cuh)."""
    dot = (xhat * ghat).sum(1)
    proj = np.where(np.logical_and(on, ss >= EPS), dot, 0.0)
    return (ghat - xhat * proj[:, None]) * inv[:, None]


def score_and_dir(u, loss_norm):
    if loss_norm == "L1":
        return np.abs(u).sum(1), np.sign(u)
    return (u * u).sum(1), 2.0 * u


def loss_grad(e_pos, e_neg, loss, margin=0.0, neg_margin=0.0, balance=1.0, scale=1.0):
    """(loss, d loss/d e_pos, d loss/d e_neg) as loss_of() / the margin kernels compute them (relu'(0) = 0)."""
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    sp = lambda x: np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))
    if loss == "margin-based":
        v = margin + e_pos - e_neg
        act = (v > 0).astype(np.float64)
        return scale * np.maximum(v, 0).sum(), scale * act, -scale * act
    if loss == "limited":
        L = np.maximum(e_pos - margin, 0).sum() + balance * np.maximum(neg_margin - e_neg, 0).sum()
        return scale * L, scale * (e_pos > margin), -scale * balance * (e_neg < neg_margin)
    if loss == "logistic":
        L = sp(e_pos).sum() + (sp(-e_neg).sum() if e_neg is not None else 0.0)
        return scale * L, scale * sig(e_pos), (None if e_neg is None else -scale * sig(-e_neg))
    if loss == "positive":
        return scale * e_pos.sum(), scale * np.ones_like(e_pos), None
    raise ValueError(loss)


class _Acc:
    def __init__(self, tabs):
        self.g = {k: np.zeros_like(np.asarray(v, dtype=np.float64)) for k, v in tabs.items()}

    def add(self, name, ids, rows):
        np.add.at(self.g[name], ids, rows)


def forward_backward(model, tabs, norms, hrt, gE, loss_norm, acc):
    """Energies of triples hrt [3, n]; when gE (d loss / d energy, [n]) is given, accumulates the raw-variable
    gradients into acc.  Mirrors Triple<MODEL>::forward / ::backward of oea_triple_ext.cu."""
    tabs = {k: np.asarray(v, dtype=np.float64) for k, v in tabs.items()}
    h, r, t = (np.asarray(x, dtype=np.int64) for x in hrt)
    col = lambda a: a[:, None]
    dot = lambda a, b: (a * b).sum(1)

    if model == "TransE":
        xh, ih, ssh = norm_fwd(tabs["ent"][h], norms["ent"])
        xt, it, sst = norm_fwd(tabs["ent"][t], norms["ent"])
        xr, ir, ssr = norm_fwd(tabs["rel"][r], norms["rel"])
        u = xh + xr - xt
        E, dirn = score_and_dir(u, loss_norm)
        if gE is not None:
            du = dirn * col(gE)
            acc.add("ent", h, through_norm(du, xh, ih, ssh, norms["ent"]))
            acc.add("ent", t, through_norm(-du, xt, it, sst, norms["ent"]))
            acc.add("rel", r, through_norm(du, xr, ir, ssr, norms["rel"]))
        return E

    if model == "TransH":
        xh, ih, ssh = norm_fwd(tabs["ent"][h], norms["ent"])
        xt, it, sst = norm_fwd(tabs["ent"][t], norms["ent"])
        xr, ir, ssr = norm_fwd(tabs["rel"][r], norms["rel"])
        n1, in1, ssn1 = norm_fwd(tabs["normal"][r], norms["normal"])     # the table's own normalisation
        n2, in2, ssn2 = norm_fwd(n1, True)                               # _calc normalises again
        a, b = dot(xh, n2), dot(xt, n2)
        u = xh + xr - xt - col(a - b) * n2
        E, dirn = score_and_dir(u, loss_norm)
        if gE is not None:
            du = dirn * col(gE)
            c = dot(du, n2)
            dh = du - col(c) * n2                                        # d/dĥ ; d/dt̂ = −dh
            dn2 = -col(a - b) * du - col(c) * (xh - xt)
            dn1 = through_norm(dn2, n2, in2, ssn2, True)
            acc.add("ent", h, through_norm(dh, xh, ih, ssh, norms["ent"]))
            acc.add("ent", t, through_norm(-dh, xt, it, sst, norms["ent"]))
            acc.add("rel", r, through_norm(du, xr, ir, ssr, norms["rel"]))
            acc.add("normal", r, through_norm(dn1, n1, in1, ssn1, norms["normal"]))
        return E

    if model == "TransD":
        xh, ih, ssh = norm_fwd(tabs["ent"][h], norms["ent"])
        xt, it, sst = norm_fwd(tabs["ent"][t], norms["ent"])
        ph, iph, ssph = norm_fwd(tabs["ent_transfer"][h], norms["ent_transfer"])
        pt, ipt, sspt = norm_fwd(tabs["ent_transfer"][t], norms["ent_transfer"])
        xr, ir, ssr = norm_fwd(tabs["rel"][r], norms["rel"])
        pr, ipr, sspr = norm_fwd(tabs["rel_transfer"][r], norms["rel_transfer"])
        a, b = dot(xh, ph), dot(xt, pt)
        vh, ivh, ssvh = norm_fwd(xh + col(a) * pr, True)                 # h⊥
        vt, ivt, ssvt = norm_fwd(xt + col(b) * pr, True)                 # t⊥
        u = vh + xr - vt
        E, dirn = score_and_dir(u, loss_norm)
        if gE is not None:
            du = dirn * col(gE)
            dvh = through_norm(du, vh, ivh, ssvh, True)
            dvt = through_norm(-du, vt, ivt, ssvt, True)
            eh, et = dot(dvh, pr), dot(dvt, pr)
            acc.add("ent", h, through_norm(dvh + col(eh) * ph, xh, ih, ssh, norms["ent"]))
            acc.add("ent", t, through_norm(dvt + col(et) * pt, xt, it, sst, norms["ent"]))
            acc.add("ent_transfer", h, through_norm(col(eh) * xh, ph, iph, ssph, norms["ent_transfer"]))
            acc.add("ent_transfer", t, through_norm(col(et) * xt, pt, ipt, sspt, norms["ent_transfer"]))
            acc.add("rel", r, through_norm(du, xr, ir, ssr, norms["rel"]))
            acc.add("rel_transfer", r, through_norm(col(a) * dvh + col(b) * dvt, pr, ipr, sspr, norms["rel_transfer"]))
        return E

    if model == "DistMult":
        xh, ih, ssh = norm_fwd(tabs["ent"][h], norms["ent"])
        xt, it, sst = norm_fwd(tabs["ent"][t], norms["ent"])
        xr, ir, ssr = norm_fwd(tabs["rel"][r], norms["rel"])
        E = -dot(xh * xr, xt)
        if gE is not None:
            w = -col(gE)                                                 # d loss / d score
            acc.add("ent", h, through_norm(w * xr * xt, xh, ih, ssh, norms["ent"]))
            acc.add("ent", t, through_norm(w * xh * xr, xt, it, sst, norms["ent"]))
            acc.add("rel", r, through_norm(w * xh * xt, xr, ir, ssr, norms["rel"]))
        return E

    if model == "SimplE":
        Hh, iHh, sHh = norm_fwd(tabs["head_ent"][h], norms["head_ent"])
        Ht, iHt, sHt = norm_fwd(tabs["head_ent"][t], norms["head_ent"])
        Th, iTh, sTh = norm_fwd(tabs["tail_ent"][h], norms["tail_ent"])
        Tt, iTt, sTt = norm_fwd(tabs["tail_ent"][t], norms["tail_ent"])
        r1, ir1, sr1 = norm_fwd(tabs["rel1"][r], norms["rel1"])
        r2, ir2, sr2 = norm_fwd(tabs["rel2"][r], norms["rel2"])
        q1, iq1, sq1 = norm_fwd(Hh * r1, True)                           # l2_normalize(hs·rs)
        q2, iq2, sq2 = norm_fwd(Ht * r2, True)
        E = -0.5 * (dot(q1, Tt) + dot(q2, Th))
        if gE is not None:
            w = -0.5 * col(gE)                                           # d loss / d each direction's score
            dp1 = through_norm(w * Tt, q1, iq1, sq1, True)
            dp2 = through_norm(w * Th, q2, iq2, sq2, True)
            acc.add("head_ent", h, through_norm(dp1 * r1, Hh, iHh, sHh, norms["head_ent"]))
            acc.add("head_ent", t, through_norm(dp2 * r2, Ht, iHt, sHt, norms["head_ent"]))
            acc.add("tail_ent", t, through_norm(w * q1, Tt, iTt, sTt, norms["tail_ent"]))
            acc.add("tail_ent", h, through_norm(w * q2, Th, iTh, sTh, norms["tail_ent"]))
            acc.add("rel1", r, through_norm(dp1 * Hh, r1, ir1, sr1, norms["rel1"]))
            acc.add("rel2", r, through_norm(dp2 * Ht, r2, ir2, sr2, norms["rel2"]))
        return E
    raise ValueError(model)


def fwd_bwd(model, tabs, norms, pos, neg, loss, loss_norm="L2", **kw):
    """Same contract as oracle.triple_ext.fwd_bwd."""
    has_neg = neg is not None and np.asarray(neg).shape[1] > 0
    e_pos = forward_backward(model, tabs, norms, pos, None, loss_norm, None)
    e_neg = forward_backward(model, tabs, norms, neg, None, loss_norm, None) if has_neg else None
    val, g_pos, g_neg = loss_grad(e_pos, e_neg, loss, **kw)
    acc = _Acc(tabs)
    forward_backward(model, tabs, norms, pos, g_pos, loss_norm, acc)
    if has_neg:
        forward_backward(model, tabs, norms, neg, g_neg, loss_norm, acc)
    return float(val), acc.g, np.concatenate([e_pos] + ([e_neg] if has_neg else []))
