"""Shared test helpers: float64 torch-autograd statement of the triple-scoring graph (an independent
derivation used to check the C oracle's hand-written backward) and seeded batch generators."""
import numpy as np
import torch


def torch_triple_loss(ent, rel, pos, neg, loss, loss_norm, ent_norm, rel_norm, margin=0.0, neg_margin=0.0,
                      balance=1.0, dtype=torch.float64):
    """Formulas of modules/base/losses.py + initializers.py:26 in torch; returns (loss, g_ent, g_rel, scores)."""
    E = torch.tensor(np.asarray(ent), dtype=dtype, requires_grad=True)
    R = torch.tensor(np.asarray(rel), dtype=dtype, requires_grad=True)

    def l2n(x):
        ss = (x * x).sum(1, keepdim=True)
        return x * torch.rsqrt(torch.clamp(ss, min=1e-12))

    Eh = l2n(E) if ent_norm else E
    Rh = l2n(R) if rel_norm else R

    def score(hrt):
        hrt = torch.as_tensor(np.asarray(hrt), dtype=torch.long)
        u = Eh[hrt[0]] + Rh[hrt[1]] - Eh[hrt[2]]
        return u.abs().sum(1) if loss_norm == "L1" else (u * u).sum(1)

    sp = score(pos)
    sn = score(neg) if neg is not None and np.asarray(neg).size else torch.zeros(0, dtype=dtype)
    relu = torch.nn.functional.relu
    if loss == "margin-based":
        val = relu(margin + sp - sn).sum()
    elif loss == "limited":
        val = relu(sp - margin).sum() + balance * relu(neg_margin - sn).sum()
    elif loss == "logistic":
        val = torch.log(1 + torch.exp(sp)).sum() + torch.log(1 + torch.exp(-sn)).sum()
    elif loss == "positive":
        val = sp.sum()
    elif loss == "logsigmoid":
        val = -torch.log(torch.sigmoid(-sp)).sum()
    else:
        raise ValueError(loss)
    val.backward()
    g_e = E.grad if E.grad is not None else torch.zeros_like(E)
    g_r = R.grad if R.grad is not None else torch.zeros_like(R)
    return float(val.detach()), g_e.numpy(), g_r.numpy(), torch.cat([sp, sn]).detach().numpy()


def make_tables(rng, n_ent, n_rel, d, scale=1.0):
    ent = (rng.standard_normal((n_ent, d)) * scale / np.sqrt(d)).astype(np.float32)
    rel = (rng.standard_normal((n_rel, d)) * scale / np.sqrt(d)).astype(np.float32)
    return ent, rel


def make_batch(rng, n_ent, n_rel, n_pos, k, hot=True):
    """Index batch with repeated rows (a few hot entities/relations) as Zipf-like KGs produce."""
    def ents(n):
        e = rng.integers(0, n_ent, size=n)
        if hot and n > 4:
            m = rng.random(n) < 0.3
            e[m] = rng.integers(0, max(1, min(5, n_ent)), size=int(m.sum()))
        return e
    pos = np.stack([ents(n_pos), rng.integers(0, n_rel, size=n_pos), ents(n_pos)]).astype(np.int32)
    if k == 0:
        return pos, None
    neg = np.repeat(pos, k, axis=1)
    side = rng.random(n_pos * k) < 0.5
    repl = ents(n_pos * k)
    neg[0, side] = repl[side]
    neg[2, ~side] = repl[~side]
    return pos, neg.astype(np.int32)
