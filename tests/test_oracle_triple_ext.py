"""CPU tests of the oracle for the projected / bilinear score family (oracle/triple_ext.py): hand-computed known
answers, and agreement of the torch-autograd restatement with the closed-form gradients the CUDA kernels
implement (tests/kernel_model_ext.py)."""
import numpy as np
import pytest

from oracle import triple_ext as ox
from tests import kernel_model_ext as km


def make_case(model, seed, n_ent=40, n_rel=7, d=12, n_pos=50, k=3, all_norm=None):
    rng = np.random.default_rng(seed)
    slots = [s for s in ox.SLOTS[model] if s]
    tabs, norms = {}, {}
    for i, s in enumerate(slots):
        rows = n_rel if s.startswith("rel") or s == "normal" else n_ent
        tabs[s] = rng.standard_normal((rows, d)) * (0.3 + 0.4 * i)
        norms[s] = bool(rng.integers(0, 2)) if all_norm is None else all_norm
    if model == "TransH":
        norms["normal"] = True        # transh.py:21-22
    pos = np.stack([rng.integers(0, n_ent, n_pos), rng.integers(0, n_rel, n_pos), rng.integers(0, n_ent, n_pos)])
    neg = np.repeat(pos, k, axis=1)
    side = rng.integers(0, 2, neg.shape[1]).astype(bool)
    neg[0, side] = rng.integers(0, n_ent, side.sum())
    neg[2, ~side] = rng.integers(0, n_ent, (~side).sum())
    return tabs, norms, pos, neg


LOSSES = [("limited", dict(margin=0.4, neg_margin=2.5, balance=0.3)),
          ("logistic", dict()),
          ("logistic", dict(scale=1.0 / 200)),
          ("positive", dict()),
          ("margin-based", dict(margin=1.5))]


@pytest.mark.parametrize("model", ox.MODELS)
@pytest.mark.parametrize("loss,kw", LOSSES)
@pytest.mark.parametrize("loss_norm", ["L2", "L1"])
def test_closed_form_gradients_match_autograd(model, loss, kw, loss_norm):
    if model in ("DistMult", "SimplE") and (loss_norm == "L1" or loss not in ("logistic",)):
        pytest.skip("similarity models use the softplus losses only")
    for seed in range(3):
        tabs, norms, pos, neg = make_case(model, 100 * seed + len(model), k=1 if loss == "margin-based" else 3)
        if loss == "positive":
            neg = None
        want_l, want_g, want_e = ox.fwd_bwd(model, tabs, norms, pos, neg, loss, loss_norm=loss_norm, **kw)
        got_l, got_g, got_e = km.fwd_bwd(model, tabs, norms, pos, neg, loss, loss_norm=loss_norm, **kw)
        np.testing.assert_allclose(got_e, want_e, rtol=1e-10, atol=1e-12)
        assert got_l == pytest.approx(want_l, rel=1e-10)
        for name in want_g:
            scale = np.abs(want_g[name]).max() + 1e-30
            np.testing.assert_allclose(got_g[name], want_g[name], rtol=1e-8, atol=1e-10 * scale, err_msg=name)


def test_transh_known_answer():
    # n = (0, 1): projection removes the y component.  h=(3,4)→(3,0), t=(1,7)→(1,0), r=(0.5,2): u=(2.5,2), s=6.25+4
    tabs = {"ent": np.array([[3.0, 4.0], [1.0, 7.0]]), "rel": np.array([[0.5, 2.0]]), "normal": np.array([[0.0, 5.0]])}
    norms = {"ent": False, "rel": False, "normal": True}
    pos = np.array([[0], [0], [1]])
    val, g, e = ox.fwd_bwd("TransH", tabs, norms, pos, None, "positive", loss_norm="L2")
    assert e[0] == pytest.approx(10.25) and val == pytest.approx(10.25)
    # du = 2u = (5, 4); dĥ = du − (du·n)n = (5, 0); dt̂ = (−5, 0); dr = (5, 4)
    np.testing.assert_allclose(g["ent"], [[5, 0], [-5, 0]], atol=1e-12)
    np.testing.assert_allclose(g["rel"], [[5, 4]], atol=1e-12)
    # dn̂ = −(a−b)·du − c·(h−t) with a−b = 4−7 = −3, c = 4: (15,12) − 4·(2,−3) = (7, 24); off n̂=(0,1): (7, 0); / ‖n‖ = 5
    np.testing.assert_allclose(g["normal"], [[7 / 5, 0]], atol=1e-12)


def test_transd_known_answer():
    # h=(1,0), h_p=(2,0), r_p=(0,1): <h,h_p>=2 → v=(1,2) → h⊥=(1,2)/√5 ; t=(0,1), t_p=(0,0) → t⊥=(0,1) ; r=(0,0)
    tabs = {"ent": np.array([[1.0, 0.0], [0.0, 1.0]]), "ent_transfer": np.array([[2.0, 0.0], [0.0, 0.0]]),
            "rel": np.zeros((1, 2)), "rel_transfer": np.array([[0.0, 1.0]])}
    norms = {k: False for k in tabs}
    pos = np.array([[0], [0], [1]])
    _, _, e = ox.fwd_bwd("TransD", tabs, norms, pos, None, "positive", loss_norm="L1")
    assert e[0] == pytest.approx(1 / np.sqrt(5) + abs(2 / np.sqrt(5) - 1))


def test_distmult_and_simple_known_answers():
    tabs = {"ent": np.array([[1.0, 2.0], [3.0, -1.0]]), "rel": np.array([[2.0, 0.5]])}
    norms = {"ent": False, "rel": False}
    pos = np.array([[0], [0], [1]])                    # score = 1·2·3 + 2·.5·(−1) = 5
    neg = np.array([[1], [0], [1]])                    # score = 9·2 + 1·.5 = 18.5
    val, g, e = ox.fwd_bwd("DistMult", tabs, norms, pos, neg, "logistic", scale=0.5)
    np.testing.assert_allclose(e, [-5.0, -18.5])
    assert val == pytest.approx(0.5 * (np.log1p(np.exp(-5.0)) + np.log1p(np.exp(18.5))))
    # SimplE with unit-norm products: H[h]∘r1 = (1,0) ; T[t] = (0.6, 0.8) → dir1 = 0.6 ; H[t]∘r2 = (0,2)→(0,1); T[h] = (0, 0.5) → dir2 = 0.5
    tabs = {"head_ent": np.array([[1.0, 0.0], [0.0, 2.0]]), "tail_ent": np.array([[0.0, 0.5], [0.6, 0.8]]),
            "rel1": np.array([[1.0, 1.0]]), "rel2": np.array([[1.0, 1.0]])}
    norms = {k: False for k in tabs}
    val, _, e = ox.fwd_bwd("SimplE", tabs, norms, pos, None, "logistic")
    assert e[0] == pytest.approx(-0.55)
    assert val == pytest.approx(np.log1p(np.exp(-0.55)))


def test_dense_state_adagrad_matches_tf_rule():
    tabs = {"ent": np.array([[1.0, 2.0], [3.0, -1.0]]), "rel": np.array([[2.0, 0.5]])}
    st = ox.DenseState(tabs, "Adagrad")
    g = {"ent": np.array([[0.5, 0.0], [0.0, 0.0]]), "rel": np.zeros((1, 2))}
    st.apply(g, lr=0.1)
    # acc = 0.1 + 0.25 → w −= 0.1·0.5/√0.35 ; rows with g = 0 are unchanged
    assert st.w["ent"][0, 0] == pytest.approx(1.0 - 0.05 / np.sqrt(0.35))
    assert st.w["ent"][1, 0] == 3.0 and st.w["rel"][0, 0] == 2.0
