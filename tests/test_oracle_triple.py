"""CPU tests of the path-(i) oracle (oracle/oea_oracle.c): hand-computed known answers for every loss of
modules/base/losses.py, and agreement with a float64 torch-autograd statement of the same formulas."""
import numpy as np
import pytest

from oracle import triple as orc
from tests.helpers import make_batch, make_tables, torch_triple_loss


def test_known_answers_l2_no_norm():
    # h + r - t by hand: triple0 u=(1,0) s=1 ; triple1 u=(0,2) s=4 ; neg0 u=(3,0) s=9 ; neg1 u=(0,1) s=1
    ent = np.array([[1, 0], [0, 0], [0, 2], [3, 0], [0, 1]], dtype=np.float32)
    rel = np.zeros((1, 2), dtype=np.float32)
    pos = np.array([[0, 2], [0, 0], [1, 1]], dtype=np.int32)
    neg = np.array([[3, 4], [0, 0], [1, 1]], dtype=np.int32)
    loss, ge, gr, sc = orc.fwd_bwd(ent, rel, pos, neg, "limited", "L2", False, False, margin=2.0, neg_margin=4.0, balance=0.5)
    np.testing.assert_allclose(sc, [1, 4, 9, 1])
    # pos: relu(1-2)=0, relu(4-2)=2 ; neg: .5*relu(4-9)=0, .5*relu(4-1)=1.5
    assert loss == pytest.approx(3.5)
    # active: pos1 (g=+1, du=2u=(0,4)) on h=2,t=1 ; neg1 (g=-.5, du=-.5*2*(0,1)=(0,-1)) on h=4,t=1
    np.testing.assert_allclose(ge[2], [0, 4]); np.testing.assert_allclose(ge[4], [0, -1])
    np.testing.assert_allclose(ge[1], [0, -4 + 1]); np.testing.assert_allclose(ge[0], [0, 0]); np.testing.assert_allclose(ge[3], [0, 0])
    np.testing.assert_allclose(gr[0], [0, 4 - 1])

    loss, ge, gr, sc = orc.fwd_bwd(ent, rel, pos, None, "positive", "L2", False, False)
    assert loss == pytest.approx(5.0)
    loss, _, _, _ = orc.fwd_bwd(ent, rel, pos, neg, "margin-based", "L2", False, False, margin=1.5)
    # relu(1.5+1-9)=0 ; relu(1.5+4-1)=4.5
    assert loss == pytest.approx(4.5)
    loss, _, _, _ = orc.fwd_bwd(ent, rel, pos, neg, "logistic", "L2", False, False)
    want = np.log1p(np.exp(1.0)) + np.log1p(np.exp(4.0)) + np.log1p(np.exp(-9.0)) + np.log1p(np.exp(-1.0))
    assert loss == pytest.approx(want, rel=1e-6)
    loss, _, _, _ = orc.fwd_bwd(ent, rel, pos, None, "logsigmoid", "L2", False, False)
    assert loss == pytest.approx(np.log1p(np.exp(1.0)) + np.log1p(np.exp(4.0)), rel=1e-6)


def test_known_answers_l1_and_margin_edge():
    ent = np.array([[1, -2], [0, 0]], dtype=np.float32)
    rel = np.array([[0.5, 0.5]], dtype=np.float32)
    pos = np.array([[0], [0], [1]], dtype=np.int32)   # u = (1.5, -1.5) → L1 score 3
    loss, ge, gr, sc = orc.fwd_bwd(ent, rel, pos, None, "positive", "L1", False, False)
    assert sc[0] == pytest.approx(3.0) and loss == pytest.approx(3.0)
    np.testing.assert_allclose(ge[0], [1, -1]); np.testing.assert_allclose(ge[1], [-1, 1]); np.testing.assert_allclose(gr[0], [1, -1])
    # exactly at the margin: relu'(0) = 0 → no gradient (TF convention)
    loss, ge, gr, _ = orc.fwd_bwd(ent, rel, pos, None, "limited", "L1", False, False, margin=3.0, neg_margin=5.0)
    assert loss == 0.0 and not ge.any() and not gr.any()


def test_normalised_lookup_known_answer():
    # unit rows after normalisation: h=(1,0), t=(0,1), r=(0,0)->stays 0 (eps clamp) ; u=(1,-1) s=2
    ent = np.array([[3, 0], [0, 5]], dtype=np.float32)
    rel = np.zeros((1, 2), dtype=np.float32)
    pos = np.array([[0], [0], [1]], dtype=np.int32)
    loss, ge, gr, sc = orc.fwd_bwd(ent, rel, pos, None, "positive", "L2", True, True)
    assert sc[0] == pytest.approx(2.0)
    # d/dĥ = 2u = (2,-2); projected off ĥ=(1,0): (0,-2) / ||h||=3
    np.testing.assert_allclose(ge[0], [0, -2 / 3], rtol=1e-6)
    # d/dt̂ = (-2, 2); projected off t̂=(0,1): (-2, 0) / 5
    np.testing.assert_allclose(ge[1], [-2 / 5, 0], rtol=1e-6)
    # zero relation row: Σx² < eps → Jacobian is rsqrt(eps) = 1e6 scaling
    np.testing.assert_allclose(gr[0], [2e6, -2e6], rtol=1e-5)


@pytest.mark.parametrize("loss,k", [("limited", 10), ("logistic", 3), ("positive", 0), ("logsigmoid", 0), ("margin-based", 1)])
@pytest.mark.parametrize("loss_norm", ["L1", "L2"])
@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("d", [75, 100])
def test_oracle_matches_float64_autograd(loss, k, loss_norm, norm, d):
    rng = np.random.default_rng(1234 + d + k)
    ent, rel = make_tables(rng, 300, 17, d, scale=1.0 if norm else 0.9)
    pos, neg = make_batch(rng, 300, 17, 64, k)
    kw = dict(margin=1.1 if loss == "margin-based" else 0.3, neg_margin=2.2, balance=0.2)
    lo, ge, gr, sc = orc.fwd_bwd(ent, rel, pos, neg, loss, loss_norm, norm, norm, **kw)
    lt, te, tr, ts = torch_triple_loss(ent, rel, pos, neg, loss, loss_norm, norm, norm, **kw)
    np.testing.assert_allclose(sc, ts, rtol=2e-5, atol=1e-6)
    assert lo == pytest.approx(lt, rel=1e-5)
    if loss_norm == "L1":
        # sign() flips where |u| is at float32 resolution; compare away from those coordinates only via norms
        assert np.abs(ge - te).max() <= 2.0 * 1e-3 * max(1.0, np.abs(te).max())
    else:
        np.testing.assert_allclose(ge, te, rtol=1e-4, atol=2e-5 * np.abs(te).max())
        np.testing.assert_allclose(gr, tr, rtol=1e-4, atol=2e-5 * np.abs(tr).max())


def test_dense_step_rules():
    rng = np.random.default_rng(7)
    ent, rel = make_tables(rng, 50, 5, 16)
    pos, neg = make_batch(rng, 50, 5, 20, 4)
    for opt in ("Adagrad", "SGD", "Adam"):
        st = orc.DenseState(ent, rel, opt)
        _, ge, gr, _ = orc.fwd_bwd(ent, rel, pos, neg, "limited", "L2", True, True, margin=0.01, neg_margin=2.0, balance=0.2)
        orc.step(st, pos, neg, "limited", "L2", True, True, 0.01, margin=0.01, neg_margin=2.0, balance=0.2)
        if opt == "Adagrad":   # acc0 = 0.1, no epsilon
            want = ent - 0.01 * ge / np.sqrt(0.1 + ge * ge)
        elif opt == "SGD":
            want = ent - 0.01 * ge
        else:                  # TF Adam, t = 1: lr_t = lr*sqrt(1-b2)/(1-b1); m = .1 g; v = .001 g²
            lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
            want = ent - lr_t * (0.1 * ge) / (np.sqrt(0.001 * ge * ge) + 1e-8)
        np.testing.assert_allclose(st.ent, want, rtol=2e-5, atol=1e-7)
        untouched = ~np.isin(np.arange(50), np.concatenate([pos[0], pos[2], neg[0], neg[2]]))
        if opt != "Adam" and untouched.any():
            np.testing.assert_array_equal(st.ent[untouched], ent[untouched])
