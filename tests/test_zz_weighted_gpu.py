"""GPU: the weighted margin scorer, the pair-distance loss and IPTransE on them (csrc/oea_triple_weighted.cu,
approaches/iptranse.py).  The kernels' sources pass the same checks on the CPU warp emulator
(tests/test_emu_triple_core.py, tests/test_iptranse.py); these are their first runs on hardware."""
import numpy as np
import pytest
import torch

from tests.helpers import make_tables
from tests.test_e2e_gpu import tiny_kgs        # noqa: F401  (fixture: the tiny synthetic dataset folder)

pytestmark = [pytest.mark.gpu]


def _normed(x, on):
    return x * torch.rsqrt(torch.clamp((x * x).sum(1, keepdim=True), min=1e-12)) if on else x


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(got, want, what):
    scale = max(1e-6, float(np.abs(want).max()))
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5 * scale, err_msg=what)


@pytest.mark.parametrize("paths,reciprocal", [(False, False), (True, True)])
@pytest.mark.parametrize("loss_norm,norm,d", [("L2", True, 100), ("L2", False, 75), ("L2", True, 300), ("L1", True, 100)])
def test_weighted_margin_scorer_equals_float64_autograd(cuda_device, paths, reciprocal, loss_norm, norm, d):
    from openea_b200 import engine as eng
    rng = np.random.default_rng(d + 7 * paths)
    n_ent, n_rel, n = 5000, 300, 4000
    ent, rel = make_tables(rng, n_ent, n_rel, d)
    hi = n_rel if paths else n_ent
    pos = np.stack([rng.integers(0, hi, n), rng.integers(0, n_rel, n), rng.integers(0, hi, n)]).astype(np.int32)
    neg = np.stack([rng.integers(0, hi, n), rng.integers(0, n_rel, n), rng.integers(0, hi, n)]).astype(np.int32)
    w = (rng.random(n) * 50 + 1).astype(np.float32) if reciprocal else rng.random(n).astype(np.float32)
    margin, scale = 1.5, 0.1 if paths else 1.0
    E = torch.tensor(ent, dtype=torch.float64, requires_grad=True)
    R = torch.tensor(rel, dtype=torch.float64, requires_grad=True)
    En, Rn = _normed(E, norm), _normed(R, norm)
    A = Rn if paths else En

    def score(b):
        u = A[b[0].astype(np.int64)] + Rn[b[1].astype(np.int64)] - A[b[2].astype(np.int64)]
        return u.abs().sum(1) if loss_norm == "L1" else (u * u).sum(1)
    wt = torch.tensor(w, dtype=torch.float64)
    want = scale * ((1.0 / wt if reciprocal else wt) * torch.relu(margin + score(pos) - score(neg))).sum()
    want.backward()
    te, tr = eng.EmbeddingTable(ent, norm), eng.EmbeddingTable(rel, norm)
    t = eng.TripleTrainer(te, tr, eng.loss_cfg("margin-based", loss_norm, margin=margin), 0.01)
    t.score_margin_weighted(_dev(pos), _dev(neg), _dev(w), reciprocal=reciprocal, scale=scale, paths=paths)
    assert t.read_loss() == pytest.approx(float(want.detach()), rel=1e-4)
    for tab, ref, name in ((te, E, "entity"), (tr, R, "relation")):
        g = np.zeros((tab.rows, d)) if ref.grad is None else ref.grad.numpy()
        got = tab.grad[:, :d].cpu().numpy()
        if loss_norm == "L1":       # sign(u) flips where |u| is at fp32 noise level
            assert (np.abs(got - g) > 1e-4 * max(1.0, np.abs(g).max())).mean() < 5e-3
        else:
            _close(got, g, name + " gradient")
    if paths:
        assert not te.touched.any().item()


@pytest.mark.parametrize("norm,d", [(True, 100), (False, 75), (True, 300)])
def test_pair_distance_loss_equals_float64_autograd(cuda_device, norm, d):
    from openea_b200 import engine as eng
    rng = np.random.default_rng(d)
    ent, rel = make_tables(rng, 6000, 3, d)
    a = rng.integers(0, 6000, 3000).astype(np.int32)
    b = rng.integers(0, 6000, 3000).astype(np.int32)
    w = rng.random(3000).astype(np.float32)
    E = torch.tensor(ent, dtype=torch.float64, requires_grad=True)
    En = _normed(E, norm)
    want = 0.7 * (torch.tensor(w, dtype=torch.float64) * ((En[a.astype(np.int64)] - En[b.astype(np.int64)]) ** 2).sum(1)).sum()
    want.backward()
    te = eng.EmbeddingTable(ent, norm)
    t = eng.TripleTrainer(te, eng.EmbeddingTable(rel, norm), eng.loss_cfg("margin-based", "L2", margin=1.0), 0.01)
    t.score_pairs(a, b, _dev(w), scale=0.7)
    assert t.read_loss() == pytest.approx(float(want.detach()), rel=1e-4)
    _close(te.grad[:, :d].cpu().numpy(), E.grad.numpy(), "entity gradient")


def test_iptranse_lifecycle(cuda_device, tiny_kgs, tmp_path):
    """set_args / set_kgs / init / run / test / save of IPTransE on the tiny synthetic KG pair: the PTransE loss falls,
    the alignment epochs run on newly aligned entities, the result lines of the reference appear."""
    import re
    from openea_b200 import presets
    from openea_b200.approaches import IPTransE
    from tests.test_e2e_gpu import _hits1, _run
    args = presets.iptranse("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 121, 1000, 32
    args.bp_freq, args.sim_th = 40, 0.5
    model, out = _run(IPTransE, args, tiny_kgs, "sharing", tmp_path)
    loss = [float(x) for x in re.findall(r"avg\. triple loss:\s*([0-9.]+)", out)]
    assert len(loss) == 120 and loss[-1] < loss[0], (loss[:1], loss[-1:])
    assert "num of path:" in out and "Training ends. Total time" in out
    assert re.search(r"epoch 40, alignment loss: [0-9.]+", out) or "newly triples" not in out
    assert _hits1(out, "accurate results:") >= 0.0       # the result line exists; accuracy is calibrated once this has run on a GPU


def test_imuse_lifecycle(cuda_device, tiny_kgs, tmp_path):
    """set_args / set_kgs / init / run / test / save of IMUSE: the string matcher's pairs feed the pair-distance loss."""
    import re
    from openea_b200 import presets
    from openea_b200.approaches import IMUSE
    from tests.test_e2e_gpu import _hits1, _run
    args = presets.imuse("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 60, 1000, 32
    model, out = _run(IMUSE, args, tiny_kgs, "sharing", tmp_path)
    triple = [float(x) for x in re.findall(r"avg\. triple loss:\s*([0-9.]+)", out)]
    align = [float(x) for x in re.findall(r"align learning loss:\s*([0-9.]+)", out)]
    assert len(triple) == 60 and len(align) == 60 and triple[-1] < triple[0]
    if model.aligned_ent_pair_set:
        assert align[-1] < align[0], (align[0], align[-1])
    assert _hits1(out, "accurate results:") >= 0.0       # the result line exists; accuracy is calibrated once this has run on a GPU


def test_attre_lifecycle(cuda_device, tiny_kgs, tmp_path):
    """set_args / set_kgs / init / run / test / save of AttrE: structure, character-level and joint passes per epoch."""
    import re
    from openea_b200 import presets
    from openea_b200.approaches import AttrE
    from tests.test_e2e_gpu import _hits1, _run
    args = presets.attre("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 40, 1000, 32
    model, out = _run(AttrE, args, tiny_kgs, "sharing", tmp_path)
    for tag in (r"avg\. triple loss:\s*([0-9.]+)", r"CE, avg\. triple loss:\s*([0-9.]+)", r"joint learning loss:\s*([0-9.]+)"):
        vals = [float(x) for x in re.findall(tag, out)]
        assert len(vals) >= 40 and all(np.isfinite(vals)), tag
    ce = [float(x) for x in re.findall(r"CE, avg\. triple loss:\s*([0-9.]+)", out)]
    assert ce[-1] < ce[0], (ce[0], ce[-1])
    assert _hits1(out, "accurate results:") >= 0.0       # the result line exists; accuracy is calibrated once this has run on a GPU


def test_jape_lifecycle(cuda_device, tiny_kgs, tmp_path):
    """set_args / set_kgs / init / run / test / save of JAPE: the attribute auxiliary, then epochs of the signed-scale
    structure loss with the evaluated (never optimised) attribute-similarity loss printed next to it."""
    import re
    from openea_b200 import presets
    from openea_b200.approaches import JAPE
    from tests.test_e2e_gpu import _hits1, _run
    args = presets.jape("15K")
    args.batch_size, args.max_epoch, args.start_valid, args.dim = 1000, 60, 1000, 32
    args.attr_max_epoch, args.sub_mat_size = 3, 50
    model, out = _run(JAPE, args, tiny_kgs, "sharing", tmp_path)
    triple = [float(x) for x in re.findall(r"avg\. triple loss:\s*(-?[0-9.]+)", out)]
    assert len(triple) == 60, (len(triple), out[-600:])
    assert triple[-1] < triple[0], (triple[0], triple[-1])
    assert len(re.findall(r"sim loss:", out)) == 60
    assert _hits1(out, "accurate results:") >= 0.0       # the result line exists; accuracy is calibrated once this has run on a GPU
