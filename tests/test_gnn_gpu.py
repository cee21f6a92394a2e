"""GPU parity of path (ii): SpMM (K2), the L1 alignment loss and one full GCN-Align unit step against the
CPU oracle (torch-CPU autograd restatement, oracle/gnn.py).  fp32 tolerance 1e-4 relative."""
import contextlib
import io

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gnn as orc

pytestmark = pytest.mark.gpu


def _rand_csr(rng, n_rows, n_cols, avg, hubs=3):
    rows = rng.integers(0, n_rows, size=avg * n_rows)
    cols = rng.integers(0, n_cols, size=avg * n_rows)
    # a few hub rows with > 256 non-zeros exercise the CTA-per-row path; also leave some rows empty
    for h in range(hubs):
        extra = rng.integers(0, n_cols, size=700 + 300 * h)
        rows = np.concatenate([rows, np.full(extra.size, h * 7 % n_rows)])
        cols = np.concatenate([cols, extra])
    keep = rows % 11 != 5
    vals = rng.standard_normal(rows.size).astype(np.float32)
    return sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n_rows, n_cols))


@pytest.mark.parametrize("d", [100, 200, 300, 500])
def test_spmm_matches_scipy(cuda_device, d):
    from openea_b200 import gnn
    rng = np.random.default_rng(d)
    A = _rand_csr(rng, 3000, 2500, 6)
    X = rng.standard_normal((2500, d)).astype(np.float32)
    dA = gnn.DeviceCsr(A)
    assert dA.long_rows.numel() >= 1
    Xd = torch.from_numpy(X).cuda()
    want = A @ X
    got = gnn.spmm(dA, Xd).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    got = gnn.spmm(dA, Xd, relu=True).cpu().numpy()
    np.testing.assert_allclose(got, np.maximum(want, 0), rtol=1e-4, atol=1e-4)
    mask = rng.standard_normal(want.shape).astype(np.float32)
    got = gnn.spmm(dA, Xd, mask_src=torch.from_numpy(mask).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, np.where(mask > 0, want, 0), rtol=1e-4, atol=1e-4)
    y0 = rng.standard_normal(want.shape).astype(np.float32)
    out = torch.from_numpy(y0.copy()).cuda()
    gnn.spmm(dA, Xd, out=out, beta=0.5)
    np.testing.assert_allclose(out.cpu().numpy(), want + 0.5 * y0, rtol=1e-4, atol=1e-4)
    # transpose (the backward operand)
    G = rng.standard_normal((3000, d)).astype(np.float32)
    got = gnn.spmm(dA.transpose(), torch.from_numpy(G).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, A.T @ G, rtol=1e-4, atol=2e-4)


def test_spmm_empty_rows_and_single_row(cuda_device):
    from openea_b200 import gnn
    A = sp.csr_matrix((np.array([2.0], dtype=np.float32), (np.array([3]), np.array([1]))), shape=(5, 4))
    X = torch.arange(16, dtype=torch.float32).reshape(4, 4).cuda()
    got = gnn.spmm(gnn.DeviceCsr(A), X).cpu().numpy()
    want = np.zeros((5, 4), dtype=np.float32); want[3] = 2 * np.arange(4, 8)
    np.testing.assert_array_equal(got, want)


def _negs(rng, ill, n, k):
    t = len(ill)
    return (np.repeat(ill[:, 0], k).astype(np.int32), rng.integers(0, n, t * k).astype(np.int32),
            rng.integers(0, n, t * k).astype(np.int32), np.repeat(ill[:, 1], k).astype(np.int32))


def test_align_loss_l1_matches_autograd(cuda_device):
    from openea_b200 import gnn
    rng = np.random.default_rng(1)
    n, d, t, k, gamma = 800, 100, 120, 5, 3.0
    x = rng.standard_normal((n, d)).astype(np.float32)
    ill = np.stack([rng.choice(n, t, replace=False), rng.choice(n, t, replace=False)], 1)
    negs = _negs(rng, ill, n, k)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    want = orc.align_loss(xt, ill, gamma, k, *negs)
    want.backward()
    xd = torch.from_numpy(x).cuda()
    grad = torch.zeros_like(xd)
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    gnn.align_loss_l1(xd, d, dev(ill[:, 0].astype(np.int32)), dev(ill[:, 1].astype(np.int32)), k, *[dev(a) for a in negs],
                      gamma, grad, loss)
    assert float(loss.item()) == pytest.approx(float(want.detach()), rel=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), xt.grad.numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("with_features", [False, True])
def test_gcn_unit_step_matches_oracle(cuda_device, with_features):
    """One session.run([loss, opt_op]) of a GCN_Align_Unit (SE: featureless, AE: sparse features)."""
    from openea_b200 import gnn
    from openea_b200.approaches.gcn_align import GCNAlignUnit
    from openea_b200.engine import EmbeddingTable
    from openea_b200.synth import synth_id_arrays
    arr = synth_id_arrays("tiny", swapping=False)
    n, d, k, gamma, lr = arr["n_ent"], 100, 5, 3.0, 8.0
    triples = np.concatenate([arr["triples1"], arr["triples2"]])
    support = gnn.preprocess_adj(gnn.weighted_adjacency(n, triples))
    rng = np.random.default_rng(3)
    feats = None
    rows = n
    if with_features:
        feats = sp.csr_matrix((rng.random((n, 40)) < 0.1).astype(np.float32))
        rows = 40
    W0 = (rng.standard_normal((rows, d)) / np.sqrt(rows)).astype(np.float32)
    ill = arr["train_links"].astype(np.int64)
    negs = _negs(rng, ill, n, k)
    want_loss, want_W, want_out = orc.unit_train_step(support, W0, feats, ill, gamma, k, negs, lr)

    table = EmbeddingTable(W0, True, "SGD")
    unit = GCNAlignUnit(gnn.DeviceCsr(support), table, None if feats is None else gnn.DeviceCsr(feats), ill, gamma, k, lr)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    loss = unit.train_step(*[dev(a) for a in negs])
    assert float(loss.item()) == pytest.approx(want_loss, rel=1e-4)
    np.testing.assert_allclose(unit.outputs[:, :d].cpu().numpy(), want_out, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(table.raw().cpu().numpy(), want_W, rtol=1e-4, atol=2e-6 * np.abs(want_W).max() + 1e-7)


def test_gcn_align_lifecycle(cuda_device, tmp_path):
    import os
    import re
    from openea_b200 import presets
    from openea_b200.approaches import GCN_Align
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    folder = str(tmp_path) + "/data/"
    write_dataset(folder, "tiny")
    args = presets.gcn_align()
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.max_epoch, args.start_valid, args.se_dim, args.ae_dim = 60, 30, 64, 32
    np.random.seed(0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kgs = read_kgs_from_folder(folder, args.dataset_division, "mapping", True)
        m = GCN_Align(); m.set_args(args); m.set_kgs(kgs); m.init(); m.run(); m.test(); m.save()
    out = buf.getvalue()
    assert "avg. relation triple loss" in out
    h1 = float(re.findall(r"accurate results: hits@\[1, 5, 10, 50\] = \[\s*([0-9.]+)", out)[-1])
    assert h1 > 5.0, h1          # chance = 0.24 %
    assert os.path.exists(m.out_folder + "ent_embeds.npy") and os.path.exists(m.out_folder + "attr_embeds.npy")


def _small_graph(rng, n, avg):
    rows = rng.integers(0, n, n * avg); cols = rng.integers(0, n, n * avg)
    hub = rng.integers(0, n, 600)
    rows = np.concatenate([rows, np.zeros(600, dtype=np.int64), np.arange(n)]); cols = np.concatenate([cols, hub, np.arange(n)])
    m = sp.coo_matrix((rng.random(rows.size) + 0.1, (rows, cols)), shape=(n, n)).tocsr()
    return m


def test_gat_aggregate_fwd_bwd_matches_autograd(cuda_device):
    """Edge-softmax attention aggregation (alinet.py:656-677): kernels vs a float64 torch scatter restatement."""
    from openea_b200 import gnn
    rng = np.random.default_rng(9)
    n, d = 700, 100
    adj = _small_graph(rng, n, 5)
    A = gnn.DeviceCsr(adj)
    assert A.long_rows.numel() >= 1
    s1h, s2h = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    Mh, Gh = rng.standard_normal((n, d)).astype(np.float32), rng.standard_normal((n, d)).astype(np.float32)
    s1 = torch.tensor(s1h, device="cuda", requires_grad=True); s2 = torch.tensor(s2h, device="cuda", requires_grad=True)
    M = torch.tensor(Mh, device="cuda", requires_grad=True)
    out = gnn.GatAggregateFn.apply(s1, s2, M, A, 0.2)
    (out * torch.tensor(Gh, device="cuda")).sum().backward()
    o1 = torch.tensor(s1h, dtype=torch.float64, requires_grad=True); o2 = torch.tensor(s2h, dtype=torch.float64, requires_grad=True)
    oM = torch.tensor(Mh, dtype=torch.float64, requires_grad=True)
    want = orc.edge_softmax_aggregate(sp.csr_matrix(adj, dtype=np.float32), o1, o2, oM)
    (want * torch.tensor(Gh, dtype=torch.float64)).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), want.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(M.grad.cpu().numpy(), oM.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(s1.grad.cpu().numpy(), o1.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(s2.grad.cpu().numpy(), o2.grad.numpy(), rtol=2e-4, atol=2e-5)


def test_alinet_step_matches_oracle(cuda_device):
    """One session.run([loss, optimizer]) of the AliNet graph: loss, gradients and the TF-Adam update."""
    from openea_b200 import gnn
    from openea_b200.approaches.alinet import AliNetModel, DenseAdam
    rng = np.random.default_rng(4)
    n, dims = 600, [64, 48, 32]
    adj1 = gnn.normalize_adj(_small_graph(rng, n, 4)).tocsr()
    adj2 = gnn.normalize_adj(_small_graph(rng, n, 6)).tocsr()
    model = AliNetModel(n, dims, gnn.DeviceCsr(adj1), gnn.DeviceCsr(adj2), torch.device("cuda"), seed=1)
    pos = np.stack([rng.integers(0, n, 80), rng.integers(0, n, 80)], 1)
    neg = np.stack([rng.integers(0, n, 600), rng.integers(0, n, 600)], 1)
    rel_win = 5
    hs, ts = rng.integers(0, n, 7 * rel_win), rng.integers(0, n, 7 * rel_win)
    tl = lambda a: torch.as_tensor(a, dtype=torch.long, device="cuda")
    before = {k: v.detach().cpu().numpy().copy() for k, v in model.params.items()}
    oparams = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in before.items()}
    oouts = orc.alinet_forward(oparams, adj1, adj2, 2)
    oloss = orc.alinet_loss(oparams, oouts, pos, neg, 1.5, 0.1, hs, ts, rel_win, 0.01)
    oloss.backward()
    opt = DenseAdam(list(model.params.values()), 0.001)
    outs = model.forward()
    loss = model.loss(outs, tl(pos), tl(neg), 1.5, 0.1, tl(hs), tl(ts), rel_win, 0.01)
    assert float(loss.detach().item()) == pytest.approx(float(oloss.detach()), rel=1e-4)
    loss.backward()
    for k in ("init_embedding", "gcn0.kernel", "gat0.kernel1", "hw0.kernel", "gcn1.bias", "gat0.bn_gamma"):
        g, w = model.params[k].grad.cpu().numpy(), oparams[k].grad.numpy()
        np.testing.assert_allclose(g, w, rtol=2e-3, atol=2e-5 * max(1e-6, np.abs(w).max()), err_msg=k)
    opt.step()
    # TF Adam, t = 1: lr_t = lr·√(1−β2)/(1−β1); m = (1−β1) g; v = (1−β2) g² → x −= lr_t·m/(√v + ε)
    lr_t = 0.001 * np.sqrt(1 - 0.999) / (1 - 0.9)
    for k in ("init_embedding", "gcn0.kernel"):
        g = oparams[k].grad.numpy()
        want = before[k] - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
        np.testing.assert_allclose(model.params[k].detach().cpu().numpy(), want, rtol=1e-3, atol=2e-6)


def test_alinet_lifecycle(cuda_device, tmp_path):
    import os
    import re
    from openea_b200 import presets
    from openea_b200.approaches import AliNet
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    folder = str(tmp_path) + "/data/"
    write_dataset(folder, "tiny")
    args = presets.alinet()
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.layer_dims, args.batch_size, args.max_epoch, args.start_valid, args.eval_freq = [64, 48, 32], 200, 60, 20, 20
    args.truncated_epsilon, args.min_rel_win = 0.9, 5
    np.random.seed(0)
    import random
    random.seed(0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kgs = read_kgs_from_folder(folder, args.dataset_division, "mapping", True)
        m = AliNet(); m.set_args(args); m.set_kgs(kgs); m.init(); m.run(); m.test(); m.save()
    out = buf.getvalue()
    assert "epoch 60, loss:" in out and "neighbors num" in out
    h1 = float(re.findall(r"accurate results: hits@\[1, 5, 10, 50\] = \[\s*([0-9.]+)", out)[-1])
    assert h1 > 3.0, h1          # chance = 0.24 %
    assert os.path.exists(m.out_folder + "ent_embeds.npy")
    assert np.load(m.out_folder + "ent_embeds.npy").shape[1] == 64 + 48 + 32


class _FakeKgs:
    """The attributes RDGCNLayer reads from a KGs object."""

    def __init__(self, arr):
        K = type("K", (), {})
        self.kg1, self.kg2 = K(), K()
        self.kg1.relation_triples_list = [tuple(x) for x in arr["triples1"].tolist()]
        self.kg2.relation_triples_list = [tuple(x) for x in arr["triples2"].tolist()]
        self.entities_num, self.relations_num = arr["n_ent"], arr["n_rel"]
        self.train_links = [tuple(x) for x in arr["train_links"].tolist()]


def test_rdgcn_step_matches_oracle(cuda_device):
    """One session.run([optimizer, loss]) of the RDGCN graph (rdgcn.py:317-338): outputs, loss, gradients."""
    from openea_b200.approaches import rdgcn as R
    from openea_b200.modules.args.args_hander import ARGs
    from openea_b200.synth import synth_id_arrays
    arr = synth_id_arrays("tiny", swapping=False)
    kgs = _FakeKgs(arr)
    rng = np.random.default_rng(2)
    d, k = 32, 5
    args = ARGs(dict(dim=d, alpha=0.1, beta=0.3, gamma=1.0, neg_triple_num=k))
    emb = rng.standard_normal((arr["n_ent"], d)).astype(np.float32)
    layer = R.RDGCNLayer(args, kgs, emb, torch.device("cuda"), seed=3)
    for name in ("sp1", "sp2", "self.f1", "dual.f2"):     # make the attention parameters non-trivial
        with torch.no_grad():
            layer.params[name + ".b"][:, 0] = 0.3
    ill = np.array(kgs.train_links)
    t = len(ill)
    negs_h = (np.repeat(ill[:, 0], k), rng.integers(0, arr["n_ent"], t * k), rng.integers(0, arr["n_ent"], t * k), np.repeat(ill[:, 1], k))
    negs = tuple(torch.as_tensor(a.astype(np.int32), device="cuda") for a in negs_h)
    out = layer.forward()
    loss = layer.loss(out, negs)
    loss.backward()
    # oracle: same parameters in float64 on the CPU
    triples = kgs.kg1.relation_triples_list + kgs.kg2.relation_triples_list
    M = R.get_sparse_matrix(triples, arr["n_ent"])
    head_r, tail_r, tri = R.rfunc(triples, arr["n_ent"], arr["n_rel"])
    norm = lambda m: sp.diags(1.0 / np.maximum(np.asarray(m.sum(1)).reshape(-1), 1e-30)) @ m
    P = {kk: torch.tensor(v.detach().cpu().numpy(), dtype=torch.float64, requires_grad=True) for kk, v in layer.params.items()}
    oout = orc.rdgcn_forward(P, M, norm(head_r), norm(tail_r), R.dual_adjacency(head_r, tail_r), tri, 0.1, 0.3)
    oloss = orc.align_loss(oout, ill, 1.0, k, *negs_h)
    oloss.backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), oout.detach().numpy(), rtol=2e-4, atol=2e-5)
    assert float(loss.detach().item()) == pytest.approx(float(oloss.detach()), rel=1e-4)
    for name in ("X0", "hw1.W", "diag2.w", "self.W", "dual.W"):
        g, w = layer.params[name].grad.cpu().numpy(), P[name].grad.numpy()
        np.testing.assert_allclose(g, w, rtol=5e-3, atol=5e-5 * max(1e-9, np.abs(w).max()), err_msg=name)
    for name in ("sp1", "sp2", "self.f1"):
        g, w = layer.params[name + ".w"].grad.cpu().numpy()[:, 0], P[name + ".w"].grad.numpy()[:, 0]
        np.testing.assert_allclose(g, w, rtol=5e-3, atol=5e-5 * max(1e-9, np.abs(w).max()), err_msg=name)


def test_rdgcn_lifecycle(cuda_device, tmp_path):
    import os
    import re
    from openea_b200 import presets
    from openea_b200.approaches import RDGCN
    from openea_b200.modules.load.kgs import read_kgs_from_folder
    from openea_b200.synth import write_dataset
    folder = str(tmp_path) + "/data/"
    write_dataset(folder, "tiny")
    args = presets.rdgcn()
    args.training_data, args.output = folder, str(tmp_path) + "/out/"
    args.dim, args.max_epoch, args.start_valid, args.neg_triple_num = 64, 40, 20, 40     # k > 32: block + radix-select path
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kgs = read_kgs_from_folder(folder, args.dataset_division, "mapping", True)
        m = RDGCN(); m.set_args(args); m.set_kgs(kgs); m.init(); m.run(); m.test(); m.save()
    out = buf.getvalue()
    losses = [float(x) for x in re.findall(r"avg. relation triple loss: ([0-9.]+)", out)]
    assert len(losses) == 40 and losses[-1] < losses[0]
    h1 = float(re.findall(r"accurate results: hits@\[1, 5, 10, 50\] = \[\s*([0-9.]+)", out)[-1])
    assert h1 > 30.0, h1          # name vectors carry most of the signal (as in the reference)
    assert np.load(m.out_folder + "ent_embeds.npy").shape == (kgs.entities_num, 64)
    # with RDGCN's own setting at the 100K scale (k = 10) the hard negatives come from the fused L1 top-k kernel
    args.neg_triple_num, args.max_epoch = 10, 12
    with contextlib.redirect_stdout(io.StringIO()):
        m2 = RDGCN(); m2.set_args(args); m2.set_kgs(kgs); m2.init(); m2.run()
