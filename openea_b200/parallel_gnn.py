"""Path (ii) across GPUs: the adjacency A, the layer outputs and — for the featureless first layer —
the entity table itself are ROW-SHARDED; one exchange per layer.

Partition: contiguous blocks of `block = ceil(N / G)` rows; rank r owns rows [r·block, min((r+1)·block, N)); every
per-row tensor is padded to `block` rows so the collectives are fixed-size (`N_pad = G·block`; padding rows of A are
empty and padding columns are never referenced).

One GCN-Align unit step (approaches/gcn_align.py:239-267,298-320,498-539), rank-local work in liboea kernels:

    forward   Ŵ_r = lookup(table shard)                                  [block, p]
              pre = all-gather(Ŵ_r)              (featureless SE unit)   [N_pad, p]     ← exchange 1
                    or X_r·Ŵ (AE unit: the small [n_attr, d] table is replicated), then all-gather
              H1_r = relu(A_r · pre)             oea_spmm_csr            [block, p]
              H1   = all-gather(H1_r)                                                  ← exchange 2
              OUT_r = A_r · H1 ; OUT = all-gather(OUT_r)                                ← exchange 3
    loss      seed pairs are block-sharded too: every rank runs oea_align_loss_l1 on its t_r pairs over OUT
              (negatives index all N rows) → loss_r and a partial dOUT [N_pad, p], both weighted t_r / t
              dOUT_r = reduce-scatter(partial dOUT)                                     ← exchange 4
    backward  dH1_r = reduce-scatter(relu'(H1) ∘ (A_rᵀ · dOUT_r))                       ← exchange 5
              dpre_r = reduce-scatter(A_rᵀ · dH1_r)                                     ← exchange 6
              SE: scatter dpre_r through the normalisation of the owned rows, SGD on the shard
              AE: all-reduce(X_rᵀ · dpre_r) [n_attr, p], identical SGD step on every replica

Each exchange moves N_pad·p·4 B (12 MB at the 15K shape with d = 100, 80 MB at 100K) over NVLink.  The result equals
the single-GPU unit up to fp32 summation order (the loss is a sum over pairs, the adjacency products are exact row
partitions).  NCCL on GPUs; under gloo (CPU tests) reduce-scatter is an all-reduce + slice.

`ops` abstracts the rank-local numerics: `KernelOps` (the default, liboea.so kernels — the product path) or a
test-provided stand-in so that the collective algebra can be checked with gloo where no GPU exists
(tests/test_parallel_gnn_gloo.py).
"""
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from . import parallel as par


class RowShard:
    """Contiguous, padded row partition of n rows over the process group."""

    def __init__(self, n, rank=None, world_size=None):
        r, w = par.world()
        self.rank = r if rank is None else rank
        self.world = w if world_size is None else world_size
        self.n = int(n)
        self.block = -(-self.n // self.world)
        self.n_pad = self.block * self.world
        self.lo = min(self.rank * self.block, self.n)
        self.hi = min(self.lo + self.block, self.n)

    def rows_of(self, mat):
        """Rows [lo, hi) of a scipy matrix as CSR [block, n_pad-or-original columns] (empty padding rows)."""
        m = sp.csr_matrix(mat)[self.lo:self.hi]
        if m.shape[0] < self.block:
            m = sp.vstack([m, sp.csr_matrix((self.block - m.shape[0], m.shape[1]), dtype=m.dtype)]).tocsr()
        return m

    def square_rows_of(self, mat):
        """Rows [lo, hi) of a square [n, n] matrix, columns padded to n_pad (the layer input is all-gathered)."""
        m = self.rows_of(mat)
        if m.shape[1] < self.n_pad:
            m = sp.hstack([m, sp.csr_matrix((m.shape[0], self.n_pad - m.shape[1]), dtype=m.dtype)]).tocsr()
        return m

    def local_rows(self, full):
        """Rows [lo, hi) of a host array padded with zero rows to `block`."""
        full = np.asarray(full)
        out = np.zeros((self.block,) + full.shape[1:], dtype=full.dtype)
        out[:self.hi - self.lo] = full[self.lo:self.hi]
        return out


def all_gather_rows(x_local, shard):
    """[block, p] per rank → [n_pad, p] on every rank."""
    if shard.world == 1:
        return x_local
    out = torch.empty((shard.n_pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, x_local.contiguous())
    return out


def reduce_scatter_rows(x_partial, shard):
    """Σ over ranks of [n_pad, p] partial results, each rank keeping its [block, p] rows."""
    if shard.world == 1:
        return x_partial
    x_partial = x_partial.contiguous()
    if dist.get_backend() == "gloo":            # gloo has no reduce-scatter
        dist.all_reduce(x_partial)
        return x_partial[shard.rank * shard.block:(shard.rank + 1) * shard.block].clone()
    out = torch.empty((shard.block,) + tuple(x_partial.shape[1:]), dtype=x_partial.dtype, device=x_partial.device)
    dist.reduce_scatter_tensor(out, x_partial)
    return out


class KernelOps:
    """Rank-local numerics through liboea.so (the product path; raises when the library or the GPU is missing)."""

    def csr(self, mat, device, keep_duplicates=False):
        from . import gnn
        return gnn.DeviceCsr(mat, device, keep_duplicates=keep_duplicates)

    def spmm(self, A, X, relu=False, mask_src=None):
        from . import gnn
        return gnn.spmm(A, X, relu=relu, mask_src=mask_src)

    def lookup(self, table):
        return table.lookup(padded=True)

    def update(self, table, grad_rows, lr):
        table.scatter_grad(grad_rows)          # through the row normalisation (path (i) table kernels)
        table.apply(lr)                        # GradientDescentOptimizer (gcn_align.py:511)

    def align_loss(self, x, dim, left, right, k, negs, gamma, grad, loss_out):
        from . import gnn
        gnn.align_loss_l1(x, dim, left, right, k, negs[0], negs[1], negs[2], negs[3], gamma, grad, loss_out)


class ShardedGCNAlignUnit:
    """GCNAlignUnit of approaches/gcn_align.py with rows sharded over the process group.

    support  : scipy [N, N] normalised adjacency (every rank passes the same matrix, keeps its row block)
    table    : featureless unit → EmbeddingTable over THIS rank's `block` rows of the [N, d] entity table
               (`shard.local_rows(global_init)`); feature unit → the replicated [n_attr, d] table
    features : scipy [N, n_attr] or None
    ill      : [t, 2] seed pairs (global entity ids), identical on every rank
    """

    def __init__(self, support, table, features, ill, gamma, k, lr, shard=None, ops=None):
        self.ops = ops or KernelOps()
        self.shard = shard or RowShard(support.shape[0])
        sh, dev = self.shard, table.device
        self.n = support.shape[0]
        self.A = self.ops.csr(sh.square_rows_of(support), dev)          # [block, n_pad]
        self.At = self.A.transpose()                                    # [n_pad, block]
        self.table = table
        self.X = self.ops.csr(sh.rows_of(features), dev) if features is not None else None    # [block, n_attr]
        self.Xt = self.X.transpose() if self.X is not None else None
        ill = np.asarray(ill)
        self.t = len(ill)
        self.p_lo, self.p_hi = par.block_range(self.t, sh.rank, sh.world)                     # my seed pairs
        mine = ill[self.p_lo:self.p_hi]
        self.left = torch.as_tensor(mine[:, 0], dtype=torch.int32, device=dev).contiguous()
        self.right = torch.as_tensor(mine[:, 1], dtype=torch.int32, device=dev).contiguous()
        self.gamma, self.k, self.lr = float(gamma), int(k), float(lr)
        self.loss_dev = torch.zeros(1, dtype=torch.float64, device=dev)
        self.outputs = None            # [N, p]: all rows on every rank (evaluation reads arbitrary rows)
        self._h1 = self._out_full = None

    def forward(self):
        sh = self.shard
        wn = self.ops.lookup(self.table)
        pre_local = wn if self.X is None else self.ops.spmm(self.X, wn)
        pre = all_gather_rows(pre_local, sh)
        self._h1 = all_gather_rows(self.ops.spmm(self.A, pre, relu=True), sh)
        self._out_full = all_gather_rows(self.ops.spmm(self.A, self._h1), sh)
        self.outputs = self._out_full[:self.n]
        return self.outputs

    def _my_negs(self, negs):
        lo, hi = self.p_lo * self.k, self.p_hi * self.k
        return [n[lo:hi].contiguous() for n in negs]

    def train_step(self, neg_left, neg_right, neg2_left, neg2_right):
        """One session.run([loss, opt_op]) of the unit; returns the global loss (device fp64 scalar tensor)."""
        sh = self.shard
        self.forward()
        out = self._out_full
        g_out = torch.zeros_like(out)
        self.loss_dev.zero_()
        t_mine = self.p_hi - self.p_lo
        if t_mine > 0:
            self.ops.align_loss(out, self.table.dim, self.left, self.right, self.k,
                                self._my_negs((neg_left, neg_right, neg2_left, neg2_right)), self.gamma, g_out,
                                self.loss_dev)
        w = t_mine / float(self.t)           # the kernel averages over ITS pairs; the unit's loss averages over all t
        self.loss_dev *= w
        g_out *= w
        if sh.world > 1:
            dist.all_reduce(self.loss_dev)
        g_out_local = reduce_scatter_rows(g_out, sh)
        g_h1_local = reduce_scatter_rows(self.ops.spmm(self.At, g_out_local, mask_src=self._h1), sh)
        g_pre_local = reduce_scatter_rows(self.ops.spmm(self.At, g_h1_local), sh)
        if self.X is None:
            g_w = g_pre_local                                   # gradient of the owned rows of Ŵ
        else:
            g_w = self.ops.spmm(self.Xt, g_pre_local)           # [n_attr, p] partial over my rows
            if sh.world > 1:
                dist.all_reduce(g_w)
        self.ops.update(self.table, g_w, self.lr)
        return self.loss_dev


# ---- AliNet (approaches/alinet.py:539-677,784-866) across GPUs ---------------------------------------------------------------
class AllGatherRows(torch.autograd.Function):
    """[block, p] per rank → [n_pad, p] on every rank, inside an autograd graph.

    The backward depends on who consumes the gathered rows:
      partitioned consumer (each rank aggregates into ITS rows, e.g. A_r · X): the gradients of the ranks are partial
          sums over disjoint outputs → reduce-scatter;
      replicated consumer (every rank computes the same thing from the full tensor, e.g. the loss on a batch of
          pairs): every rank already holds the complete gradient → keep the own rows, no communication."""

    @staticmethod
    def forward(ctx, x_local, shard, replicated_consumer):
        ctx.shard, ctx.replicated = shard, replicated_consumer
        return all_gather_rows(x_local.contiguous(), shard)

    @staticmethod
    def backward(ctx, g_full):
        sh = ctx.shard
        if ctx.replicated or sh.world == 1:
            return g_full[sh.rank * sh.block:(sh.rank + 1) * sh.block].contiguous(), None, None
        return reduce_scatter_rows(g_full, sh), None, None


def make_sharded_alinet(base_cls):
    """ShardedAliNetModel over the (late-imported) AliNetModel: same parameters, same forward code, with
      * the input embedding table, both adjacencies and every layer's rows sharded by `RowShard`;
      * an all-gather in front of every aggregation (partitioned consumer) and after the last layer op of each block
        (replicated consumer: loss, evaluation and neighbour search run unchanged on all rows on every rank);
      * the small dense weights replicated, their partial gradients all-reduced by `sync_grads()` before the optimiser.
    Exchanges per forward with L = 2 aggregation layers: 1 (X·W) + 2 (attention scores, mapped rows) per layer + one
    gather per layer output + the input table = 8 all-gathers of N_pad·d·4 B."""

    class ShardedAliNetModel(base_cls):

        def __init__(self, n_ent, layer_dims, adj1, adj2, device, seed=0, shard=None, ops=None):
            self.shard = shard or RowShard(n_ent)
            self.ops = ops or KernelOps()
            self.n_ent = n_ent
            sh = self.shard
            super().__init__(n_ent, layer_dims, self.ops.csr(sh.square_rows_of(adj1), device),
                             self.ops.csr(sh.square_rows_of(adj2), device), device, seed=seed)
            # every rank drew the same full table from the same generator: keep the owned rows only
            full = self.params["init_embedding"].detach()
            local = torch.zeros(sh.block, full.shape[1], dtype=full.dtype, device=full.device)
            local[:sh.hi - sh.lo] = full[sh.lo:sh.hi]
            self.params["init_embedding"] = local.requires_grad_(True)

        def _layer_input(self, x):
            return AllGatherRows.apply(x, self.shard, False)

        def _layer_outputs(self, outs):
            return [AllGatherRows.apply(o, self.shard, True)[:self.n_ent] for o in outs]

        def input_embedding(self):
            return AllGatherRows.apply(self.params["init_embedding"], self.shard, True)[:self.n_ent]

        def set_adj1(self, mat, device):
            self.adj1 = self.ops.csr(self.shard.square_rows_of(mat), device)

        def sync_grads(self):
            """Sum the partial gradients of the replicated parameters (everything but the sharded input table)."""
            if self.shard.world == 1:
                return
            for name, p in self.params.items():
                if name != "init_embedding" and p.grad is not None:
                    dist.all_reduce(p.grad)

    return ShardedAliNetModel


# ---- RDGCN (approaches/rdgcn.py:184-338) across GPUs ------------------------------------------------------------------------
class AllReduceSum(torch.autograd.Function):
    """Σ over ranks of partial results of a reduction over sharded rows, inside an autograd graph.  Every rank's part
    enters the sum with weight one, and what consumes the sum on each rank contributes a partial gradient (its own
    rows' share), so the backward is the same all-reduce."""

    @staticmethod
    def forward(ctx, x):
        out = x.clone()
        dist.all_reduce(out)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g)
        return g


def row_block_keep_duplicates(mat, shard):
    """Rows [lo, hi) of a CSR matrix as [block, n_pad] WITHOUT merging duplicate entries (RDGCN's r_mat holds one entry
    per triple, several of them at the same (h, t))."""
    m = sp.csr_matrix(mat) if not sp.isspmatrix_csr(mat) else mat
    a, b = m.indptr[shard.lo], m.indptr[shard.hi]
    indptr = np.full(shard.block + 1, b - a, dtype=m.indptr.dtype)
    indptr[:shard.hi - shard.lo + 1] = m.indptr[shard.lo:shard.hi + 1] - a
    return sp.csr_matrix((m.data[a:b], m.indices[a:b], indptr), shape=(shard.block, shard.n_pad))


def make_sharded_rdgcn(base_cls):
    """ShardedRDGCNLayer over the (late-imported) RDGCNLayer: the entity rows of X0, of every intermediate and of the
    entity-side matrices (M, r_mat) are sharded; the relation-side tensors (dual input / attention, [R, ·] with R ≈ 500)
    are replicated.  Exchanges per forward: 2 all-reduces of [R, 2d] (the relation-incidence averages reduce over all
    entity rows, each rank holding its column block of head_r / tail_r), 4 all-gathers of [N_pad, d] in front of the
    entity-row aggregations, 1 all-gather of the output (replicated consumers: loss, hard-negative search, evaluation)."""

    class ShardedRDGCNLayer(base_cls):

        def __init__(self, args, kgs, embedding, device, seed=0, shard=None, ops=None):
            self.shard = shard or RowShard(kgs.entities_num)
            self.ops = ops or KernelOps()
            self.n_ent = kgs.entities_num
            super().__init__(args, kgs, embedding, device, seed=seed)

        def _entity_rows(self, mat, device, keep_duplicates=False):
            sh = self.shard
            if keep_duplicates:
                return self.ops.csr(row_block_keep_duplicates(mat, sh), device, keep_duplicates=True)
            return self.ops.csr(sh.square_rows_of(mat), device)

        def _entity_cols(self, mat, device):
            sh = self.shard
            m = sp.csc_matrix(mat)[:, sh.lo:sh.hi]
            if m.shape[1] < sh.block:
                m = sp.hstack([m, sp.csc_matrix((m.shape[0], sh.block - m.shape[1]), dtype=m.dtype)])
            return self.ops.csr(sp.csr_matrix(m), device)

        def _own_rows(self, x):
            sh = self.shard
            local = torch.zeros(sh.block, x.shape[1], dtype=x.dtype)
            local[:sh.hi - sh.lo] = x[sh.lo:sh.hi]
            return local

        def _all_rows(self, x):
            return AllGatherRows.apply(x, self.shard, False)

        def _sum_over_owners(self, x):
            return AllReduceSum.apply(x) if self.shard.world > 1 else x

        def _full_output(self, x):
            return AllGatherRows.apply(x, self.shard, True)[:self.n_ent]

        def sync_grads(self):
            """Sum the partial gradients of the replicated parameters (everything but the sharded X0)."""
            if self.shard.world == 1:
                return
            for name, p in self.params.items():
                if name != "X0" and p.grad is not None:
                    dist.all_reduce(p.grad)

    return ShardedRDGCNLayer


def __getattr__(name):            # built on first use: the approach modules import this module's siblings
    if name == "ShardedAliNetModel":
        from .approaches.alinet import AliNetModel
        cls = make_sharded_alinet(AliNetModel)
    elif name == "ShardedRDGCNLayer":
        from .approaches.rdgcn import RDGCNLayer
        cls = make_sharded_rdgcn(RDGCNLayer)
    else:
        raise AttributeError(name)
    globals()[name] = cls
    return cls
