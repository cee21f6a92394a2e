"""Path (ii) across GPUs: the adjacency A, the layer outputs and — for the featureless first layer —
the entity table itself are ROW-SHARDED; one exchange per layer.

Partition: CYCLIC — rank r owns the ids r, r + G, … (ids are ordered by descending frequency, so this balances the
non-zeros); every per-row tensor is padded to `block = ceil(N / G)` rows so the collectives are fixed-size
(`N_pad = G·block`; padding rows of A are empty and padding columns are never referenced); gathered tensors are in
rank-after-rank "position" order and adjacency columns are renumbered to positions once (`RowShard`).

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
    """Cyclic, padded row partition of n rows over the process group: rank r owns the ids r, r + G, r + 2G, …

    Entity ids interleave the two KGs by DESCENDING frequency (read.py:69-79), so contiguous blocks would give rank 0
    every hub and most of the adjacency's non-zeros; cyclic ownership gives every rank the same degree profile.
    Internally a gathered tensor is laid out rank after rank ("position" order): id i sits at position
    pos_of_id[i] = (i mod G)·block + i div G, each rank's `block = ceil(n / G)` rows padded at the end.  Adjacency
    columns are renumbered to positions once at build time; `to_global` returns a gathered tensor in id order."""

    def __init__(self, n, rank=None, world_size=None):
        r, w = par.world()
        self.rank = r if rank is None else rank
        self.world = w if world_size is None else world_size
        self.n = int(n)
        self.block = -(-self.n // self.world)
        self.n_pad = self.block * self.world
        ids = np.arange(self.n, dtype=np.int64)
        self.pos_of_id = (ids % self.world) * self.block + ids // self.world
        self.my_ids = np.arange(self.rank, self.n, self.world, dtype=np.int64)
        self.n_local = len(self.my_ids)
        self._pos_t = {}

    def pos_tensor(self, device):
        """pos_of_id as an int64 tensor on `device` (cached)."""
        key = str(device)
        if key not in self._pos_t:
            self._pos_t[key] = torch.as_tensor(self.pos_of_id, device=device)
        return self._pos_t[key]

    def to_global(self, x_full):
        """[n_pad, …] in position order → [n, …] in id order (drops the padding rows)."""
        return x_full.index_select(0, self.pos_tensor(x_full.device))

    def _pad_rows(self, m):
        if m.shape[0] < self.block:
            m = sp.vstack([m, sp.csr_matrix((self.block - m.shape[0], m.shape[1]), dtype=m.dtype)]).tocsr()
        return m

    def rows_of(self, mat):
        """The owned rows of a scipy matrix [n, c] as CSR [block, c] (empty padding rows); columns untouched."""
        return self._pad_rows(sp.csr_matrix(mat)[self.my_ids])

    def square_rows_of(self, mat, keep_duplicates=False):
        """The owned rows of an [n, n] matrix whose columns index a GATHERED tensor: columns renumbered to positions,
        shape [block, n_pad].  keep_duplicates: repeated (row, col) entries stay separate entries (RDGCN's r_mat holds
        one entry per triple) instead of being summed."""
        m = sp.csr_matrix(mat) if not sp.isspmatrix_csr(mat) else mat
        counts = np.diff(m.indptr)[self.my_ids]
        indptr = np.full(self.block + 1, counts.sum(), dtype=np.int64)
        indptr[:self.n_local + 1] = np.concatenate([[0], np.cumsum(counts)])
        take = np.concatenate([np.arange(m.indptr[i], m.indptr[i + 1]) for i in self.my_ids]) if self.n_local and counts.sum() \
            else np.zeros(0, dtype=np.int64)
        out = sp.csr_matrix((m.data[take], self.pos_of_id[m.indices[take]], indptr), shape=(self.block, self.n_pad))
        if not keep_duplicates:
            out.sum_duplicates()
        return out

    def cols_of(self, mat):
        """The owned columns of a scipy matrix [c, n] as CSR [c, block] (a reduction OVER rows of a sharded tensor)."""
        m = sp.csc_matrix(mat)[:, self.my_ids]
        if m.shape[1] < self.block:
            m = sp.hstack([m, sp.csc_matrix((m.shape[0], self.block - m.shape[1]), dtype=m.dtype)])
        return sp.csr_matrix(m)

    def local_rows(self, full):
        """The owned rows of a host array (or CPU tensor) padded with zero rows to `block`."""
        full = np.asarray(full)
        out = np.zeros((self.block,) + full.shape[1:], dtype=full.dtype)
        out[:self.n_local] = full[self.my_ids]
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
        mine = sh.pos_of_id[ill[self.p_lo:self.p_hi]]              # the loss kernel indexes the gathered (position-order) rows
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
        self.outputs = sh.to_global(self._out_full)
        return self.outputs

    def _my_negs(self, negs):
        lo, hi = self.p_lo * self.k, self.p_hi * self.k
        pos = self.shard.pos_tensor(negs[0].device)
        return [pos[n[lo:hi].long()].to(torch.int32).contiguous() for n in negs]

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
            local[:sh.n_local] = full[torch.as_tensor(sh.my_ids, device=full.device)]
            self.params["init_embedding"] = local.requires_grad_(True)

        def _layer_input(self, x):
            return AllGatherRows.apply(x, self.shard, False)

        def _layer_outputs(self, outs):
            return [self.shard.to_global(AllGatherRows.apply(o, self.shard, True)) for o in outs]

        def input_embedding(self):
            return self.shard.to_global(AllGatherRows.apply(self.params["init_embedding"], self.shard, True))

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
                return self.ops.csr(sh.square_rows_of(mat, keep_duplicates=True), device, keep_duplicates=True)
            return self.ops.csr(sh.square_rows_of(mat), device)

        def _entity_cols(self, mat, device):
            return self.ops.csr(self.shard.cols_of(mat), device)

        def _own_rows(self, x):
            sh = self.shard
            local = torch.zeros(sh.block, x.shape[1], dtype=x.dtype)
            local[:sh.n_local] = x[torch.as_tensor(sh.my_ids)]
            return local

        def _all_rows(self, x):
            return AllGatherRows.apply(x, self.shard, False)

        def _sum_over_owners(self, x):
            return AllReduceSum.apply(x) if self.shard.world > 1 else x

        def _full_output(self, x):
            return self.shard.to_global(AllGatherRows.apply(x, self.shard, True))

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
