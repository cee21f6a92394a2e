"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The reference has no multi-device code at all.  What is sharded here, and the only collectives:

 path (i)   triple training: tables are replicated (12–80 MB), every rank OWNS the rows with id % G == rank and
            trains on the triples whose head it owns; once per epoch the owners' copies of the seed-pair rows are
            all-gathered and written into every replica (`SeedRowSync`), the only data-path collective
            (BASELINE.json north_star).  Replicas are otherwise stale ⇒ statistical, not bit-wise, parity with
            one GPU.  `assemble_owned_rows` builds the final table from the owners' rows.
 path (iii) evaluation: E1 rows are block-sharded, E2 is replicated; row top-k / rank are local; the CSLS column
            means need each column's k best over ALL rows: every rank contributes its partial [n2, k] list,
            one all-gather, then a k-way merge (`merge_partial_topk`); Hits/MR/MRR are 3 all-reduced scalars.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# Set by the approaches whose ranks hold identical embeddings (the row-sharded GNN models gather every layer output on
# every rank).  Only then may the evaluation be sharded: it combines row blocks and column lists computed on different
# ranks, which is meaningless for independently trained replicas.
_replicas_in_sync = False


def mark_replicas_in_sync(value=True):
    global _replicas_in_sync
    _replicas_in_sync = bool(value)


def replicas_in_sync():
    return _replicas_in_sync and world()[1] > 1


def owner_of(ids, world_size):
    """Cyclic row ownership: ids interleave the two KGs by descending frequency (read.py:69-79), so contiguous
    blocks would put every hub on rank 0."""
    return np.asarray(ids) % world_size


def shard_triples(triples, rank, world_size):
    """Triples whose HEAD is owned by `rank` (the rank that will produce most of that row's gradient)."""
    tri = np.asarray(triples).reshape(-1, 3)
    return tri[owner_of(tri[:, 0], world_size) == rank]


def block_range(n, rank, world_size):
    """Contiguous row block of `rank` when n rows are split as evenly as possible."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SeedRowSync:
    """Per-epoch exchange of the seed-pair rows: every rank sends the rows it owns, all ranks receive all."""

    def __init__(self, weight, seed_ids, rank, world_size):
        self.weight = weight                    # [rows, pitch] tensor (CUDA with NCCL, CPU with gloo)
        self.rank, self.world = rank, world_size
        ids = np.unique(np.asarray(seed_ids, dtype=np.int64))
        own = owner_of(ids, world_size)
        per_owner = [ids[own == g] for g in range(world_size)]
        self.counts = [len(p) for p in per_owner]
        self.max_cnt = max(1, max(self.counts))
        dev = weight.device
        self.mine = torch.as_tensor(per_owner[rank], dtype=torch.long, device=dev)
        # destination row of every received slot (padding slots point at a scratch row index −1 → masked out)
        dst = np.full((world_size, self.max_cnt), -1, dtype=np.int64)
        for g, p in enumerate(per_owner):
            dst[g, :len(p)] = p
        flat = dst.reshape(-1)
        self.valid = torch.as_tensor(np.flatnonzero(flat >= 0), dtype=torch.long, device=dev)
        self.dst_rows = torch.as_tensor(flat[flat >= 0], dtype=torch.long, device=dev)
        pitch = weight.shape[1]
        self.send = torch.zeros(self.max_cnt, pitch, dtype=weight.dtype, device=dev)
        self.recv = torch.zeros(world_size * self.max_cnt, pitch, dtype=weight.dtype, device=dev)
        self.bytes_per_sync = self.recv.numel() * self.recv.element_size()

    def sync(self):
        if self.world == 1:
            return
        n = self.mine.numel()
        if n:
            self.send[:n] = self.weight.index_select(0, self.mine)
        dist.all_gather_into_tensor(self.recv, self.send)
        self.weight.index_copy_(0, self.dst_rows, self.recv.index_select(0, self.valid))


def assemble_owned_rows(weight, rank, world_size):
    """Final table: every row taken from its owner (all rows, same mechanism as the seed sync)."""
    if world_size == 1:
        return weight
    SeedRowSync(weight, np.arange(weight.shape[0]), rank, world_size).sync()
    return weight


def merge_partial_topk(partial, k):
    """partial: [G, n, k] values (each rank's k best of a column over its row block) → mean of the k best overall.
    torch.topk on a [n, G·k] view: a k-way merge of tiny lists."""
    g, n, kk = partial.shape
    allv = partial.permute(1, 0, 2).reshape(n, g * kk)
    return allv.topk(k, dim=1).values.mean(dim=1)


def allgather_partial(values):
    """[n, k] per rank → [G, n, k] on every rank."""
    rank, ws = world()
    if ws == 1:
        return values.unsqueeze(0)
    n = values.shape[0]
    out = torch.empty((ws * n,) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
    dist.all_gather_into_tensor(out, values.contiguous())      # concatenation along dim 0 (gloo and NCCL)
    return out.view((ws, n) + tuple(values.shape[1:]))


def allreduce_sum(t):
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allgather_blocks(x_local, n_total):
    """Every rank holds the rows block_range(n_total, rank, G) of a tensor: returns the whole [n_total, …] tensor on every
    rank (blocks differ by at most one row: padded to the largest, gathered, re-assembled)."""
    rank, ws = world()
    if ws == 1:
        return x_local
    biggest = -(-n_total // ws)
    send = torch.zeros((biggest,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    send[:x_local.shape[0]] = x_local
    recv = torch.empty((ws * biggest,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(recv, send)
    parts = []
    for r in range(ws):
        lo, hi = block_range(n_total, r, ws)
        parts.append(recv[r * biggest:r * biggest + (hi - lo)])
    return torch.cat(parts, 0)
