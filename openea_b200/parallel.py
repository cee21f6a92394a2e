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
    """Per-epoch exchange of the seed-pair rows: every rank publishes the rows it owns, every rank receives the others'.

    Three transports, same result:
      'p2p'   (GPUs of one box) — oea_seed_push / oea_seed_pull (csrc/oea_p2p.cu): one kernel stores the owned rows straight
              into every peer's exchange window over NVLink and raises a release flag; the consumer kernel acquires the
              flags and copies the rows into its table.  No library collective, no host synchronisation.
      'nccl'  — oea_seed_pack → ncclAllGather (asynchronous, overlaps the next training step) → oea_seed_unpack; the
              fallback where peer mapping (CUDA IPC) is not available.
      'torch' — index_select / all_gather / index_copy_ (CPU tensors with gloo: the host-logic tests).
    `push()` publishes, `pull()` applies the last publication; a training loop calls push() after the last step of an
    epoch and pull() one step later so the transfer overlaps a step; `sync()` = push + pull."""

    def __init__(self, weight, seed_ids, rank, world_size, mode="auto", timeout_s=20.0):
        self.weight = weight                    # [rows, pitch] tensor (CUDA with NCCL, CPU with gloo)
        self.rank, self.world = rank, world_size
        ids = np.unique(np.asarray(seed_ids, dtype=np.int64))
        own = owner_of(ids, world_size)
        per_owner = [ids[own == g] for g in range(world_size)]
        self.counts = [len(p) for p in per_owner]
        self.max_cnt = max(1, max(self.counts))
        dev = weight.device
        pitch = weight.shape[1]
        slot = np.full((world_size, self.max_cnt), -1, dtype=np.int64)
        for g, p in enumerate(per_owner):
            slot[g, :len(p)] = p
        self.n_own = self.counts[rank]
        self.bytes_per_sync = world_size * self.max_cnt * pitch * weight.element_size()
        self.epoch = 0
        self._pending = None
        self._x = None
        self.timeout_ns = int(timeout_s * 1e9)
        if mode == "auto":
            mode = "p2p" if weight.is_cuda else "torch"
        if world_size == 1:
            mode = "torch"
        self.mode = mode
        if mode == "torch":
            self.mine = torch.as_tensor(per_owner[rank], dtype=torch.long, device=dev)
            flat = slot.copy()
            flat[rank, :] = -1                  # own rows are authoritative: never overwritten
            flat = flat.reshape(-1)
            self.valid = torch.as_tensor(np.flatnonzero(flat >= 0), dtype=torch.long, device=dev)
            self.dst_rows = torch.as_tensor(flat[flat >= 0], dtype=torch.long, device=dev)
            self.send = torch.zeros(self.max_cnt, pitch, dtype=weight.dtype, device=dev)
            self.recv = torch.zeros(world_size * self.max_cnt, pitch, dtype=weight.dtype, device=dev)
            return
        from . import lib as L
        self._L, self._lib = L, L.load()
        self.own_ids = torch.as_tensor(per_owner[rank], dtype=torch.int32, device=dev)
        self.slot_ids = torch.as_tensor(slot.reshape(-1), dtype=torch.int32, device=dev)
        if mode == "p2p" and not self._open_windows(pitch):
            self.mode = mode = "nccl"           # every rank takes the same decision (all-reduced inside)
        if mode == "nccl":
            self.send = torch.zeros(self.max_cnt, pitch, dtype=weight.dtype, device=dev)
            self.recv = torch.zeros(world_size * self.max_cnt, pitch, dtype=weight.dtype, device=dev)

    # ---- p2p plumbing: windows are created by the library (IPC-exportable), handles travel through torch.distributed ----
    def _open_windows(self, pitch):
        import ctypes as C
        L, lib = self._L, self._lib
        dev = self.weight.device
        nbytes = lib.oea_seed_xchg_window_bytes(self.world, self.max_cnt, pitch)
        ok = 1
        local = C.c_void_p(0)
        handle = (C.c_ubyte * 64)()
        if self.world > L.P2P_MAX_WORLD or lib.oea_p2p_window_create(nbytes, C.byref(local), handle) != 0:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if ok else None)
        opened = {}
        if ok and all(h is not None for h in handles):
            for g, h in enumerate(handles):
                if g == self.rank:
                    continue
                ptr = C.c_void_p(0)
                if lib.oea_p2p_window_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(ptr)) != 0:
                    ok = 0
                    break
                opened[g] = ptr.value
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            for ptr in opened.values():
                lib.oea_p2p_window_close(C.c_void_p(ptr))
            if local.value:
                lib.oea_p2p_window_destroy(local)
            return False
        self._local_window, self._opened = local.value, opened
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        win = (C.c_void_p * 16)()
        for g in range(self.world):
            win[g] = local.value if g == self.rank else opened[g]
        self._x = L.SeedXchg(self.rank, self.world, pitch, self.max_cnt, win, self.own_ids.data_ptr(), self.n_own,
                             self.slot_ids.data_ptr(), self._ticket.data_ptr())
        dist.barrier()                          # every window is mapped everywhere before the first push
        return True

    def _stream(self):
        import ctypes as C
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def push(self):
        """Publish this rank's owned rows (epoch += 1)."""
        if self.world == 1:
            return
        self.epoch += 1
        if self.mode == "p2p":
            import ctypes as C
            self._L.check(self._lib.oea_seed_push(C.byref(self._x), C.c_void_p(self.weight.data_ptr()), self.epoch,
                                                  self._stream()), "oea_seed_push")
        elif self.mode == "nccl":
            import ctypes as C
            self._L.check(self._lib.oea_seed_pack(C.c_void_p(self.weight.data_ptr()), self.weight.shape[1],
                                                  C.c_void_p(self.own_ids.data_ptr()), self.n_own,
                                                  C.c_void_p(self.send.data_ptr()), self._stream()), "oea_seed_pack")
            self._pending = dist.all_gather_into_tensor(self.recv, self.send, async_op=True)
        else:
            n = self.mine.numel()
            if n:
                self.send[:n] = self.weight.index_select(0, self.mine)
            dist.all_gather_into_tensor(self.recv, self.send)

    def pull(self):
        """Apply the last publication of every peer to this replica."""
        if self.world == 1 or self.epoch == 0:
            return
        if self.mode == "p2p":
            import ctypes as C
            self._L.check(self._lib.oea_seed_pull(C.byref(self._x), C.c_void_p(self.weight.data_ptr()), self.epoch,
                                                  self.timeout_ns, self._stream()), "oea_seed_pull")
        elif self.mode == "nccl":
            import ctypes as C
            if self._pending is not None:
                self._pending.wait()            # stream-level dependency, no host block
                self._pending = None
            self._L.check(self._lib.oea_seed_unpack(C.c_void_p(self.weight.data_ptr()), self.weight.shape[1],
                                                    C.c_void_p(self.recv.data_ptr()), C.c_void_p(self.slot_ids.data_ptr()),
                                                    self.world, self.max_cnt, self.rank, self._stream()), "oea_seed_unpack")
        else:
            if self.dst_rows.numel():
                self.weight.index_copy_(0, self.dst_rows, self.recv.index_select(0, self.valid))

    def sync(self):
        self.push()
        self.pull()

    def status(self):
        """0 = ok; 1 = a device-side wait for a peer's publication timed out (p2p mode)."""
        if self.mode != "p2p" or self._x is None:
            return 0
        import ctypes as C
        st = C.c_int32(0)
        self._L.check(self._lib.oea_seed_xchg_status(C.byref(self._x), C.byref(st)), "oea_seed_xchg_status")
        return int(st.value)

    def close(self):
        """Unmap the peers' windows and free the local one (collective: every rank calls it)."""
        if self.mode != "p2p" or self._x is None:
            return
        import ctypes as C
        torch.cuda.synchronize()
        dist.barrier()                          # nobody is still storing into a window that is about to go away
        for ptr in self._opened.values():
            self._lib.oea_p2p_window_close(C.c_void_p(ptr))
        dist.barrier()
        self._lib.oea_p2p_window_destroy(C.c_void_p(self._local_window))
        self._x = None


class ExactReplicaStep:
    """Exact-parity multi-GPU mode of path (i) ("exact-parity alternative"): tables REPLICATED, the BATCH
    sharded.  Rank g scores the positives p ≡ g (mod G) of the very batch one GPU would draw (the sampler's draws depend
    on p, not on the shard), then gradients, row flags and the loss are all-reduced and the identical optimiser step is
    applied on every replica.  Replicas stay bit-identical to each other; against one GPU the only difference is the
    fp32 summation order of the gradient (1e-4 relative, the same tolerance as the oracle tests)."""

    def __init__(self, trainer):
        self.trainer = trainer
        self.loss_step = torch.zeros(1, dtype=torch.float64, device=trainer.ent.device)

    def step(self, kg1, kg2, tset, batch_size, neg_per_pos, step, epoch_seed, max_try=10):
        tr = self.trainer
        rank, ws = world()
        if ws == 1:
            tr.score_sampled(kg1, kg2, tset, batch_size, neg_per_pos, step, epoch_seed, max_try=max_try)
            tr.apply()
            return
        self.loss_step.zero_()
        tr.score_sampled(kg1, kg2, tset, batch_size, neg_per_pos, step, epoch_seed, max_try=max_try,
                         loss_out=self.loss_step, shard=(rank, ws))
        for tab in (tr.ent, tr.rel):
            dist.all_reduce(tab.grad)
            dist.all_reduce(tab.touched, op=dist.ReduceOp.MAX)
        dist.all_reduce(self.loss_step)
        tr.loss_dev += self.loss_step
        tr.apply()


class ReplicaDeltaSum:
    """Periodic combination of independently trained replicas (local-update data parallelism): every replica adds up what
    ALL replicas changed since the last combination,  x ← x_ref + Σ_g (x_g − x_ref),  for the variables and for their
    optimiser slots (Adagrad's accumulator is a sum of g², so the same rule gives the accumulator one process would
    hold).  Unlike the owner-wins seed-row exchange no gradient contribution is discarded — a row trained as a tail on
    one rank and as a head on another receives both — at the price of an all-reduce of the tables per combination.
    Replicas are identical after every sync()."""

    def __init__(self, tables):
        self.items = []
        for t in tables:
            for name in ("weight", "state1", "state2"):
                x = getattr(t, name, None)
                if x is not None:
                    self.items.append((x, x.clone()))
        self.bytes_per_sync = sum(x.numel() * x.element_size() for x, _ in self.items)

    def sync(self):
        if world()[1] == 1:
            return
        for x, ref in self.items:
            x -= ref                      # this replica's change since the last combination
            dist.all_reduce(x)            # everybody's changes
            x += ref
            ref.copy_(x)


def assemble_owned_rows(weight, rank, world_size):
    """Final table: every row taken from its owner (all rows, same mechanism as the seed sync)."""
    if world_size == 1:
        return weight
    # one-off, whole table: the library all-gather is the right tool (no window of table size is mapped for it)
    x = SeedRowSync(weight, np.arange(weight.shape[0]), rank, world_size, mode="nccl" if weight.is_cuda else "torch")
    x.sync()
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
