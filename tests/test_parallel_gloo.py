"""CPU: the N>1 host logic under torch.distributed/gloo with world_size 2 — triple sharding by head owner,
the per-epoch seed-row all-gather (SeedRowSync), owner assembly, and the k-way merge of partial column top-k
lists used by the sharded CSLS evaluation (values checked against the NumPy oracle)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openea_b200 import parallel as par
    from oracle import finding as orf
    try:
        # ---- triple sharding: a partition by head owner
        rng = np.random.default_rng(0)
        tri = np.stack([rng.integers(0, 101, 500), rng.integers(0, 7, 500), rng.integers(0, 101, 500)], 1)
        mine = par.shard_triples(tri, rank, world)
        assert (mine[:, 0] % world == rank).all()
        cnt = torch.tensor([len(mine)])
        dist.all_reduce(cnt)
        assert int(cnt) == 500
        # ---- seed-row sync: after the sync every replica holds the OWNER's copy of every seed row
        rows, pitch = 101, 8
        w = torch.full((rows, pitch), float(rank + 1))          # replica r holds value r+1 everywhere
        seeds = np.array([3, 4, 10, 11, 11, 50, 99, 100])
        before = w.clone()
        sync = par.SeedRowSync(w, seeds, rank, world)
        sync.sync()
        for i in range(rows):
            if i in set(seeds.tolist()):
                assert (w[i] == float(i % world + 1)).all(), (rank, i, w[i])
            else:
                assert (w[i] == before[i]).all()
        assert sync.bytes_per_sync == world * sync.max_cnt * pitch * 4
        # ---- owner assembly: every row from its owner
        w2 = torch.full((rows, pitch), float(rank + 1))
        par.assemble_owned_rows(w2, rank, world)
        want = torch.tensor([(i % world) + 1.0 for i in range(rows)])
        assert (w2[:, 0] == want).all()
        # ---- sharded CSLS column means: partial top-k per row block, all-gather, merge == oracle
        rng = np.random.default_rng(5)
        s = rng.standard_normal((37, 23)).astype(np.float32)     # [n1, n2]
        k = 5
        lo, hi = par.block_range(37, rank, world)
        part_v, _ = orf.topk_rows(np.ascontiguousarray(s[lo:hi].T), k)       # columns over my rows → [n2, k]
        merged = par.merge_partial_topk(par.allgather_partial(torch.from_numpy(part_v)), k).numpy()
        np.testing.assert_allclose(merged, orf.nearest_k_mean(np.ascontiguousarray(s.T), k), rtol=1e-6)
        # ---- ragged row blocks re-assembled on every rank (the arg-max columns of the sharded evaluation)
        full = torch.arange(37 * 3, dtype=torch.int32).reshape(37, 3)
        assert torch.equal(par.allgather_blocks(full[lo:hi].clone(), 37), full)
        assert torch.equal(par.allgather_blocks(full[lo:hi, 0].clone(), 37), full[:, 0])
        ranges = [par.block_range(37, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 37 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
