"""Row-sharding host logic on CPU with a real world-size-2 gloo process group: the partition,
the CSR slices and the gather of per-rank row blocks reproduce the unsharded product."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_rows_balances_nnz():
    from selfrec_b200.sharded import LocalShard, partition_rows
    rng = np.random.default_rng(0)
    deg = np.concatenate([rng.zipf(1.5, 500) % 300, np.zeros(20, int), [4000]])  # hubs, empty rows, one giant row
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    for world in (1, 2, 3, 8):
        b = partition_rows(rowptr, world)
        assert b[0] == 0 and b[-1] == len(deg) and (np.diff(b) >= 0).all() and len(b) == world + 1
        per = np.diff(rowptr[b])
        assert per.sum() == rowptr[-1]
        assert per.max() <= rowptr[-1] / world + deg.max()  # at most one row over the ideal share
    # slices tile the matrix exactly
    A = sp.random(300, 300, density=0.05, random_state=1, format="csr", dtype=np.float32)
    parts = [LocalShard(A, r, 4) for r in range(4)]
    assert [p.row_begin for p in parts[1:]] == [p.row_end for p in parts[:-1]]
    assert abs(sp.vstack([p.local_csr() for p in parts]) - A).max() == 0
    assert all(np.all(np.diff(np.diff(p.rowptr)[p.row_order]) <= 0) for p in parts)  # degree-descending order


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from selfrec_b200.sharded import LocalShard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    n, d = 400, 64
    A = sp.random(n, n, density=0.03, random_state=2, format="csr", dtype=np.float32)
    A = (A + A.T).tocsr()
    X = rng.standard_normal((n, d)).astype(np.float32)
    sh = LocalShard(A, rank, world)
    # two propagation layers: every rank computes its row block, blocks are all-gathered (the CPU
    # stand-in for the fused NVLink push), the next layer consumes the gathered table
    cur = X
    for _ in range(2):
        mine = torch.from_numpy((sh.local_csr() @ cur).astype(np.float32))
        sizes = [int(sh.bounds[g + 1] - sh.bounds[g]) for g in range(world)]
        blocks = [torch.empty((s, d)) for s in sizes]
        dist.all_gather(blocks, mine) if len(set(sizes)) == 1 else [dist.broadcast(blocks[g].copy_(mine) if g == rank else blocks[g], src=g) for g in range(world)]
        cur = torch.cat(blocks).numpy()
    ref = A @ (A @ X)
    ok = np.allclose(cur, ref, rtol=1e-5, atol=1e-6)
    out = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(out, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(float(out.item()))
    dist.destroy_process_group()


def test_sharded_propagation_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1.0
