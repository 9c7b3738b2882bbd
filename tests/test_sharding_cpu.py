"""Bipartite-sharding host logic on CPU, incl. a real world-size-2 gloo process group: user partition, per-rank
blocks Ru / Rt, and the layer exchange (partial item products summed over ranks) reproduce the unsharded product."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bipartite(n_users, n_items, nnz, seed):
    rng = np.random.default_rng(seed)
    u = np.minimum((rng.pareto(1.2, nnz) * 3).astype(np.int64), n_users - 1)  # power-law: hubs at low ids
    i = rng.integers(0, n_items, nnz)
    R = sp.csr_matrix((np.ones(nnz, np.float32), (u, i)), shape=(n_users, n_items))
    R.data[:] = rng.random(R.nnz).astype(np.float32) + 0.1
    n = n_users + n_items
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    A.sort_indices()
    assert A.shape == (n, n)
    return R.tocsr(), A


def test_partition_rows_and_users():
    from selfrec_b200.sharded import item_bounds, partition_rows, partition_users
    rng = np.random.default_rng(0)
    deg = np.concatenate([rng.zipf(1.5, 500) % 300, np.zeros(20, int), [4000]])  # hubs, empty rows, one giant row
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    for world in (1, 2, 3, 8):
        b = partition_rows(rowptr, world)
        assert b[0] == 0 and b[-1] == len(deg) and (np.diff(b) >= 0).all() and len(b) == world + 1
        per = np.diff(rowptr[b])
        assert per.sum() == rowptr[-1]
        assert per.max() <= rowptr[-1] / world + deg.max()  # at most one row over the ideal share
        bu = partition_users(rowptr, world)
        assert bu[0] == 0 and bu[-1] == len(deg) and (np.diff(bu) >= 32).all()
        assert all(x % 32 == 0 for x in bu[1:-1])
        ib = item_bounds(1001, world)
        assert ib[0] == 0 and ib[-1] == 1001 and (np.diff(ib) > 0).all() and np.diff(ib).max() <= -(-1001 // world)
    # a giant first row must not starve the other ranks
    bu = partition_users(np.concatenate([[0], np.cumsum([10**6] + [1] * 511)]), 8)
    assert (np.diff(bu) >= 32).all()
    with pytest.raises(Exception):
        partition_users(np.arange(41), 2)  # 40 users cannot give two blocks of >= 32


def test_extract_blocks_tile_the_adjacency():
    import torch
    from selfrec_b200.sharded import extract_blocks, partition_users
    U, I = 400, 150
    R, A = _bipartite(U, I, 6000, 1)
    rp, ci, vv = (torch.from_numpy(np.asarray(x)) for x in (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data))
    for world in (1, 2, 4):
        b = partition_users(A.indptr[:U + 1], world)
        rus, rts = [], []
        for g in range(world):
            (p1, c1, v1), (p2, c2, v2) = extract_blocks(rp, ci, vv, U, I, int(b[g]), int(b[g + 1]))
            ug = int(b[g + 1] - b[g])
            rus.append(sp.csr_matrix((v1.numpy(), c1.numpy(), p1.numpy()), shape=(ug, I)))
            rts.append(sp.csr_matrix((v2.numpy(), c2.numpy(), p2.numpy()), shape=(I, ug)))
            assert rts[-1].has_sorted_indices and rus[-1].has_sorted_indices
        assert abs(sp.vstack(rus) - R).max() == 0
        assert abs(sp.hstack(rts) - R.T).max() == 0


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from selfrec_b200.sharded import extract_blocks, item_bounds, partition_users
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, I, d = 400, 150, 32
    R, A = _bipartite(U, I, 6000, 2)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((U + I, d)).astype(np.float32)
    rp, ci, vv = (torch.from_numpy(np.asarray(x)) for x in (A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data))
    b = partition_users(A.indptr[:U + 1], world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    (p1, c1, v1), (p2, c2, v2) = extract_blocks(rp, ci, vv, U, I, lo, hi)
    Ru = sp.csr_matrix((v1.numpy(), c1.numpy(), p1.numpy()), shape=(hi - lo, I))
    Rt = sp.csr_matrix((v2.numpy(), c2.numpy(), p2.numpy()), shape=(I, hi - lo))
    ib = item_bounds(I, world)
    xu, xi = X[lo:hi], X[U:]
    for _ in range(2):  # two propagation layers
        part = torch.from_numpy((Rt @ xu).astype(np.float32))  # this rank's partial item product
        yu = (Ru @ xi).astype(np.float32)                      # local user half
        # reduce-scatter to the slice owners, then all-gather of the finished slices (the CPU stand-in for the
        # P2P partial pushes + owner-side reduction + multicast stores of sharded.cu)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        mine = sum(p[ib[rank]:ib[rank + 1]] for p in parts)   # fixed rank order
        sizes = [int(ib[g + 1] - ib[g]) for g in range(world)]
        pad = torch.zeros((max(sizes), d))
        pad[: sizes[rank]] = mine
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        xi = torch.cat([o[:n] for o, n in zip(outs, sizes)]).numpy()
        xu = yu
    ref = A @ (A @ X)
    ok = np.allclose(xu, ref[lo:hi], rtol=1e-5, atol=1e-5) and np.allclose(xi, ref[U:], rtol=1e-5, atol=1e-5)
    out = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(out, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(float(out.item()))
    dist.destroy_process_group()


def test_bipartite_sharded_propagation_world2_gloo():
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
