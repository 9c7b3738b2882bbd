"""Row-sharded multi-GPU path: one process per GPU, tables split by contiguous,
nnz-balanced row blocks, the per-layer all-gather fused into the SpMM epilogue.

Host logic (partitioning, CSR slicing) is plain numpy and is exercised on CPU with a world-size-2
gloo group (tests/test_sharding_cpu.py).  The device path needs NVLink-connected GPUs: peer
pointers come from torch.distributed._symmetric_memory, every finished row of a propagated layer
is stored by the SpMM kernel into each rank's copy (srb_spmm_csr_allgather), and a device-side
symmetric-memory barrier separates producers from consumers.  torch.distributed (NCCL) is only the
plumbing: rendezvous and the barrier; the data never goes through a NCCL collective.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib


def partition_rows(rowptr, world):
    """Contiguous row blocks with (nearly) equal non-zero counts: bounds[g] .. bounds[g+1]."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = len(rowptr) - 1
    nnz = int(rowptr[-1])
    targets = (np.arange(1, world) * nnz) // world
    cuts = np.searchsorted(rowptr, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


class LocalShard:
    """The CSR slice A[R_r, :] of one rank (row pointers rebased, column ids global)."""

    def __init__(self, csr, rank, world, long_row_nnz=64, vlong_row_nnz=256):
        csr = sp.csr_matrix(csr, dtype=np.float32)
        csr.sort_indices()
        self.n = csr.shape[0]
        self.bounds = partition_rows(csr.indptr, world)
        self.rank, self.world = rank, world
        self.row_begin, self.row_end = int(self.bounds[rank]), int(self.bounds[rank + 1])
        lo, hi = csr.indptr[self.row_begin], csr.indptr[self.row_end]
        self.rowptr = (csr.indptr[self.row_begin:self.row_end + 1] - lo).astype(np.int32)
        self.colidx = csr.indices[lo:hi].astype(np.int32)
        self.vals = csr.data[lo:hi].astype(np.float32)
        deg = np.diff(self.rowptr)
        self.row_order = np.argsort(-deg, kind="stable").astype(np.int32)
        self.n_vlong = int((deg >= vlong_row_nnz).sum())
        self.n_long = int((deg >= long_row_nnz).sum()) - self.n_vlong

    @property
    def n_rows(self):
        return self.row_end - self.row_begin

    def local_csr(self):
        return sp.csr_matrix((self.vals, self.colidx, self.rowptr), shape=(self.n_rows, self.n))


class ShardedPropagator:
    """Device side of one rank: local CSR + symmetric [N, d] buffers every rank can store into."""

    def __init__(self, csr, d, n_buffers, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        _lib.require_device()
        self.torch, self.dist = torch, dist
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise _lib.SrbError("row-sharded path supports up to 8 ranks (one NVSwitch domain)")
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.d = d
        sh = LocalShard(csr, self.rank, self.world)
        self.shard = sh
        self.N = sh.n
        to = lambda a: torch.from_numpy(a).to(self.dev)
        self.rowptr, self.colidx, self.vals, self.row_order = to(sh.rowptr), to(sh.colidx), to(sh.vals), to(sh.row_order)
        self.bufs, self.handles = [], []
        for _ in range(n_buffers):
            t = symm.empty((self.N, d), dtype=torch.float32, device=self.dev)
            h = symm.rendezvous(t, self.group)
            t.zero_()
            self.bufs.append(t)
            self.handles.append(h)
        # NVSwitch multicast (NVLS): one store to the multicast mapping of a symmetric buffer lands in every
        # rank's copy, so the SpMM epilogue issues 1 store per row instead of `world` and each GPU's NVLink
        # egress drops from (world-1)/world of a layer to 1/world of it.  Measured at 2 GPUs it is slower than
        # unicast (the local copy also travels through the switch: 2984 vs 3281 steps/s), so it is the default
        # from 4 ranks up; SRB_MULTICAST=0 / 1 forces unicast / multicast.
        import os
        self.mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.handles]
        want = os.environ.get("SRB_MULTICAST", "auto")
        self.use_mc = self.world > 1 and all(m != 0 for m in self.mc) and (want == "1" or (want == "auto" and self.world >= 4))
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def peer_ptrs(self, buf_index):
        h = self.handles[buf_index]
        arr = (C.c_void_p * 8)()
        for g in range(self.world):
            arr[g] = int(h.buffer_ptrs[g])
        return arr

    def barrier(self):
        """Device-side barrier across ranks on the current stream (pushed rows become visible)."""
        self.handles[0].barrier(channel=0)

    def spmm(self, x, push_y=None, push_sum=None, push_p=None, row_list=None, **epi):
        """Own rows of A @ x with the fused pushes; x and all epilogue tensors are full [N, d].
        row_list = (rows, counters, capacity): only the listed local rows (srb_build_batch_rows)."""
        from . import ops
        torch = self.torch
        lib = _lib.load()
        sd = _lib.SpmmShardedDesc()
        loc = sd.local
        loc.rowptr, loc.colidx, loc.vals = ops._p(self.rowptr), ops._p(self.colidx), ops._p(self.vals)
        loc.row_order, loc.n_long_rows, loc.n_vlong_rows = ops._p(self.row_order), self.shard.n_long, self.shard.n_vlong
        loc.n_rows, loc.n_cols, loc.d = self.shard.n_rows, self.N, self.d
        loc.X = ops._p(x)
        loc.extra_scale, loc.sum_scale = 1.0, 1.0
        keep = [x]
        for k, v in epi.items():
            if isinstance(v, torch.Tensor):
                keep.append(v)
                setattr(loc, k, ops._p(v))
            else:
                setattr(loc, k, v)
        if row_list is not None:
            rows, counters, cap = row_list
            keep += [rows, counters]
            loc.row_order, loc.n_rows, loc.n_long_rows, loc.n_vlong_rows = ops._p(rows), cap, 0, 0
            loc.n_vlong_dev = ops._p(counters)
        sd.row_begin, sd.world = self.shard.row_begin, (1 if self.use_mc else self.world)
        for idx, field in ((push_y, "peer_Y"), (push_sum, "peer_sum"), (push_p, "peer_p")):
            if idx is not None:
                if self.use_mc:
                    getattr(sd, field)[0] = self.mc[idx]  # the one "peer" is the multicast address
                    continue
                arr = self.peer_ptrs(idx)
                for g in range(self.world):
                    getattr(sd, field)[g] = arr[g]
        _lib.check(lib.srb_spmm_csr_allgather(C.byref(sd), ops._stream()), "srb_spmm_csr_allgather")


class ShardedXSimGCL:
    """XSimGCL / LightGCN training step on row-sharded tables.

    SpMMs and Adam are sharded by rows (each rank computes and pushes its block); the batch losses
    (BPR, L2, InfoNCE over <= 3B + 2B gathered rows) are replicated on every rank from the gathered
    layers -- they touch ~2 MB and would cost more to distribute than to recompute.  All ranks hold
    bit-identical parameters after every step because every rank consumes the same pushed rows.
    Buffers (symmetric): 0 params, 1/2 layer ping-pong, 3 cl view, 4 final (running layer sum), 5/6 backward
    ping-pong, 7 batch rows of the final mean (training steps evaluate the last layer on the batch rows only).
    """

    P, W0, W1, CL, FIN, A0, A1, FINB = range(8)

    def __init__(self, model, data, emb_size, n_layers, batch_size, lr, reg, *, eps=0.0, tau=0.2, cl_rate=0.0, layer_cl=0,
                 l2_div=1.0, init_user=None, init_item=None, group=None):
        import torch
        from . import ops
        if model not in ("XSimGCL", "LightGCN"):
            raise _lib.SrbError("sharded engine covers XSimGCL and LightGCN")
        self.torch, self.ops = torch, ops
        self.model = model
        self.prop = ShardedPropagator(data.norm_adj, emb_size, 8, group)
        p = self.prop
        self.U, self.I, self.d, self.L, self.B = int(data.user_num), int(data.item_num), int(emb_size), int(n_layers), int(batch_size)
        self.N = self.U + self.I
        self.lr, self.reg, self.eps, self.tau, self.cl_rate, self.layer_cl, self.l2_div = lr, reg, eps, tau, cl_rate, layer_cl, l2_div
        dev = p.dev
        if init_user is None:
            g = torch.Generator().manual_seed(0)  # every rank must start from the same table
            bound_u = (6.0 / (self.U + self.d)) ** 0.5
            bound_i = (6.0 / (self.I + self.d)) ** 0.5
            init_user = (torch.rand(self.U, self.d, generator=g) * 2 - 1) * bound_u
            init_item = (torch.rand(self.I, self.d, generator=g) * 2 - 1) * bound_i
        self.params = p.bufs[self.P]
        self.params[: self.U].copy_(torch.as_tensor(init_user))
        self.params[self.U:].copy_(torch.as_tensor(init_item))
        self.m = torch.zeros((self.N, self.d), device=dev)
        self.v = torch.zeros((self.N, self.d), device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(16, device=dev)
        self.losses = torch.zeros(4, device=dev)
        self.noise = None
        self.gd = torch.zeros((self.N, self.d), device=dev)
        self._alloc_step_buffers()
        torch.cuda.synchronize()
        p.dist.barrier(p.group)

    def set_noise_tensor(self, noise):
        self.noise = self.ops._f32c(noise, "noise")  # [1, L, N, d]

    def _forward(self, perturbed, philox_seed=0x5EED, batch_rows_only=False):
        p, ops = self.prop, self.ops
        L = self.L
        ego = self.model == "LightGCN"
        inv = 1.0 / (L + 1 if ego else L)
        x = p.bufs[self.P]
        for k in range(L):
            last = k == L - 1
            is_cl = perturbed and self.layer_cl == k + 1
            ybuf = self.CL if is_cl else (None if last else (self.W1 if x is p.bufs[self.W0] else self.W0))
            epi = dict(sum_out=p.bufs[self.FIN], sum_scale=inv if last else 1.0)
            if k == 0:
                if ego:
                    epi["sum_in"] = p.bufs[self.P]
            else:
                epi["sum_in"] = p.bufs[self.FIN]
            if perturbed and self.model == "XSimGCL":
                epi["eps"] = self.eps
                if self.noise is not None:
                    epi.update(noise_mode=1, noise=self.noise[0, k])
                else:
                    epi.update(noise_mode=2, philox_seed=philox_seed, philox_offset=0x10 + k, philox_step_dev=self.step_dev)
            # the running sum only needs to travel once it is final
            if last and batch_rows_only:
                # out of place (the list may hold a row twice): final rows go to FINB on every rank
                epi["sum_out"] = p.bufs[self.FINB]
                p.spmm(x, push_y=ybuf, push_sum=self.FINB, row_list=(self._brows, self._bcnt, 3 * self.B), **epi)
            else:
                p.spmm(x, push_y=ybuf, push_sum=self.FIN if last else None, **epi)
            p.barrier()
            if ybuf is not None:
                x = p.bufs[ybuf]
        return p.bufs[self.FIN], p.bufs[self.CL]

    def forward_clean(self):
        fin, _ = self._forward(False)
        out = fin.clone()
        return out[: self.U], out[self.U:]

    def _alloc_step_buffers(self):
        """Everything a step touches is allocated once: a step then makes no allocation and no host read,
        so it can be captured in a CUDA graph (counts are read by the kernels from the batch header)."""
        torch, ops, p = self.torch, self.ops, self.prop
        lib = _lib.load()
        B, d, U, dev = self.B, self.d, self.U, p.dev
        self.words = _lib.BATCH_HEADER + 5 * B
        self.batch_dev = torch.zeros(self.words, dtype=torch.int32, device=dev)
        w, H = self.batch_dev, _lib.BATCH_HEADER
        self._cnt = [w[0:1], w[1:2], w[2:3]]                       # b, n_uniq_u, n_uniq_i (device)
        self._idx = [w[H + q * B: H + (q + 1) * B] for q in range(5)]  # u, i, j, uniq_u, uniq_i (capacity B)
        self.g_emb = torch.zeros((3, B, d), device=dev)
        self.g_l2 = torch.zeros((3, B, d), device=dev) if self.model == "LightGCN" else None
        self._scratch = torch.zeros(8, device=dev)
        self._bl = torch.zeros(2, device=dev)
        self._nl = torch.zeros(2, device=dev)
        self._gn = [torch.zeros((B, d), device=dev) for _ in range(4)]  # g1 / g2 of the user and item problems
        # batch rows of this rank's block by degree class + bitmap of all batch rows (srb_build_batch_rows)
        self._brows = torch.zeros(12 * B, dtype=torch.int32, device=dev)
        self._bcnt = torch.zeros(8, dtype=torch.int32, device=dev)
        self._rmask = torch.zeros((self.N + 31) // 32, dtype=torch.int32, device=dev)
        # the final mean is only read at the batch rows: unless the last layer is the CL view, it is evaluated there only
        self._subset = not (self.model == "XSimGCL" and self.layer_cl == self.L) and self.L >= 1
        # measured and parity-checked at 2 ranks (unicast pushes, +1 %); with the multicast pushes of >= 4 ranks the
        # layer exchange is no longer what bounds the step, and the combination has not been run: keep the plain route
        self._sparse_tricks = not p.use_mc
        self._subset = self._subset and self._sparse_tricks
        fin, cl = p.bufs[self.FINB if self._subset else self.FIN], p.bufs[self.CL]
        u_idx, i_idx, j_idx, uq_u, uq_i = self._idx
        bd = _lib.BprDesc()
        bd.emb, bd.n_users, bd.d = ops._p(fin), U, d
        bd.l2_emb = ops._p(p.bufs[self.P]) if self.model == "LightGCN" else ops._p(fin)
        bd.u_idx, bd.i_idx, bd.j_idx, bd.b_dev, bd.b = ops._p(u_idx), ops._p(i_idx), ops._p(j_idx), ops._p(self._cnt[0]), B
        bd.emb_scale, bd.reg, bd.grad_scale = 1.0, self.reg, 1.0
        bd.l2_terms = 2 if self.model == "XSimGCL" else 3
        bd.l2_div = self.l2_div
        bd.losses, bd.g_emb, bd.g_l2, bd.scratch = ops._p(self._bl), ops._p(self.g_emb), ops._p(self.g_l2), ops._p(self._scratch)
        self._bpr_desc = bd
        self._nce_desc = None
        if self.model == "XSimGCL":
            ws_bytes = lib.srb_infonce_workspace_bytes(B, d, 2)
            self._nce_ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            nd = _lib.InfoNceDesc()
            nd.n_problems, nd.d, nd.b_cos, nd.temperature = 2, d, 1, float(self.tau)
            for q, (idx, cnt, off) in enumerate(((uq_u, self._cnt[1], 0), (uq_i, self._cnt[2], U))):
                pr = nd.prob[q]
                pr.table1, pr.table2, pr.row_off1, pr.row_off2 = ops._p(fin), ops._p(cl), off, off
                pr.scale1, pr.scale2 = 1.0, 1.0
                pr.idx, pr.n_dev, pr.n, pr.weight = ops._p(idx), ops._p(cnt), B, float(self.cl_rate)
                pr.g1, pr.g2 = ops._p(self._gn[2 * q]), ops._p(self._gn[2 * q + 1])
                pr.loss = C.c_void_p(self._nl.data_ptr() + 4 * q)
            nd.workspace, nd.workspace_bytes = ops._p(self._nce_ws), ws_bytes
            self._nce_desc = nd
        cm = 1.0 / (self.L + 1 if self.model == "LightGCN" else self.L)
        b_dev, nu_dev, ni_dev = self._cnt
        self._final_segs = [(self.g_emb[0], u_idx, b_dev, B, 0, cm), (self.g_emb[1], i_idx, b_dev, B, U, cm),
                            (self.g_emb[2], j_idx, b_dev, B, U, cm)]
        self._cl_segs, self._ego_segs = [], []
        if self.model == "XSimGCL":
            self._final_segs += [(self._gn[0], uq_u, nu_dev, B, 0, cm), (self._gn[2], uq_i, ni_dev, B, U, cm)]
            tgt = self._cl_segs if 1 <= self.layer_cl <= self.L else self._ego_segs
            tgt += [(self._gn[1], uq_u, nu_dev, B, 0, 1.0), (self._gn[3], uq_i, ni_dev, B, U, 1.0)]
        else:
            self._ego_segs += [(self.g_l2[0], u_idx, b_dev, B, 0, 1.0), (self.g_l2[1], i_idx, b_dev, B, U, 1.0),
                               (self.g_l2[2], j_idx, b_dev, B, U, 1.0)]

    def step(self, words=None, words_dev=None):
        """One training step; the batch buffer (srb_sampler_next_batch layout) must be the same on all ranks.
        words: host buffer (copied to the device), or words_dev: the buffer already resident on the device."""
        torch = self.torch
        if words_dev is not None:
            self.batch_dev.copy_(words_dev, non_blocking=True)
        elif words is not None:
            self.batch_dev.copy_(torch.as_tensor(np.asarray(words, dtype=np.int32)), non_blocking=True)
        self.step_resident()

    def step_resident(self):
        """Step on whatever self.batch_dev holds: no allocation, no host read (CUDA-graph capturable)."""
        torch, ops, p = self.torch, self.ops, self.prop
        lib = _lib.load()
        L = self.L
        ops.adam_prepare(self.step_dev, self.scalars, self.lr)
        sh = p.shard
        _lib.check(lib.srb_build_batch_rows(ops._p(self.batch_dev), self.B, self.U, ops._p(p.rowptr), sh.row_begin, sh.n_rows, self.N,
                                            ops._p(self._brows), ops._p(self._bcnt), ops._p(self._rmask), None, None, 0, ops._stream()),
                   "srb_build_batch_rows")
        self._forward(True, batch_rows_only=self._subset)
        # ---- replicated batch losses on the gathered layers ----
        _lib.check(lib.srb_bpr_l2_fwd_bwd(C.byref(self._bpr_desc), ops._stream()), "srb_bpr_l2_fwd_bwd")
        if self._nce_desc is not None:
            _lib.check(lib.srb_infonce_fwd_bwd(C.byref(self._nce_desc), ops._stream()), "srb_infonce_fwd_bwd")
        # ---- Horner backward, rows sharded, every level pushed to all ranks ----
        final_segs, cl_segs, ego_segs = self._final_segs, self._cl_segs, self._ego_segs
        acc = p.bufs[self.A0]
        acc.zero_()
        ops.scatter_add_segments(acc, final_segs + (cl_segs if self.layer_cl == L else []))
        x_idx = self.A0
        for k in range(L - 1, 0, -1):
            y_idx = self.A1 if x_idx == self.A0 else self.A0
            # the seed of the chain is non-zero at the batch rows only: the first product skips every other column
            p.spmm(p.bufs[x_idx], push_y=y_idx, **(dict(col_mask=self._rmask) if (k == L - 1 and self._sparse_tricks) else {}))
            p.barrier()
            # replicated: every rank adds the same sparse rows to its copy
            ops.scatter_add_segments(p.bufs[y_idx], final_segs + (cl_segs if self.layer_cl == k else []))
            x_idx = y_idx
        extra = None
        if self.model == "LightGCN" or ego_segs:
            self.gd.zero_()
            ops.scatter_add_segments(self.gd, (final_segs if self.model == "LightGCN" else []) + ego_segs)
            extra = self.gd
        epi = dict(adam_p=p.bufs[self.P], adam_m=self.m, adam_v=self.v, adam_scalars=self.scalars, beta1=0.9, beta2=0.999, adam_eps=1e-8)
        if extra is not None:
            epi["extra"] = extra
        if L == 1 and self._sparse_tricks:
            epi["col_mask"] = self._rmask
        p.spmm(p.bufs[x_idx], push_p=self.P, **epi)
        p.barrier()
        ls = self.losses
        ls[0:2].copy_(self._bl)
        if self._nce_desc is not None:
            torch.add(self._nl[0:1], self._nl[1:2], out=ls[2:3])
            ls[2:3].mul_(self.cl_rate)
        else:
            ls[2:3].zero_()
        torch.add(ls[0:1], ls[1:2], out=ls[3:4])
        ls[3:4].add_(ls[2:3])
