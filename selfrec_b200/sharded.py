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

    def spmm(self, x, push_y=None, push_sum=None, push_p=None, **epi):
        """Own rows of A @ x with the fused pushes; x and all epilogue tensors are full [N, d]."""
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
        sd.row_begin, sd.world = self.shard.row_begin, self.world
        for idx, field in ((push_y, "peer_Y"), (push_sum, "peer_sum"), (push_p, "peer_p")):
            if idx is not None:
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
    Buffers (symmetric): 0 params, 1/2 layer ping-pong, 3 cl view, 4 final, 5/6 backward ping-pong.
    """

    P, W0, W1, CL, FIN, A0, A1 = range(7)

    def __init__(self, model, data, emb_size, n_layers, batch_size, lr, reg, *, eps=0.0, tau=0.2, cl_rate=0.0, layer_cl=0,
                 l2_div=1.0, init_user=None, init_item=None, group=None):
        import torch
        from . import ops
        if model not in ("XSimGCL", "LightGCN"):
            raise _lib.SrbError("sharded engine covers XSimGCL and LightGCN")
        self.torch, self.ops = torch, ops
        self.model = model
        self.prop = ShardedPropagator(data.norm_adj, emb_size, 7, group)
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
        torch.cuda.synchronize()
        p.dist.barrier(p.group)

    def set_noise_tensor(self, noise):
        self.noise = self.ops._f32c(noise, "noise")  # [1, L, N, d]

    def _forward(self, perturbed, philox_seed=0x5EED):
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
            p.spmm(x, push_y=ybuf, push_sum=self.FIN if last else None, **epi)
            p.barrier()
            if ybuf is not None:
                x = p.bufs[ybuf]
        return p.bufs[self.FIN], p.bufs[self.CL]

    def forward_clean(self):
        fin, _ = self._forward(False)
        out = fin.clone()
        return out[: self.U], out[self.U:]

    def step(self, words, words_dev=None):
        """One training step; `words` = batch buffer (srb_sampler_next_batch layout), same on all ranks.
        words_dev: the same buffer already resident on the device (then only the 3 header ints of the
        host copy are read)."""
        torch, ops, p = self.torch, self.ops, self.prop
        lib = _lib.load()
        B, d, U, L = self.B, self.d, self.U, self.L
        w = words_dev if words_dev is not None else torch.as_tensor(np.asarray(words, dtype=np.int32)).to(p.dev)
        b, nu, ni = (int(x) for x in np.asarray(words[:3]))
        u_idx, i_idx, j_idx = w[4:4 + b], w[4 + B:4 + B + b], w[4 + 2 * B:4 + 2 * B + b]
        uq_u, uq_i = w[4 + 3 * B:4 + 3 * B + nu], w[4 + 4 * B:4 + 4 * B + ni]
        ops.adam_prepare(self.step_dev, self.scalars, self.lr)
        fin, cl = self._forward(True)
        # ---- replicated batch losses on the gathered layers ----
        g_emb = torch.empty((3, b, d), device=p.dev)
        g_l2 = torch.empty((3, b, d), device=p.dev) if self.model == "LightGCN" else None
        scratch = torch.empty(8, device=p.dev)
        bl = torch.empty(2, device=p.dev)
        bd = _lib.BprDesc()
        bd.emb, bd.n_users, bd.d = ops._p(fin), U, d
        bd.l2_emb = ops._p(p.bufs[self.P]) if self.model == "LightGCN" else ops._p(fin)
        bd.u_idx, bd.i_idx, bd.j_idx, bd.b = ops._p(u_idx), ops._p(i_idx), ops._p(j_idx), b
        bd.emb_scale, bd.reg, bd.grad_scale = 1.0, self.reg, 1.0
        bd.l2_terms = 2 if self.model == "XSimGCL" else 3
        bd.l2_div = self.l2_div
        bd.losses, bd.g_emb, bd.g_l2, bd.scratch = ops._p(bl), ops._p(g_emb), ops._p(g_l2), ops._p(scratch)
        _lib.check(lib.srb_bpr_l2_fwd_bwd(C.byref(bd), ops._stream()), "srb_bpr_l2_fwd_bwd")
        cm = 1.0 / (L + 1 if self.model == "LightGCN" else L)
        final_segs = [(g_emb[0], u_idx, 0), (g_emb[1], i_idx, U), (g_emb[2], j_idx, U)]
        cl_segs, ego_segs = [], []
        cl_loss = None
        if self.model == "XSimGCL":
            nl, outs = ops.infonce_raw(
                [dict(table1=fin, table2=cl, idx=uq_u, n=nu, weight=self.cl_rate),
                 dict(table1=fin, table2=cl, idx=uq_i, n=ni, weight=self.cl_rate, row_off1=U, row_off2=U)], d, self.tau)
            cl_loss = nl
            final_segs += [(outs[0][0], uq_u, 0), (outs[1][0], uq_i, U)]
            tgt = cl_segs if 1 <= self.layer_cl <= L else ego_segs
            tgt += [(outs[0][1], uq_u, 0), (outs[1][1], uq_i, U)]
        else:
            ego_segs += [(g_l2[0], u_idx, 0), (g_l2[1], i_idx, U), (g_l2[2], j_idx, U)]
        # ---- Horner backward, rows sharded, every level pushed to all ranks ----
        def scatter(dst, segs, scale):
            for src, rows, off in segs:
                ops.scatter_add_rows(dst, src, rows, off, scale)

        acc = p.bufs[self.A0]
        acc.zero_()
        scatter(acc, final_segs, cm)
        if self.layer_cl == L:
            scatter(acc, cl_segs, 1.0)
        x_idx = self.A0
        for k in range(L - 1, 0, -1):
            y_idx = self.A1 if x_idx == self.A0 else self.A0
            p.spmm(p.bufs[x_idx], push_y=y_idx)
            p.barrier()
            y = p.bufs[y_idx]
            scatter(y, final_segs, cm)  # replicated: every rank adds the same sparse rows to its copy
            if self.layer_cl == k:
                scatter(y, cl_segs, 1.0)
            x_idx = y_idx
        extra = None
        if self.model == "LightGCN" or ego_segs:
            self.gd.zero_()
            if self.model == "LightGCN":
                scatter(self.gd, final_segs, cm)
            scatter(self.gd, ego_segs, 1.0)
            extra = self.gd
        epi = dict(adam_p=p.bufs[self.P], adam_m=self.m, adam_v=self.v, adam_scalars=self.scalars, beta1=0.9, beta2=0.999, adam_eps=1e-8)
        if extra is not None:
            epi["extra"] = extra
        p.spmm(p.bufs[x_idx], push_p=self.P, **epi)
        p.barrier()
        cl_val = (self.cl_rate * cl_loss.sum()) if cl_loss is not None else torch.zeros((), device=p.dev)
        self.losses = torch.stack([bl[0], bl[1], cl_val, bl[0] + bl[1] + cl_val])
