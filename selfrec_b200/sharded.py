"""Bipartite-sharded multi-GPU path: one process per GPU.

The normalised adjacency is A = [[0, R], [R^T, 0]].  Rank g owns the USERS u with u % world == g (local row
u // world; their rows of every [U, d] table never leave the GPU); the ITEM tables are replicated.  Per propagation
layer only the item half crosses NVLink: every rank's partial product R_g^T X_u is stored by the SpMM epilogue into the
staging area of the rank owning that item slice, the owner adds the partials, applies the epilogue and stores the
finished rows into every rank's copy (selfrec_b200/csrc/sharded.cu).  torch.distributed is plumbing only: rendezvous
of the symmetric-memory region (peer pointers, multicast mapping); no NCCL collective touches the data path.

Why cyclic and not contiguous nnz-balanced blocks: ids follow first appearance in the training file
(ui_graph.py:29-40), so on a power-law graph the hubs have the low ids -- at config-5 size an nnz-balanced 2-way split
is 142 k users against 9.86 M, and the second rank's products (cold gathers, 70x the rows to write) take 1.5x longer.
Every rank taking each world-th user gets the same mix of degrees, rows and non-zeros.

Host logic here (user assignment, block extraction) is plain tensor code that also runs on CPU tensors and is
exercised with a world-size-2 gloo group in tests/test_sharding_cpu.py.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


def local_user_count(n_users, rank, world):
    """Users rank, rank + world, rank + 2 world, ... below n_users."""
    return (int(n_users) - int(rank) + int(world) - 1) // int(world)


def user_ids_of(n_users, rank, world):
    """Global ids of rank's users in local-row order (numpy int64)."""
    return np.arange(int(rank), int(n_users), int(world), dtype=np.int64)


def item_bounds(n_items, world):
    """Item slice whose reduction rank g owns: [g * I / world, (g + 1) * I / world) (sharded.cu)."""
    return np.array([g * n_items // world for g in range(world + 1)], dtype=np.int64)


def extract_blocks(rowptr, colidx, vals, n_users, n_items, rank, world):
    """Rank-local blocks of the normalised (U+I)^2 adjacency given as CSR tensors (any device), for the cyclic user
    assignment: Ru [Ug x I] = rows rank, rank + world, ... of A[:U, U:] (columns: item ids) and Rt [I x Ug] = the
    columns of A[U:, :U] with col % world == rank, renumbered col // world (monotone: the rows stay sorted).
    Returns two (rowptr, colidx, vals) triples of int32 / int32 / fp32 tensors."""
    import torch
    U, I, G, g = int(n_users), int(n_items), int(world), int(rank)
    rp = rowptr.to(torch.int64)
    dev = rp.device
    if G == 1:
        lo, hi = 0, int(rp[U])
        ru = (rp[:U + 1].to(torch.int32), (colidx[lo:hi] - U).to(torch.int32).contiguous(), vals[lo:hi].contiguous())
    else:
        rows = torch.arange(g, U, G, device=dev)
        beg = rp[rows]
        deg = rp[rows + 1] - beg
        ru_ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(deg, 0, out=ru_ptr[1:])
        total = int(ru_ptr[-1])
        pos = torch.arange(total, device=dev) + torch.repeat_interleave(beg - ru_ptr[:-1], deg, output_size=total)
        ru = (ru_ptr.to(torch.int32), (colidx[pos] - U).to(torch.int32).contiguous(), vals[pos].contiguous())
        del pos, beg, deg, rows
    ilo, ihi = int(rp[U]), int(rp[U + I])
    cols = colidx[ilo:ihi]
    if G == 1:
        rt = ((rp[U:U + I + 1] - ilo).to(torch.int32), cols.to(torch.int32).contiguous(), vals[ilo:ihi].contiguous())
        return ru, rt
    keep = (cols % G) == g
    pref = torch.zeros(ihi - ilo + 1, dtype=torch.int64, device=dev)
    torch.cumsum(keep, 0, out=pref[1:])
    rt_ptr = pref[rp[U:U + I + 1] - ilo].to(torch.int32)
    rt = (rt_ptr, torch.div(cols[keep], G, rounding_mode="floor").to(torch.int32).contiguous(), vals[ilo:ihi][keep].contiguous())
    return ru, rt


class ShardedEngine:
    """LightGCN / SimGCL / XSimGCL training on bipartite-sharded tables; world == 1 works without torch.distributed.

    Same constructor surface as TrainEngine.  Every rank must feed the SAME batch buffer to step().  Parameters:
    `user_emb` = this rank's users [Ug, d] (global ids `user_ids`: rank, rank + world, ...), `item_emb` = the full
    replicated [I, d] item table.  The in-kernel Philox noise is keyed by global row ids, so a sharded run with the same
    philox_seed draws the noise the single-GPU TrainEngine draws."""

    def __init__(self, model, data, emb_size, n_layers, batch_size, lr, reg, *, eps=0.0, tau=0.2, cl_rate=0.0, layer_cl=0,
                 l2_div=1.0, init_user=None, init_item=None, group=None, philox_seed=0x5EED, device=None, multicast=None, nvls=None):
        import torch
        from . import ops
        lib = _lib.require_device()
        if model not in ("LightGCN", "SimGCL", "XSimGCL"):
            raise _lib.SrbError("the sharded engine covers LightGCN, SimGCL and XSimGCL")
        if int(emb_size) not in ops._SUPPORTED_D:
            raise _lib.SrbError(f"embedding.size {emb_size} is not supported by the CUDA path {ops._SUPPORTED_D}")
        self.torch, self.ops, self.lib = torch, ops, lib
        self.model_name = model
        self.dist = None
        self.group = None
        self.rank, self.world = 0, 1
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.dist = dist
            self.group = dist.group.WORLD if group is None else group
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise _lib.SrbError("the sharded engine supports up to 8 ranks (one NVSwitch domain)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.dev
        self.U, self.I, self.d = int(data.user_num), int(data.item_num), int(emb_size)
        self.N, self.L, self.B = self.U + self.I, int(n_layers), int(batch_size)
        # ---- graph blocks of this rank ----
        na = data.norm_adj
        adj = na if isinstance(na, ops.SparseAdj) else ops.SparseAdj(na)
        had = adj.rowptr is not None
        adj.cuda(dev)
        if self.U < self.world:
            raise _lib.SrbError(f"{self.U} users cannot be spread over {self.world} ranks")
        self.Ug = local_user_count(self.U, self.rank, self.world)
        self.user_ids = torch.arange(self.rank, self.U, self.world, device=dev)  # global id of every local row
        ru, rt = extract_blocks(adj.rowptr, adj.colidx, adj.vals, self.U, self.I, self.rank, self.world)
        self.Ru = ops.SparseAdj.from_device(*ru, (self.Ug, self.I), symmetric=False)
        self.Rt = ops.SparseAdj.from_device(*rt, (self.I, self.Ug), symmetric=False)
        self.nnzA = adj.nnz
        if not had and self.world > 1:
            adj.rowptr = adj.colidx = adj.vals = adj.row_order = None  # the full matrix is not needed on the device any more
        self.ib = item_bounds(self.I, self.world)
        # ---- memory ----
        lay = _lib.ShardLayout()
        gu, gt = self.Ru.graph_struct(self.d), self.Rt.graph_struct(self.d)
        _lib.check(lib.srb_shard_plan(self.U, self.I, self.Ug, self.d, self.B, self.world, int(gu.hub.n_work), int(gt.hub.n_work),
                                      C.byref(lay)), "srb_shard_plan")
        self.layout = lay
        self.sym_handle = None
        mc_ptr = 0
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.sym = symm.empty(int(lay.sym_bytes), dtype=torch.uint8, device=dev)
            self.sym_handle = symm.rendezvous(self.sym, self.group)
            peers = [int(p) for p in self.sym_handle.buffer_ptrs]
            want = os.environ.get("SRB_MULTICAST", "auto") if multicast is None else ("1" if multicast else "0")
            mc = int(getattr(self.sym_handle, "multicast_ptr", 0) or 0)
            if mc and (want == "1" or (want == "auto" and self.world >= 4)):
                mc_ptr = mc
        else:
            self.sym = torch.empty(int(lay.sym_bytes), dtype=torch.uint8, device=dev)
            peers = [self.sym.data_ptr()]
        self.sym.zero_()
        self.use_multicast = bool(mc_ptr)
        self.workspace = torch.zeros(int(lay.workspace_bytes) + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
        self._ctrl = self.workspace[ws_ptr - self.workspace.data_ptr() + int(lay.ctrl):][:8].view(torch.int32)
        nd_i = self.I * self.d
        self.item_emb = self.sym[int(lay.item_params): int(lay.item_params) + 4 * nd_i].view(torch.float32).view(self.I, self.d)
        self._item_final = self.sym[int(lay.item_final): int(lay.item_final) + 4 * nd_i].view(torch.float32).view(self.I, self.d)
        self.user_emb = torch.empty((self.Ug, self.d), device=dev, dtype=torch.float32)
        if init_user is None:
            if self.N * self.d > (1 << 27):  # config-5 sized tables: drawn on the device, identically on every rank
                g = torch.Generator(device=dev).manual_seed(int(philox_seed) & 0x7FFFFFFF)
                bound_u, bound_i = (6.0 / (self.U + self.d)) ** 0.5, (6.0 / (self.I + self.d)) ** 0.5
                chunk = 1 << 20
                for lo in range(0, self.U, chunk):  # the same stream on every rank; keep the owned rows
                    hi = min(self.U, lo + chunk)
                    blk = torch.empty((hi - lo, self.d), device=dev).uniform_(-bound_u, bound_u, generator=g)
                    first = lo + (self.rank - lo) % self.world  # first owned id >= lo
                    if first < hi:
                        mine = blk[first - lo:: self.world]
                        self.user_emb[first // self.world: first // self.world + mine.shape[0]].copy_(mine)
                self.item_emb.uniform_(-bound_i, bound_i, generator=g)
            else:
                g = torch.Generator().manual_seed(int(philox_seed) & 0x7FFFFFFF)  # every rank starts from the same tables
                init_user = torch.nn.init.xavier_uniform_(torch.empty(self.U, self.d), generator=g)
                init_item = torch.nn.init.xavier_uniform_(torch.empty(self.I, self.d), generator=g)
        if init_user is not None:
            self.user_emb.copy_(torch.as_tensor(init_user)[self.rank:: self.world])
            self.item_emb.copy_(torch.as_tensor(init_item))
        self.mu, self.vu = torch.zeros_like(self.user_emb), torch.zeros_like(self.user_emb)
        self.mi = torch.zeros((self.I, self.d), device=dev)
        self.vi = torch.zeros((self.I, self.d), device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(16, device=dev)
        self.losses = torch.zeros(4, device=dev)
        self.words = _lib.BATCH_HEADER + 5 * self.B
        self.batch_dev = torch.zeros(self.words, dtype=torch.int32, device=dev)
        s = _lib.ShardDesc()
        s.model, s.world, s.rank = _lib.MODEL_IDS[model], self.world, self.rank
        s.n_users, s.n_items, s.d, s.n_layers, s.batch_cap, s.layer_cl = self.U, self.I, self.d, self.L, self.B, int(layer_cl)
        s.eps, s.tau, s.cl_rate, s.reg = float(eps), float(tau), float(cl_rate), float(reg)
        s.lr, s.beta1, s.beta2, s.adam_eps, s.l2_div = float(lr), 0.9, 0.999, 1e-8, float(l2_div)
        s.noise_mode = 2 if model in ("SimGCL", "XSimGCL") else 0
        s.philox_seed = int(philox_seed)
        s.Ru, s.Rt = gu, gt
        p = ops._p
        s.batch, s.pu, s.mu, s.vu, s.mi, s.vi = p(self.batch_dev), p(self.user_emb), p(self.mu), p(self.vu), p(self.mi), p(self.vi)
        s.step_dev, s.scalars, s.losses = p(self.step_dev), p(self.scalars), p(self.losses)
        for g in range(self.world):
            s.sym[g] = peers[g]
        s.sym_mc = mc_ptr or None
        s.sym_bytes = int(lay.sym_bytes)
        s.workspace, s.workspace_bytes = C.c_void_p(ws_ptr), int(lay.workspace_bytes)
        if self.world > 1 and os.environ.get("SRB_SHARD_OVERLAP", "1") != "0":
            # the owner-side reduction of a layer runs on this stream beside the user-side product
            self._fork_stream = torch.cuda.Stream(device=dev)
            self._fork_events = (torch.cuda.Event(), torch.cuda.Event())
            for ev in self._fork_events:
                ev.record(self._fork_stream)  # torch creates the CUDA event lazily, on first record
            s.fork_stream = C.c_void_p(self._fork_stream.cuda_stream)
            s.fork_event, s.join_event = (C.c_void_p(ev.cuda_event) for ev in self._fork_events)
        want_nvls = (os.environ.get("SRB_SHARD_NVLS", "0") == "1") if nvls is None else bool(nvls)
        s.nvls = 1 if (want_nvls and mc_ptr) else 0
        self.use_nvls = bool(s.nvls)
        self.desc = s
        self.graph = None
        self._warm = False
        torch.cuda.synchronize()
        self._host_barrier()

    # ---- plumbing ------------------------------------------------------------------------
    def _host_barrier(self):
        if self.dist is not None and self.world > 1:
            self.dist.barrier(self.group)

    def check_peers(self):
        """Raise if a device-side barrier ever timed out (a peer died or fell out of step)."""
        if int(self._ctrl[1].item()) != 0:
            raise _lib.SrbError("sharded step: a peer rank did not reach a device-side barrier in time")

    def nvlink_bytes_per_layer(self):
        """Bytes this rank sends per propagation layer: the partial rows it hands to the other slices' owners plus
        the finished rows of its own slice (one multicast store, or one store per peer)."""
        if self.world == 1:
            return 0
        own = int(self.ib[self.rank + 1] - self.ib[self.rank]) * self.d * 4
        part = self.I * self.d * 4 - own
        return part + own * (1 if self.use_multicast else self.world - 1)

    # ---- stepping ------------------------------------------------------------------------
    def _enqueue(self):
        _lib.check(self.lib.srb_shard_step(C.byref(self.desc), self.ops._stream()), "srb_shard_step")

    def step(self, words=None, words_dev=None):
        torch = self.torch
        if words_dev is not None:
            self.batch_dev.copy_(words_dev, non_blocking=True)
        elif words is not None:
            self.batch_dev.copy_(torch.as_tensor(np.asarray(words, dtype=np.int32)), non_blocking=True)
        self.step_resident()

    def step_resident(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()

    def capture(self):
        """CUDA graph of one step (device-side barriers included).  Collective: every rank must call it."""
        torch = self.torch
        torch.cuda.synchronize()
        state = (self.user_emb, self.item_emb, self.mu, self.vu, self.mi, self.vi, self.step_dev, self.losses)
        if not self._warm:  # a warm-up is a real step: put the trajectory back afterwards (all ranks do the same)
            saved = [t.clone() for t in state]
            self._host_barrier()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._enqueue()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._host_barrier()
            for dst, src in zip(state, saved):
                dst.copy_(src)
            del saved
            torch.cuda.synchronize()
            self._host_barrier()
            self._warm = True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue()
        self.graph = g
        self._host_barrier()
        return g

    # ---- inference -------------------------------------------------------------------------
    def forward_clean(self):
        """Clean forward -> (final embeddings of this rank's users [Ug, d], final item embeddings [I, d])."""
        torch = self.torch
        out_u = torch.empty_like(self.user_emb)
        _lib.check(self.lib.srb_shard_forward(C.byref(self.desc), self.ops._p(out_u), self.ops._stream()), "srb_shard_forward")
        return out_u, self._item_final.clone()

    def all_user_rows(self, local):
        """[U, d] table from every rank's [Ug, d] block (test / evaluation plumbing: one NCCL all_gather)."""
        torch = self.torch
        if self.world == 1:
            return local.clone()
        sizes = [local_user_count(self.U, g, self.world) for g in range(self.world)]
        pad = torch.zeros((max(sizes), self.d), device=self.dev)
        pad[: self.Ug].copy_(local)
        outs = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(outs, pad, group=self.group)
        full = torch.empty((self.U, self.d), device=self.dev)
        for g, (o, n) in enumerate(zip(outs, sizes)):
            full[g:: self.world].copy_(o[:n])
        return full
