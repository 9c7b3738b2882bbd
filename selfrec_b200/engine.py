"""TrainEngine: device state + one fused training step per call (srb_train_step).

Owns (as torch tensors, i.e. PyTorch's allocator) the single contiguous [U+I, d] parameter
table -- users first, items after, so the reference's torch.cat (LightGCN.py:69) disappears
-- the Adam moments, the step counter, the workspace and the device batch buffer.  step()
enqueues one H2D copy of the batch words plus the whole forward/backward/Adam sequence on
the current stream; nothing synchronises.  capture() wraps the same sequence in a CUDA graph.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from .util.sampler import NativePairSampler, permute_training_data, stream_epoch


class LossHandle:
    """Pinned-memory copy of a step's [rec, l2, cl, total] losses; get() waits for the D2H copy."""

    __slots__ = ("_buf", "_ev")

    def __init__(self, buf, ev):
        self._buf, self._ev = buf, ev

    def get(self):
        self._ev.synchronize()
        return self._buf.numpy().copy()


class TrainEngine:
    def __init__(self, model, data, emb_size, n_layers, batch_size, lr, reg, *, eps=0.0, tau=0.2, cl_rate=0.0,
                 layer_cl=0, l2_div=1.0, device=None, init_user=None, init_item=None, philox_seed=0x5EED):
        lib = _lib.require_device()
        self.lib = lib
        if model not in _lib.MODEL_IDS:
            raise ValueError(f"TrainEngine: unknown model {model!r} (one of {sorted(_lib.MODEL_IDS)})")
        if int(emb_size) not in ops._SUPPORTED_D:
            raise _lib.SrbError(f"TrainEngine: embedding.size {emb_size} is not supported by the CUDA path "
                                f"(supported: {ops._SUPPORTED_D}); there is no fallback")
        if int(batch_size) <= 0:
            raise ValueError("TrainEngine: batch.size must be positive")
        self.model_name = model
        self.model_id = _lib.MODEL_IDS[model]
        self.data = data
        self.U, self.I, self.d = int(data.user_num), int(data.item_num), int(emb_size)
        self.N = self.U + self.I
        self.L = int(n_layers) if model != "MF" else 0
        self.B = int(batch_size)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.dev
        self.params = torch.empty((self.N, self.d), device=dev, dtype=torch.float32)
        if init_user is None and self.N * self.d > (1 << 27):
            # config-5 sized tables: xavier-uniform drawn on the device (a 6 GB host tensor is not worth its copy)
            g = torch.Generator(device=dev).manual_seed(torch.initial_seed() & 0x7FFFFFFF)
            for lo, hi in ((0, self.U), (self.U, self.N)):
                bound = (6.0 / ((hi - lo) + self.d)) ** 0.5
                self.params[lo:hi].uniform_(-bound, bound, generator=g)
        else:
            if init_user is None:  # same initialiser calls, same order as LightGCN.py:60-66
                init_user = torch.nn.init.xavier_uniform_(torch.empty(self.U, self.d))
                init_item = torch.nn.init.xavier_uniform_(torch.empty(self.I, self.d))
            self.params[: self.U].copy_(init_user)
            self.params[self.U:].copy_(init_item)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.scalars = torch.zeros(16, device=dev, dtype=torch.float32)
        self.losses = torch.zeros(4, device=dev, dtype=torch.float32)
        self.words = _lib.BATCH_HEADER + 5 * self.B
        self.batch_dev = torch.zeros(self.words, device=dev, dtype=torch.int32)
        self.ring = [torch.zeros(self.words, dtype=torch.int32).pin_memory() for _ in range(8)]
        self.ring_ev = [None] * len(self.ring)
        self.ring_pos = 0
        self.loss_ring = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(8)]
        self.loss_pos = 0
        self.adj = None
        if model != "MF":
            na = data.norm_adj
            self.adj = na if isinstance(na, ops.SparseAdj) else ops.SparseAdj(na)
            self.adj.cuda(dev)
        ws_bytes = lib.srb_step_workspace_bytes(self.model_id, self.N, self.d, self.B, self.adj.hub_struct(self.d).n_work if self.adj is not None else 0)
        self.workspace = torch.empty(ws_bytes + 256, device=dev, dtype=torch.uint8)
        ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
        self.view_adj = [None, None]
        self.noise = None
        s = _lib.StepDesc()
        s.model, s.n_users, s.n_items, s.d, s.n_layers = self.model_id, self.U, self.I, self.d, self.L
        s.batch_cap, s.layer_cl = self.B, int(layer_cl)
        s.eps, s.tau, s.cl_rate, s.reg = float(eps), float(tau), float(cl_rate), float(reg)
        s.lr, s.beta1, s.beta2, s.adam_eps = float(lr), 0.9, 0.999, 1e-8
        s.l2_div = float(l2_div)
        s.noise_mode = 2 if model in ("SimGCL", "XSimGCL") else 0
        s.philox_seed = int(philox_seed)
        if self.adj is not None:
            s.adj = self.adj.graph_struct(self.d)
        s.batch, s.params, s.adam_m, s.adam_v = ops._p(self.batch_dev), ops._p(self.params), ops._p(self.m), ops._p(self.v)
        s.step_dev, s.scalars, s.losses = ops._p(self.step_dev), ops._p(self.scalars), ops._p(self.losses)
        s.workspace, s.workspace_bytes = C.c_void_p(ws_ptr), ws_bytes
        # this engine's own fork / join resources (BPR beside InfoNCE): two engines on one device never share events
        self._fork_stream = torch.cuda.Stream(device=self.dev)
        self._fork_events = (torch.cuda.Event(), torch.cuda.Event())
        for ev in self._fork_events:
            ev.record(self._fork_stream)  # torch creates the CUDA event lazily, on first record
        s.fork_stream = C.c_void_p(self._fork_stream.cuda_stream)
        s.fork_event, s.join_event = (C.c_void_p(ev.cuda_event) for ev in self._fork_events)
        self.desc = s
        self.eps, self.layer_cl = float(eps), int(layer_cl)
        self.sampler = None
        self.graph = None
        self._warm = False

    # ---- parameters as the reference exposes them ------------------------------------
    @property
    def user_emb(self):
        return self.params[: self.U]

    @property
    def item_emb(self):
        return self.params[self.U:]

    # ---- configuration -----------------------------------------------------------------
    def set_noise_tensor(self, noise):
        """Parity mode: noise is an INPUT, uniform[0,1) of shape [views, L, N, d]."""
        noise = ops._f32c(noise, "noise")
        views = 2 if self.model_name == "SimGCL" else 1
        if tuple(noise.shape) != (views, self.L, self.N, self.d):
            raise ValueError(f"noise must be [{views}, {self.L}, {self.N}, {self.d}]")
        self.noise = noise
        self.desc.noise_mode, self.desc.noise = 1, ops._p(noise)
        self.graph = None  # the step sequence changed: a captured graph is stale

    def set_view_graphs(self, adj1, adj2):
        """SGL: the two dropped, re-normalised graphs of this epoch (SGL.py:27-29)."""
        self.view_adj = [a if isinstance(a, ops.SparseAdj) else ops.SparseAdj(a) for a in (adj1, adj2)]
        for k, a in enumerate(self.view_adj):
            a.cuda(self.dev)
            self.desc.adj_view[k] = a.graph_struct(self.d)
        self.graph = None  # pointers changed: a captured graph is stale

    # ---- stepping ------------------------------------------------------------------------
    def _enqueue(self):
        _lib.check(self.lib.srb_train_step(C.byref(self.desc), ops._stream()), "srb_train_step")

    def step(self, batch_words, fetch_loss=False):
        """batch_words: int32 array/tensor of `words` entries laid out by srb_sampler_next_batch.
        Enqueues the H2D copy and the step (the captured CUDA graph when capture() was called) and
        returns without synchronising.  fetch_loss=True also enqueues a D2H copy of the four loss
        values into pinned memory and returns a LossHandle; .get() waits for that copy only, so the
        caller can sample the next batch while this step runs."""
        slot = self.ring_pos
        self.ring_pos = (slot + 1) % len(self.ring)
        ev = self.ring_ev[slot]
        if ev is not None:
            ev.synchronize()  # the copy that last used this pinned slot has finished
        pin = self.ring[slot]
        if isinstance(batch_words, torch.Tensor):
            pin.copy_(batch_words)
        else:
            pin.numpy()[:] = batch_words
        self.batch_dev.copy_(pin, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.ring_ev[slot] = ev
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        if not fetch_loss:
            return None
        ls = self.loss_pos
        self.loss_pos = (ls + 1) % len(self.loss_ring)
        self.loss_ring[ls].copy_(self.losses, non_blocking=True)
        lev = torch.cuda.Event()
        lev.record()
        return LossHandle(self.loss_ring[ls], lev)

    def step_resident(self):
        """Step on whatever batch_dev currently holds (inputs already in HBM)."""
        self._enqueue()

    def capture(self):
        """CUDA graph of one step on the resident batch buffer; replay with graph.replay()."""
        torch.cuda.synchronize()
        if not self._warm:
            # warm-up outside capture (lazy module load, smem attributes).  A warm-up IS a training step on whatever
            # batch_dev holds: parameters, moments and the step counter (Adam bias correction, Philox stream) are
            # put back afterwards, so capturing never changes the training trajectory.
            saved = (self.params.clone(), self.m.clone(), self.v.clone(), self.step_dev.clone(), self.losses.clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._enqueue()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for dst, src in zip((self.params, self.m, self.v, self.step_dev, self.losses), saved):
                dst.copy_(src)
            del saved
            torch.cuda.synchronize()
            self._warm = True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue()
        self.graph = g
        return g

    def batches(self, exact_lazy=False):
        """One epoch of batch words from the native sampler (advances Python's `random`).  The yielded buffer is
        reused: consume it (step() copies it into a pinned slot) before asking for the next one.  exact_lazy=True
        hands Python's `random` state back after every batch, like the reference's generator would."""
        if self.sampler is None:
            self.sampler = NativePairSampler(self.data)
        s = self.sampler
        if not exact_lazy:
            yield from stream_epoch(s, self.data, self.B, self.B)
            return
        s.pull_state()
        perm = s.begin_epoch(want_perm=True)
        permute_training_data(self.data, perm)
        s.push_state()
        buf = np.empty(self.words, dtype=np.int32)
        while True:
            s.pull_state()
            b = s.next_batch(self.B, self.B, buf)
            s.push_state()
            if b == 0:
                return
            yield buf

    # ---- inference ---------------------------------------------------------------------
    def forward_clean(self):
        """no_grad clean forward -> (user_emb, item_emb), e.g. XSimGCL.py:40-41."""
        if self.model_name == "MF":
            out = self.params.clone()
        else:
            include_ego = self.model_name in ("LightGCN", "SGL")
            out, _ = ops.encoder_forward(self.adj, self.params, self.L, include_ego)
        return out[: self.U], out[self.U:]
