"""Interaction graph resident on the GPU + device-side assembly of the normalised adjacency.

`DeviceBipartite` holds the users x items CSR of the distinct training pairs and its transpose on the device
and turns a set of kept edges into the normalised (U+I)^2 adjacency with `srb_graph_assemble`
(selfrec_b200/csrc/graphbuild.cu) -- bit-identical to the scipy route of the reference:

    Interaction.__create_sparse_bipartite_adjacency / convert_to_laplacian_mat   data/ui_graph.py:47-65
    Graph.normalize_graph_mat                                                   data/graph.py:10-24
    GraphAugmentor.edge_dropout / node_dropout                                   data/augmentor.py:11-40

Used by SGL's per-epoch views (the draw stays CPython's random.sample stream; only the kept positions travel)
and by the config-5 sized synthetic graph (10 M x 2 M x 200 M: no scipy, no host copy of the adjacency).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, ops


def dinv_table(max_rowsum):
    """float32 power(k, -0.5) for k = 0 .. max_rowsum with inf -> 0: numpy's rounding, as graph.py:13-15 gets it."""
    k = np.arange(int(max_rowsum) + 1, dtype=np.float32)
    with np.errstate(divide="ignore"):
        t = np.power(k, -0.5)
    t[np.isinf(t)] = 0.0
    return t.astype(np.float32)


class DeviceBipartite:
    def __init__(self, n_users, n_items, ui_ptr, ui_col, ui_val, iu_ptr, iu_col, iu_perm):
        _lib.require_device()
        self.U, self.I = int(n_users), int(n_items)
        self.N = self.U + self.I
        self.ui_ptr, self.ui_col, self.ui_val = ui_ptr, ui_col, ui_val
        self.iu_ptr, self.iu_col, self.iu_perm = iu_ptr, iu_col, iu_perm
        self.nnz = int(ui_col.numel())
        self.dev = ui_ptr.device
        if self.nnz >= 1 << 30:
            raise _lib.SrbError("DeviceBipartite: 2 * nnz must fit in int32")
        # row sums are bounded by the largest (weighted) degree on either side
        w = ui_val if ui_val is not None else None
        deg_u = (ui_ptr[1:] - ui_ptr[:-1]).max().item() if self.nnz else 0
        deg_i = (iu_ptr[1:] - iu_ptr[:-1]).max().item() if self.nnz else 0
        wmax = int(w.max().item()) if (w is not None and self.nnz) else 1
        self.table = torch.from_numpy(dinv_table(max(deg_u, deg_i) * wmax)).to(self.dev)
        lib = _lib.load()
        self._ws_bytes = int(lib.srb_graph_assemble_workspace_bytes(self.U, self.I, self.nnz))
        self._ws = None

    # ---- constructors ----------------------------------------------------------------------
    @classmethod
    def from_interaction_mat(cls, mat, device):
        """From the reference's data.interaction_mat (scipy, users x items, duplicates already summed)."""
        import scipy.sparse as sp
        m = sp.csr_matrix(mat, dtype=np.float32)
        m.sum_duplicates()
        m.sort_indices()
        dev = torch.device(device)
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
        return cls.from_device_csr(m.shape[0], m.shape[1], to(m.indptr, np.int32), to(m.indices, np.int32),
                                   None if np.all(m.data == 1.0) else to(m.data, np.float32))

    @classmethod
    def from_device_csr(cls, n_users, n_items, ui_ptr, ui_col, ui_val=None):
        """From a device users x items CSR (columns ascending within a row).  The transpose comes from one stable
        device sort by item id (setup code; users stay ascending within an item's row)."""
        nnz = ui_col.numel()
        order = torch.sort(ui_col.to(torch.int64), stable=True).indices  # iu order -> position in the ui order
        users_of = torch.repeat_interleave(torch.arange(n_users, device=ui_ptr.device, dtype=torch.int32),
                                           (ui_ptr[1:] - ui_ptr[:-1]).to(torch.int64), output_size=nnz)
        iu_col = users_of[order].contiguous()
        cnt = torch.bincount(ui_col.to(torch.int64), minlength=n_items)
        iu_ptr = torch.zeros(n_items + 1, dtype=torch.int32, device=ui_ptr.device)
        iu_ptr[1:] = torch.cumsum(cnt, 0).to(torch.int32)
        return cls(n_users, n_items, ui_ptr.contiguous(), ui_col.contiguous(), ui_val, iu_ptr, iu_col, order.to(torch.int32).contiguous())

    # ---- assembly ----------------------------------------------------------------------------
    def assemble(self, keep_idx=None, keep_flags=None, reset_weights=False, out=None):
        """Normalised adjacency of the kept edges as an ops.SparseAdj on the device.
        keep_idx: int64 positions in the users x items CSR order (e.g. sample_range's output, host or device);
        keep_flags: uint8 device tensor over that order; neither: all edges.
        out: optional (rowptr, colidx, vals) device buffers to fill (capacity >= 2 * kept): fixed addresses across
        epochs keep a captured CUDA graph of the training step valid."""
        lib = _lib.load()
        dev = self.dev
        if self._ws is None:
            self._ws = torch.empty(self._ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
        g = _lib.GraphAssembleDesc()
        g.n_users, g.n_items, g.nnz = self.U, self.I, self.nnz
        g.ui_ptr, g.ui_col, g.ui_val = ops._p(self.ui_ptr), ops._p(self.ui_col), ops._p(self.ui_val)
        g.iu_ptr, g.iu_col, g.iu_perm = ops._p(self.iu_ptr), ops._p(self.iu_col), ops._p(self.iu_perm)
        kept = self.nnz
        keep = []
        if keep_idx is not None:
            ki = keep_idx if isinstance(keep_idx, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(keep_idx, dtype=np.int64))
            ki = ki.to(device=dev, dtype=torch.int64).contiguous()
            keep.append(ki)
            kept = int(ki.numel())
            g.keep_idx, g.n_keep = ops._p(ki), kept
        elif keep_flags is not None:
            kf = keep_flags.to(device=dev, dtype=torch.uint8).contiguous()
            keep.append(kf)
            g.keep_flags = ops._p(kf)
        g.reset_weights = int(bool(reset_weights))
        g.dinv_table, g.dinv_table_n = ops._p(self.table), int(self.table.numel())
        if out is None:
            cap = 2 * kept
            rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=dev)
            colidx = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
            vals = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
        else:
            rowptr, colidx, vals = out
            cap = int(colidx.numel())
        dinv = torch.empty(self.N, dtype=torch.float32, device=dev)
        nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
        g.rowptr, g.colidx, g.vals, g.dinv = ops._p(rowptr), ops._p(colidx), ops._p(vals), ops._p(dinv)
        g.out_cap, g.nnz_out = cap, ops._p(nnz_out)
        g.workspace, g.workspace_bytes = C.c_void_p(ws_ptr), self._ws_bytes
        _lib.check(lib.srb_graph_assemble(C.byref(g), ops._stream()), "srb_graph_assemble")
        n_out = int(nnz_out.item()) if keep_flags is not None else 2 * kept
        adj = ops.SparseAdj.from_device(rowptr, colidx[:n_out], vals[:n_out], (self.N, self.N), symmetric=True)
        adj.dinv = dinv
        return adj

    def free_workspace(self):
        self._ws = None
