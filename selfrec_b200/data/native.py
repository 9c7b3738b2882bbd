"""Interaction built by the native dataset -> CSR builder (csrc/dataset.cpp).

load_interaction(conf, train_file, test_file) returns an object with the attributes and methods of
data/ui_graph.py's Interaction (same ids, same matrices bit for bit) without the Python dict loops of
FileIO.load_data_set + Interaction.__init__.  The name-keyed dict-of-dict members the CUDA path never
reads (training_set_u / training_set_i / training_data / test_data) are materialised on first access.
"""
import ctypes as C
from collections import defaultdict

import numpy as np
import scipy.sparse as sp

from .. import _lib
from .data import _PendingShuffles
from .graph import Graph
from .ui_graph import Interaction


def _names(lib, h, which, n, nbytes):
    blob = C.create_string_buffer(max(int(nbytes), 1))
    off = np.zeros(n + 1, dtype=np.int64)
    _lib.check(lib.srb_dataset_names(h, which, blob, off.ctypes.data_as(C.c_void_p)), "srb_dataset_names")
    raw = blob.raw[:int(nbytes)].decode()
    if len(raw) == int(nbytes):  # pure ASCII: byte offsets are character offsets
        o = off.tolist()
        return [raw[o[k]:o[k + 1]] for k in range(n)]
    b = blob.raw
    o = off.tolist()
    return [b[o[k]:o[k + 1]].decode() for k in range(n)]


class NativeInteraction(Interaction):
    """Interaction whose arrays come from srb_dataset_*; lazy name-keyed views."""

    def __init__(self, conf, train_file, test_file=None):  # noqa: super().__init__ deliberately not called
        lib = _lib.load()
        Graph.__init__(self)
        train_file = getattr(train_file, "path", train_file)  # TripleFile or a path
        test_file = getattr(test_file, "path", test_file)
        h = lib.srb_dataset_load(str(train_file).encode(), None if test_file is None else str(test_file).encode())
        if not h:
            raise _lib.SrbError("srb_dataset_load: " + lib.srb_last_error().decode())
        try:
            cnt = np.zeros(8, dtype=np.int64)
            _lib.check(lib.srb_dataset_counts(h, cnt.ctypes.data_as(C.c_void_p)), "srb_dataset_counts")
            U, I, n_tr, n_te, nnz, ub, ib, n_te_lines = (int(x) for x in cnt)
            unames, inames = _names(lib, h, 0, U, ub), _names(lib, h, 1, I, ib)
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            pu, pi, pw = np.empty(n_tr, np.int32), np.empty(n_tr, np.int32), np.empty(n_tr, np.float64)
            _lib.check(lib.srb_dataset_pairs(h, 0, p(pu), p(pi), p(pw)), "srb_dataset_pairs")
            tu, ti, tw = np.empty(n_te, np.int32), np.empty(n_te, np.int32), np.empty(n_te, np.float64)
            _lib.check(lib.srb_dataset_pairs(h, 1, p(tu), p(ti), p(tw)), "srb_dataset_pairs")
            rp, rc, rv = np.empty(U + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
            _lib.check(lib.srb_dataset_interaction_csr(h, p(rp), p(rc), p(rv)), "srb_dataset_interaction_csr")
            N = U + I
            ap, ac, av, rs = np.empty(N + 1, np.int32), np.empty(2 * nnz, np.int32), np.empty(2 * nnz, np.float32), np.empty(N, np.float32)
            _lib.check(lib.srb_dataset_adjacency_csr(h, None, p(ap), p(ac), p(av), p(rs)), "srb_dataset_adjacency_csr")
            # d = rowsum^-0.5 with numpy's float32 pow, exactly as normalize_graph_mat does (data/graph.py:13-15)
            with np.errstate(divide="ignore"):
                d_inv = np.power(rs, -0.5)
            d_inv[np.isinf(d_inv)] = 0.0
            nv = np.empty(2 * nnz, np.float32)
            ap2, ac2 = np.empty(N + 1, np.int32), np.empty(2 * nnz, np.int32)
            _lib.check(lib.srb_dataset_adjacency_csr(h, p(d_inv), p(ap2), p(ac2), p(nv), None), "srb_dataset_adjacency_csr")
        finally:
            lib.srb_dataset_free(h)
        self.config = conf
        self.user = dict(zip(unames, range(U)))
        self.item = dict(zip(inames, range(I)))
        self.id2user = dict(enumerate(unames))
        self.id2item = dict(enumerate(inames))
        self.user_num, self.item_num = U, I
        self.pair_users, self.pair_items, self.pair_weights = pu, pi, pw
        self._test_pairs = (tu, ti, tw)
        self._n_test_lines = n_te_lines
        self._unames, self._inames = unames, inames
        self.test_set = defaultdict(dict)
        self.test_set_item = set()
        for u, i in zip(tu.tolist(), ti.tolist()):  # ui_graph.py:42-45, file order
            self.test_set[unames[u]][inames[i]] = 1
        self.test_set_item = {inames[i] for i in np.unique(ti).tolist()}
        self.ui_adj = sp.csr_matrix((av, ac, ap), shape=(N, N))
        self.norm_adj = sp.csr_matrix((nv, ac2, ap2), shape=(N, N))
        self.interaction_mat = sp.csr_matrix((rv, rc, rp), shape=(U, I))
        self._rated = (rp.copy(), rc.copy())  # already sorted and unique
        self._lazy = {}
        self._td_pending = _PendingShuffles()

    # ---- name-keyed members of the reference object, built when somebody asks ----------------
    def _lazy_get(self, key, build):
        if key not in self._lazy:
            self._lazy[key] = build()
        return self._lazy[key]

    @property
    def training_data(self):
        rows = self._lazy.get("training_data")
        order = self._td_pending.take()
        if rows is None:  # first access: build the list, already in its current (shuffled) order
            pu, pi, pw = self.pair_users, self.pair_items, self.pair_weights
            if order is not None:
                pu, pi, pw = pu[order], pi[order], pw[order]
            un, inn = self._unames, self._inames
            rows = self._lazy["training_data"] = [[un[u], inn[i], w] for u, i, w in zip(pu.tolist(), pi.tolist(), pw.tolist())]
        elif order is not None:
            rows[:] = [rows[k] for k in order.tolist()]
        return rows

    @training_data.setter
    def training_data(self, v):
        self._lazy["training_data"] = v
        self._td_pending = _PendingShuffles()

    @property
    def test_data(self):
        tu, ti, tw = self._test_pairs
        un, inn = self._unames, self._inames
        return self._lazy_get("test_data", lambda: [[un[u], inn[i], w] for u, i, w in zip(tu.tolist(), ti.tolist(), tw.tolist())])

    @test_data.setter
    def test_data(self, v):
        self._lazy["test_data"] = v

    def _name_sets(self):
        def build():
            su, si = defaultdict(dict), defaultdict(dict)
            un, inn = self._unames, self._inames
            for u, i in zip(self.pair_users.tolist(), self.pair_items.tolist()):
                su[un[u]][inn[i]] = 1
                si[inn[i]][un[u]] = 1
            return su, si
        return self._lazy_get("sets", build)

    @property
    def training_set_u(self):
        return self._name_sets()[0]

    @property
    def training_set_i(self):
        return self._name_sets()[1]

    def training_size(self):
        return self.user_num, self.item_num, len(self.pair_users)

    def test_size(self):
        # the reference counts every test line, including those dropped for unseen users / items
        return len(self.test_set), len(self.test_set_item), self._n_test_lines


def load_interaction(conf, train_file, test_file=None):
    return NativeInteraction(conf, train_file, test_file)
