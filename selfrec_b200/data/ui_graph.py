"""Mirror of data/ui_graph.py:8-122 (class Interaction): same attributes and methods.

Besides the reference's dict/scipy members, the object carries what the CUDA path consumes:
  pair_users / pair_items   int32 internal ids of training_data, in file order
  rated_csr()               (ptr, idx) sorted rated item ids per user -> test-time mask
"""
from collections import defaultdict

import numpy as np
import scipy.sparse as sp

from .data import Data
from .graph import Graph


def _native_unit_laplacian(adj_mat):
    """convert_to_laplacian_mat for a U x I matrix of unit weights (SGL's edge-dropped graphs): CSR assembly by
    srb_bipartite_adjacency_csr, d = rowsum^-0.5 by numpy (what the reference calls), values (d_r * a) * d_c in
    fp32 -- bit-identical to the scipy route, ~5x faster.  None when the matrix does not qualify."""
    import ctypes as C
    from .. import _lib
    m = sp.csr_matrix(adj_mat)
    if m.nnz == 0 or m.dtype != np.float32 or not m.has_canonical_format or not np.all(m.data == 1.0):
        return None
    lib = _lib.load()  # a missing library is an error here as everywhere else
    U, I = m.shape
    n, k = U + I, int(m.nnz)
    rows = np.repeat(np.arange(U, dtype=np.int32), np.diff(m.indptr))
    cols = np.ascontiguousarray(m.indices, dtype=np.int32)
    ptr, idx, val, rs = np.empty(n + 1, np.int32), np.empty(2 * k, np.int32), np.empty(2 * k, np.float32), np.empty(n, np.float32)
    nnz = C.c_int64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.srb_bipartite_adjacency_csr(p(rows), p(cols), k, U, I, p(ptr), p(idx), p(val), p(rs), C.byref(nnz)),
               "srb_bipartite_adjacency_csr")
    idx, val = idx[: nnz.value], val[: nnz.value]
    with np.errstate(divide="ignore"):
        d_inv = np.power(rs, -0.5)  # graph.py:13-15
    d_inv[np.isinf(d_inv)] = 0.0
    r = np.repeat(np.arange(n), np.diff(ptr))
    data = ((d_inv[r] * val).astype(np.float32) * d_inv[idx]).astype(np.float32)  # left product, then right (graph.py:16-18)
    return sp.csr_matrix((data, idx, ptr), shape=(n, n))


class Interaction(Data, Graph):
    def __new__(cls, conf=None, training=None, test=None, *a, **kw):
        # triples that still know their file (FileIO.load_data_set) go to the native builder
        if cls is Interaction and getattr(training, "path", None) and (test is None or getattr(test, "path", None)):
            from .native import NativeInteraction
            return object.__new__(NativeInteraction)
        return object.__new__(cls)

    def __init__(self, conf, training, test):
        Graph.__init__(self)
        Data.__init__(self, conf, training, test)
        self.user = {}
        self.item = {}
        self.id2user = {}
        self.id2item = {}
        self.training_set_u = defaultdict(dict)
        self.training_set_i = defaultdict(dict)
        self.test_set = defaultdict(dict)
        self.test_set_item = set()
        self._index_sets()
        self.user_num = len(self.training_set_u)
        self.item_num = len(self.training_set_i)
        self.ui_adj = self._bipartite_adjacency()
        self.norm_adj = self.normalize_graph_mat(self.ui_adj)
        self.interaction_mat = self._interaction_matrix()
        self._rated = None

    def _index_sets(self):
        # ids in first-appearance order of the training file (ui_graph.py:29-40)
        user, item = self.user, self.item
        pu = np.empty(len(self.training_data), dtype=np.int32)
        pi = np.empty(len(self.training_data), dtype=np.int32)
        for k, (u, i, _r) in enumerate(self.training_data):
            uid = user.get(u)
            if uid is None:
                uid = len(user)
                user[u] = uid
                self.id2user[uid] = u
            iid = item.get(i)
            if iid is None:
                iid = len(item)
                item[i] = iid
                self.id2item[iid] = i
            self.training_set_u[u][i] = 1
            self.training_set_i[i][u] = 1
            pu[k] = uid
            pi[k] = iid
        self.pair_users, self.pair_items = pu, pi
        for u, i, _r in self.test_data:  # ui_graph.py:42-45
            if u in user and i in item:
                self.test_set[u][i] = 1
                self.test_set_item.add(i)

    def _bipartite_adjacency(self, self_connection=False):
        n = self.user_num + self.item_num
        ones = np.ones(len(self.pair_users), dtype=np.float32)
        # csr_matrix sums duplicate (u, i) lines (ui_graph.py:52-53)
        half = sp.csr_matrix((ones, (self.pair_users, self.pair_items.astype(np.int64) + self.user_num)), shape=(n, n), dtype=np.float32)
        adj = half + half.T
        if self_connection:
            adj += sp.eye(n)
        return adj

    def convert_to_laplacian_mat(self, adj_mat):
        # ui_graph.py:58-65: embed a U x I matrix into (U+I)^2 and normalise
        native = _native_unit_laplacian(adj_mat)
        if native is not None:
            return native
        rows, cols = adj_mat.nonzero()
        n = adj_mat.shape[0] + adj_mat.shape[1]
        half = sp.csr_matrix((adj_mat.data, (rows, cols + adj_mat.shape[0])), shape=(n, n), dtype=np.float32)
        return self.normalize_graph_mat(half + half.T)

    def _interaction_matrix(self):
        ones = np.ones(len(self.pair_users), dtype=np.float32)
        return sp.csr_matrix((ones, (self.pair_users, self.pair_items)), shape=(self.user_num, self.item_num), dtype=np.float32)

    def rated_csr(self):
        """(ptr[int32 U+1], idx[int32]) sorted unique rated item ids per user id."""
        if self._rated is None:
            m = sp.csr_matrix(self.interaction_mat)
            m.sum_duplicates()
            m.sort_indices()
            self._rated = (m.indptr.astype(np.int32), m.indices.astype(np.int32))
        return self._rated

    def test_csr(self):
        """(ptr[int32 U+1], idx[int32], n_test[int32 U]): per user id the sorted unique ids of its test items
        that have a training id, and len(test_set[user]) (which also counts items never seen in training)."""
        if getattr(self, "_test_csr", None) is None:
            U = len(self.user)
            rows, n_test = [[] for _ in range(U)], np.zeros(U, dtype=np.int32)
            for name, items in self.test_set.items():
                uid = self.user.get(name)
                if uid is None:
                    continue
                n_test[uid] = len(items)
                rows[uid] = sorted({self.item[i] for i in items if i in self.item})
            ptr = np.zeros(U + 1, dtype=np.int32)
            ptr[1:] = np.cumsum([len(r) for r in rows])
            idx = np.fromiter((i for r in rows for i in r), dtype=np.int32, count=int(ptr[-1]))
            self._test_csr = (ptr, idx, n_test)
        return self._test_csr

    def get_user_id(self, u):
        return self.user.get(u)

    def get_item_id(self, i):
        return self.item.get(i)

    def training_size(self):
        return len(self.user), len(self.item), len(self.training_data)

    def test_size(self):
        return len(self.test_set), len(self.test_set_item), len(self.test_data)

    def contain(self, u, i):
        return u in self.user and i in self.training_set_u[u]

    def contain_user(self, u):
        return u in self.user

    def contain_item(self, i):
        return i in self.item

    def user_rated(self, u):
        return list(self.training_set_u[u].keys()), list(self.training_set_u[u].values())

    def item_rated(self, i):
        return list(self.training_set_i[i].keys()), list(self.training_set_i[i].values())

    def row(self, u):
        vec = np.zeros(self.item_num, dtype=np.float32)
        for name, r in self.training_set_u[self.id2user[u]].items():
            vec[self.item[name]] = r
        return vec

    def col(self, i):
        vec = np.zeros(self.user_num, dtype=np.float32)
        for name, r in self.training_set_i[self.id2item[i]].items():
            vec[self.user[name]] = r
        return vec

    def matrix(self):
        m = np.zeros((self.user_num, self.item_num), dtype=np.float32)
        for name, uid in self.user.items():
            for it, r in self.training_set_u[name].items():
                m[uid, self.item[it]] = r
        return m
