"""Synthetic bipartite interaction graphs of a named |U| x |I| x nnz shape.

No dataset travels to the GPU box, so bench.py and the full-size property tests build a
power-law graph with the same shape as the named configuration (yelp2018: 31 668 x 38 048 x
1 237 259).  Ids follow the reference's convention: first-appearance order of the pair list
(ui_graph.py:29-40), every user and item has at least one interaction, pairs are unique.
"""
import numpy as np
import scipy.sparse as sp

from .data.data import Data
from .data.graph import Graph

SHAPES = {
    "douban-book": (12638, 22222, 478730),
    "yelp2018": (31668, 38048, 1237259),
    "amazon-kindle": (138333, 98572, 1525091),
}


def _first_appearance_relabel(x, n):
    _, first = np.unique(x, return_index=True)
    order = np.argsort(first)  # old ids in order of first appearance
    new = np.empty(n, dtype=np.int64)
    new[np.unique(x)[order]] = np.arange(len(order))
    return new[x]


def make_pairs(n_users, n_items, nnz, seed=0, alpha_u=0.42, alpha_i=0.40):
    """Unique (user, item) id pairs, power-law degrees on both sides, shuffled by user blocks."""
    rng = np.random.default_rng(seed)
    pu = (1.0 / np.arange(1, n_users + 1) ** alpha_u)
    pi = (1.0 / np.arange(1, n_items + 1) ** alpha_i)
    pu /= pu.sum()
    pi /= pi.sum()
    # one guaranteed edge per node, then draw until nnz unique pairs
    u0 = np.concatenate([np.arange(n_users), rng.choice(n_users, n_items, p=pu)])
    i0 = np.concatenate([rng.choice(n_items, n_users, p=pi), np.arange(n_items)])
    key = set()
    keys = u0.astype(np.int64) * n_items + i0
    keys = np.unique(keys)
    while len(keys) < nnz:
        need = int((nnz - len(keys)) * 1.3) + 1024
        k = rng.choice(n_users, need, p=pu).astype(np.int64) * n_items + rng.choice(n_items, need, p=pi)
        keys = np.unique(np.concatenate([keys, k]))
    if len(keys) > nnz:
        # drop random extras but never a node's only edge
        u, i = keys // n_items, keys % n_items
        du, di = np.bincount(u, minlength=n_users), np.bincount(i, minlength=n_items)
        perm = rng.permutation(len(keys))
        keep = np.ones(len(keys), dtype=bool)
        extra = len(keys) - nnz
        for t in perm:
            if extra == 0:
                break
            if du[u[t]] > 1 and di[i[t]] > 1:
                keep[t] = False
                du[u[t]] -= 1
                di[i[t]] -= 1
                extra -= 1
        keys = keys[keep]
    u, i = keys // n_items, keys % n_items
    # the training file is grouped by user; users in random order
    uperm = rng.permutation(n_users)
    order = np.lexsort((rng.random(len(u)), uperm[u]))
    u, i = u[order], i[order]
    return _first_appearance_relabel(u, n_users).astype(np.int32), _first_appearance_relabel(i, n_items).astype(np.int32)


class ArrayInteraction(Data, Graph):
    """The subset of data/ui_graph.py's Interaction the CUDA path consumes, built straight
    from id arrays (no name dictionaries).  training_data holds (user_id, item_id, 1.0)."""

    def __init__(self, pair_users, pair_items, n_users, n_items, test_users=None):
        self.pair_users = np.ascontiguousarray(pair_users, dtype=np.int32)
        self.pair_items = np.ascontiguousarray(pair_items, dtype=np.int32)
        self.user_num, self.item_num = int(n_users), int(n_items)
        Data.__init__(self, None, list(zip(self.pair_users.tolist(), self.pair_items.tolist())), [])
        n = self.user_num + self.item_num
        ones = np.ones(len(self.pair_users), dtype=np.float32)
        half = sp.csr_matrix((ones, (self.pair_users, self.pair_items.astype(np.int64) + self.user_num)), shape=(n, n), dtype=np.float32)
        self.ui_adj = half + half.T
        self.norm_adj = self.normalize_graph_mat(self.ui_adj)
        self.interaction_mat = sp.csr_matrix((ones, (self.pair_users, self.pair_items)), shape=(self.user_num, self.item_num), dtype=np.float32)
        self.test_user_ids = np.arange(self.user_num, dtype=np.int32) if test_users is None else np.asarray(test_users, dtype=np.int32)
        self._rated = None

    def rated_csr(self):
        if self._rated is None:
            m = sp.csr_matrix(self.interaction_mat)
            m.sum_duplicates()
            m.sort_indices()
            self._rated = (m.indptr.astype(np.int32), m.indices.astype(np.int32))
        return self._rated

    def convert_to_laplacian_mat(self, adj_mat):
        rows, cols = adj_mat.nonzero()
        n = adj_mat.shape[0] + adj_mat.shape[1]
        half = sp.csr_matrix((adj_mat.data, (rows, cols + adj_mat.shape[0])), shape=(n, n), dtype=np.float32)
        return self.normalize_graph_mat(half + half.T)


def make_interaction(shape="yelp2018", seed=0, scale=1.0):
    """ArrayInteraction of a named shape (optionally scaled down for quick tests)."""
    U, I, nnz = SHAPES[shape] if isinstance(shape, str) else shape
    U, I, nnz = max(8, int(U * scale)), max(8, int(I * scale)), max(16, int(nnz * scale))
    pu, pi = make_pairs(U, I, nnz, seed)
    return ArrayInteraction(pu, pi, U, I)
