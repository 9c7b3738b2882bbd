"""Mirror of data/augmentor.py:6-40 (GraphAugmentor), used by SGL's per-epoch views.

Consumes Python's global `random` stream exactly like the reference (random.sample over
range(n)), so a seeded run drops the same nodes/edges."""
import random

import numpy as np
import scipy.sparse as sp


class GraphAugmentor(object):
    @staticmethod
    def node_dropout(sp_adj, drop_rate):
        n_u, n_i = sp_adj.get_shape()
        rows, cols = sp_adj.nonzero()
        keep_u = np.ones(n_u, dtype=np.float32)
        keep_i = np.ones(n_i, dtype=np.float32)
        keep_u[random.sample(range(n_u), int(n_u * drop_rate))] = 0.0
        keep_i[random.sample(range(n_i), int(n_i * drop_rate))] = 0.0
        ones = sp.csr_matrix((np.ones_like(rows, dtype=np.float32), (rows, cols)), shape=(n_u, n_i))
        return sp.diags(keep_u).dot(ones).dot(sp.diags(keep_i))

    @staticmethod
    def edge_dropout(sp_adj, drop_rate):
        shape = sp_adj.get_shape()
        nnz = sp_adj.count_nonzero()
        rows, cols = sp_adj.nonzero()
        keep = random.sample(range(nnz), int(nnz * (1 - drop_rate)))
        ku, ki = np.array(rows)[keep], np.array(cols)[keep]
        return sp.csr_matrix((np.ones_like(ku, dtype=np.float32), (ku, ki)), shape=shape)
