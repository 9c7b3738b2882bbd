"""Mirror of data/augmentor.py:6-40 (GraphAugmentor), used by SGL's per-epoch views.

Consumes Python's global `random` stream exactly like the reference (random.sample over
range(n)), so a seeded run drops the same nodes/edges; the draw itself runs natively."""
import ctypes as C
import random
from math import ceil as _ceil, log as _log

import numpy as np
import scipy.sparse as sp

from .. import _lib


def sample_range(n, k):
    """random.sample(range(n), k) as an int64 array, drawn natively (srb_random_sample_range) from Python's
    global `random` state, which is advanced exactly as random.sample would advance it."""
    if not 0 <= k <= n:
        raise ValueError("Sample larger than population or is negative")
    if n >= 1 << 32:
        return np.asarray(random.sample(range(n), k), dtype=np.int64)
    setsize = 21  # Lib/random.py sample(): which of CPython's two strategies applies
    if k > 5:
        setsize += 4 ** _ceil(_log(k * 3, 4))
    st = random.getstate()
    mt = np.array(st[1], dtype=np.uint32)
    out = np.empty(k, dtype=np.int64)
    lib = _lib.load()
    _lib.check(lib.srb_random_sample_range(mt.ctypes.data_as(C.c_void_p), n, k, int(n <= setsize), out.ctypes.data_as(C.c_void_p)),
               "srb_random_sample_range")
    random.setstate((st[0], tuple(int(x) for x in mt), st[2]))
    return out


class GraphAugmentor(object):
    @staticmethod
    def node_dropout(sp_adj, drop_rate):
        n_u, n_i = sp_adj.get_shape()
        rows, cols = sp_adj.nonzero()
        keep_u = np.ones(n_u, dtype=np.float32)
        keep_i = np.ones(n_i, dtype=np.float32)
        keep_u[sample_range(n_u, int(n_u * drop_rate))] = 0.0
        keep_i[sample_range(n_i, int(n_i * drop_rate))] = 0.0
        ones = sp.csr_matrix((np.ones_like(rows, dtype=np.float32), (rows, cols)), shape=(n_u, n_i))
        return sp.diags(keep_u).dot(ones).dot(sp.diags(keep_i))

    @staticmethod
    def edge_dropout(sp_adj, drop_rate):
        shape = sp_adj.get_shape()
        nnz = sp_adj.count_nonzero()
        rows, cols = sp_adj.nonzero()
        keep = sample_range(nnz, int(nnz * (1 - drop_rate)))
        ku, ki = np.array(rows)[keep], np.array(cols)[keep]
        return sp.csr_matrix((np.ones_like(ku, dtype=np.float32), (ku, ki)), shape=shape)
