"""Mirror of data/graph.py:10-24 (Graph.normalize_graph_mat)."""
import numpy as np
import scipy.sparse as sp


def _scale_rows_cols(adj, left, right=None):
    """diag(left) @ adj [@ diag(right)] with the reference's rounding: every output entry is
    the fp32 product (left[i] * a_ij) [* right[j]] -- no sums -- so this is bit-identical to
    scipy's diags().dot() chain (graph.py:16-18) at a fraction of the cost."""
    csr = sp.csr_matrix(adj, dtype=np.float32)
    csr.sort_indices()
    rows = np.repeat(np.arange(csr.shape[0]), np.diff(csr.indptr))
    data = (left.astype(np.float32)[rows] * csr.data).astype(np.float32)
    if right is not None:
        data = (data * right.astype(np.float32)[csr.indices]).astype(np.float32)
    return sp.csr_matrix((data, csr.indices.copy(), csr.indptr.copy()), shape=csr.shape)


class Graph(object):
    def __init__(self):
        pass

    @staticmethod
    def normalize_graph_mat(adj_mat):
        """Square: D^-1/2 A D^-1/2 ; rectangular: D^-1 A.  fp32, inf -> 0 (graph.py:13-23)."""
        shape = adj_mat.get_shape()
        rowsum = np.array(adj_mat.sum(1)).flatten()
        with np.errstate(divide="ignore"):
            if shape[0] == shape[1]:
                d_inv = np.power(rowsum, -0.5)
                d_inv[np.isinf(d_inv)] = 0.0
                return _scale_rows_cols(adj_mat, d_inv, d_inv)
            d_inv = np.power(rowsum, -1)
            d_inv[np.isinf(d_inv)] = 0.0
            return _scale_rows_cols(adj_mat, d_inv)

    def convert_to_laplacian_mat(self, adj_mat):
        pass
