"""find_k_largest with the signature of util/algorithm.py:144-156, on the CUDA top-k kernel."""
import numpy as np
import torch

from .. import ops


def find_k_largest(K, candidates):
    """ids, scores of the K largest entries of one score vector (score-descending).

    Same selection rule as the reference's numba heap (strict > against the current K-th,
    earliest id kept on ties at the threshold); runs srb_topk_rows on the GPU."""
    sc = torch.as_tensor(np.asarray(candidates, dtype=np.float32)).reshape(1, -1).cuda()
    ids, scores = ops.topk_rows(sc, int(K))
    return ids[0].tolist(), [float(s) for s in scores[0].tolist()]
