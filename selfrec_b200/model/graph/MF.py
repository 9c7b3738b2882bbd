"""MF + BPR (reference model/graph/MF.py) on the fused engine: gather + BPR + L2 + dense Adam."""
from ._common import FusedGraphModel


class MF(FusedGraphModel):
    MODEL = "MF"
    EVAL_EVERY = 5  # MF.py:30-31

    def __init__(self, conf, training_set, test_set):
        super(MF, self).__init__(conf, training_set, test_set)
        # l2_reg_loss(reg, u, p, n) / batch_size  (MF.py:21)
        self._make_engine(0, l2_div=float(self.batch_size))

    def _log_line(self, epoch, n, losses):
        print("training:", epoch + 1, "batch", n, "batch_loss:", losses[3])


from ._common import OpLevelEncoder  # noqa: E402


class Matrix_Factorization(OpLevelEncoder):
    """`from model.graph.MF import Matrix_Factorization` (DirectAU.py:6) keeps resolving after install()."""

    def __init__(self, data, emb_size):
        super().__init__(data, emb_size, 0)
