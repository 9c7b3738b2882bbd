"""LightGCN (reference model/graph/LightGCN.py) on the fused engine.

Mean over L+1 layers including the ego layer (LightGCN.py:74-75); L2 on the RAW parameters
gathered by index, divided by the configured batch size (LightGCN.py:25)."""
from ._common import FusedGraphModel


class LightGCN(FusedGraphModel):
    MODEL = "LightGCN"
    EVAL_EVERY = 5  # LightGCN.py:34-35

    def __init__(self, conf, training_set, test_set):
        super(LightGCN, self).__init__(conf, training_set, test_set)
        args = self.config["LightGCN"]
        self.n_layers = int(args["n_layer"])
        self._make_engine(self.n_layers, l2_div=float(self.batch_size))

    def _log_line(self, epoch, n, losses):
        print("training:", epoch + 1, "batch", n, "batch_loss:", losses[3])


from ._common import OpLevelEncoder  # noqa: E402


class LGCN_Encoder(OpLevelEncoder):
    """`from model.graph.LightGCN import LGCN_Encoder` (DirectAU.py:7, SelfCF.py:6) keeps resolving after install()."""

    def __init__(self, data, emb_size, n_layers):
        super().__init__(data, emb_size, n_layers)
