"""XSimGCL (reference model/graph/XSimGCL.py) on the fused engine -- the north-star step.

One perturbed forward yields both the recommendation embedding (mean of layers 1..L) and the
contrastive view (layer l*), XSimGCL.py:83-101; losses per XSimGCL.py:31-33."""
from ._common import FusedGraphModel


class XSimGCL(FusedGraphModel):
    MODEL = "XSimGCL"

    def __init__(self, conf, training_set, test_set):
        super(XSimGCL, self).__init__(conf, training_set, test_set)
        config = self.config["XSimGCL"]
        self.cl_rate = float(config["lambda"])
        self.eps = float(config["eps"])
        self.temp = float(config["tau"])
        self.n_layers = int(config["n_layer"])
        self.layer_cl = int(config["l_star"])
        self._make_engine(self.n_layers, eps=self.eps, tau=self.temp, cl_rate=self.cl_rate, layer_cl=self.layer_cl)
