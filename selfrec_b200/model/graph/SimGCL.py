"""SimGCL (reference model/graph/SimGCL.py) on the fused engine.

Clean encoder for BPR, two perturbed encoders for the InfoNCE views (SimGCL.py:43-50) with
the hard-coded temperature 0.2 (SimGCL.py:48-49); mean over layers 1..L (ego excluded)."""
from ._common import FusedGraphModel


class SimGCL(FusedGraphModel):
    MODEL = "SimGCL"

    def __init__(self, conf, training_set, test_set):
        super(SimGCL, self).__init__(conf, training_set, test_set)
        args = self.config["SimGCL"]
        self.cl_rate = float(args["lambda"])
        self.eps = float(args["eps"])
        self.n_layers = int(args["n_layer"])
        self._make_engine(self.n_layers, eps=self.eps, tau=0.2, cl_rate=self.cl_rate)
