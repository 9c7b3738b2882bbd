"""SGL (reference model/graph/SGL.py:1-126) on the fused engine.

Per epoch two augmented graphs (SGL.py:27-29, via GraphAugmentor + convert_to_laplacian_mat),
three encoders per step (clean + 2 views), one InfoNCE over cat(users, items) (SGL.py:120-125).
The reference's `aug_type==0 or 1` test is always true (SGL.py:81), so every aug_type yields
a single graph per view; that behaviour is kept."""
from ...data.augmentor import GraphAugmentor
from ._common import FusedGraphModel


class SGL(FusedGraphModel):
    MODEL = "SGL"
    EVAL_FROM = 5  # SGL.py:45-46

    def __init__(self, conf, training_set, test_set):
        super(SGL, self).__init__(conf, training_set, test_set)
        args = self.config["SGL"]
        self.cl_rate = float(args["lambda"])
        self.aug_type = int(args["aug_type"])
        self.drop_rate = float(args["drop_rate"])
        self.n_layers = int(args["n_layer"])
        self.temp = float(args["temp"])
        self._make_engine(self.n_layers, tau=self.temp, cl_rate=self.cl_rate)

    def random_graph_augment(self):
        if self.aug_type == 0:
            dropped = GraphAugmentor.node_dropout(self.data.interaction_mat, self.drop_rate)
        else:
            dropped = GraphAugmentor.edge_dropout(self.data.interaction_mat, self.drop_rate)
        return self.data.convert_to_laplacian_mat(dropped)

    def _epoch_prologue(self, epoch):
        self.engine.set_view_graphs(self.random_graph_augment(), self.random_graph_augment())
