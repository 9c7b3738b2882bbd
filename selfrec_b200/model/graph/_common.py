"""Shared scaffolding of the fused in-scope models (MF, LightGCN, SimGCL, XSimGCL, SGL).

Each model keeps the reference class name, constructor signature, YAML keys, attributes
(`model.embedding_dict`, `user_emb`, `item_emb`, `best_user_emb`, ...) and train()/save()/
predict() contract; the batch loop body is one TrainEngine.step()."""
import torch
import torch.nn as nn

from ...base.graph_recommender import GraphRecommender
from ...engine import TrainEngine


class _EncoderView(nn.Module):
    """Exposes the engine's single [U+I, d] table as the reference's embedding_dict."""

    def __init__(self, engine):
        super().__init__()
        self.engine = engine
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(engine.user_emb, requires_grad=False),
            "item_emb": nn.Parameter(engine.item_emb, requires_grad=False),
        })

    def forward(self, *args, **kwargs):
        with torch.no_grad():
            return self.engine.forward_clean()

    def cuda(self, device=None):
        return self


class FusedGraphModel(GraphRecommender):
    MODEL = None
    EVAL_EVERY = 1      # fast_evaluation cadence (epochs)
    EVAL_FROM = 0

    def _engine_kwargs(self):
        return {}

    def _make_engine(self, n_layers, **kw):
        self.engine = TrainEngine(self.MODEL, self.data, self.emb_size, n_layers, self.batch_size, self.lRate, self.reg, **kw)
        self.model = _EncoderView(self.engine)

    def _epoch_prologue(self, epoch):
        pass

    def _log_line(self, epoch, n, losses):
        print("training:", epoch + 1, "batch", n, "rec_loss:", losses[0], "cl_loss", losses[2])

    def train(self):
        eng = self.engine
        for epoch in range(self.maxEpoch):
            self._epoch_prologue(epoch)
            if eng.graph is None:
                eng.capture()  # one CUDA graph launch per batch from here on
            for n, words in enumerate(eng.batches()):
                if n % 100 == 0 and n > 0:
                    self._log_line(epoch, n, eng.step(words, fetch_loss=True).get().tolist())
                else:
                    eng.step(words)
            with torch.no_grad():
                self.user_emb, self.item_emb = eng.forward_clean()
            if epoch >= self.EVAL_FROM and epoch % self.EVAL_EVERY == 0:
                self.fast_evaluation(epoch)
        self.user_emb, self.item_emb = self.best_user_emb, self.best_item_emb

    def save(self):
        with torch.no_grad():
            ue, ie = self.engine.forward_clean()
            self.best_user_emb, self.best_item_emb = ue.clone(), ie.clone()

    def predict(self, u):
        u = self.data.get_user_id(u)
        # one user's full-catalog scores (reference predict(), e.g. XSimGCL.py:57-60); test()
        # never calls this -- it uses the fused scoring + top-k kernel.
        from ... import ops
        return ops.score_rows(self.user_emb, self.item_emb, [u])[0].cpu().numpy()


class OpLevelEncoder(nn.Module):
    """Autograd-visible encoder on the drop-in ops, for the reference models that import another model's encoder
    class (`from model.graph.LightGCN import LGCN_Encoder`: DirectAU.py:7, SelfCF.py:6; `from model.graph.MF import
    Matrix_Factorization`: DirectAU.py:6).  Same attributes the callers touch (data, latent_size, layers, norm_adj,
    embedding_dict, sparse_norm_adj) and the same forward() contract: (user embeddings, item embeddings), the mean of
    the ego layer and `layers` propagated ones (LightGCN.py:68-78); zero layers is matrix factorisation (MF.py:52-53).
    Propagation is torch.sparse.mm on the SparseAdj handle, i.e. the CUDA SpMM, differentiable w.r.t. the table."""

    def __init__(self, data, emb_size, n_layers=0):
        super().__init__()
        from ...base.torch_interface import TorchGraphInterface
        self.data = data
        self.latent_size = emb_size
        self.layers = n_layers
        init = nn.init.xavier_uniform_  # users first, then items: the reference's draw order (LightGCN.py:60-66)
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(init(torch.empty(data.user_num, emb_size))),
            "item_emb": nn.Parameter(init(torch.empty(data.item_num, emb_size))),
        })
        if n_layers > 0:
            self.norm_adj = data.norm_adj
            self.sparse_norm_adj = TorchGraphInterface.convert_sparse_mat_to_tensor(self.norm_adj).cuda()

    def forward(self):
        pd = self.embedding_dict
        if self.layers == 0:
            return pd["user_emb"], pd["item_emb"]
        x = torch.cat([pd["user_emb"], pd["item_emb"]], 0)
        total = x
        for _ in range(self.layers):
            x = torch.sparse.mm(self.sparse_norm_adj, x)
            total = total + x
        out = total / float(self.layers + 1)
        n_u = self.data.user_num
        return out[:n_u], out[n_u:]
