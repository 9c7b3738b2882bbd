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
