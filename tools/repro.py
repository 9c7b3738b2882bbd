import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selfrec_b200 import ops, synth
from selfrec_b200.engine import TrainEngine
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
rng = np.random.default_rng(0)
for n in (100, 1900, 2048):
    v1 = torch.from_numpy((rng.standard_normal((n, 64)) * 0.1).astype(np.float32)).cuda().requires_grad_(True)
    v2 = torch.from_numpy((rng.standard_normal((n, 64)) * 0.1).astype(np.float32)).cuda().requires_grad_(True)
    l = ops.InfoNCE(v1, v2, 0.2); sync(f"infonce fwd n={n} loss={l.item():.5f}")
    l.backward(); sync(f"infonce bwd n={n}")
data = synth.make_interaction((3000, 4000, 60000), seed=3)
for model, kw in (("LightGCN", dict(l2_div=512.0)), ("XSimGCL", dict(eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1)), ("SimGCL", dict(eps=0.1, tau=0.2, cl_rate=0.5)), ("SGL", dict(tau=0.2, cl_rate=0.1))):
    eng = TrainEngine(model, data, 64, 3, 512, 1e-3, 1e-4, **kw)
    if model == "SGL":
        eng.set_view_graphs(data.norm_adj, data.norm_adj)
    import random; random.seed(1)
    w = next(eng.batches()).copy()
    eng.step(w); sync(f"{model} step losses={eng.losses.tolist()}")
    eng.forward_clean(); sync(f"{model} clean forward")
