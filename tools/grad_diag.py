"""Diagnostic: where does the fused XSimGCL step's E0 gradient differ from the float64 oracle?"""
import os, sys, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle
from selfrec_b200 import synth
from selfrec_b200.engine import TrainEngine
yelp = synth.make_interaction("yelp2018", seed=0)
torch.manual_seed(0)
eng = TrainEngine("XSimGCL", yelp, 64, 3, 2048, 1e-3, 1e-4, eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1)
U, N = eng.U, eng.N
rng = np.random.default_rng(1)
noise = rng.random((1, 3, N, 64), dtype=np.float32)
eng.set_noise_tensor(torch.from_numpy(noise).cuda())
E0 = eng.params.cpu().numpy().copy()
b = 2048
u = yelp.pair_users[:b].copy(); i = yelp.pair_items[:b].copy()
rp, ri = yelp.rated_csr()
j = np.array([next(x for x in rng.integers(0, eng.I, 64) if x not in ri[rp[uu]:rp[uu + 1]]) for uu in u], dtype=np.int32)
w = np.zeros(eng.words, dtype=np.int32)
uq, iq = np.unique(u), np.unique(i)
w[0], w[1], w[2] = b, len(uq), len(iq)
for s, arr in enumerate((u, i, j, uq, iq)):
    w[4 + s * b:4 + s * b + len(arr)] = arr
eng.step(w); torch.cuda.synchronize()
g_gpu = eng.m.cpu().numpy().astype(np.float64) * 10.0   # m = 0.1 g after step 1
out = oracle.train_step("XSimGCL", yelp.norm_adj.tocsr(), E0, U, u, i, j, n_layers=3, reg=1e-4, batch_size=2048, eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1, noise=noise)
g = out["grad"]
d = np.abs(g_gpu - g)
print("rms(g)", np.sqrt((g**2).mean()), "max|g|", np.abs(g).max(), "max abs diff", d.max(), "rms diff", np.sqrt((d**2).mean()))
rel = d / (np.abs(g) + 1e-12)
for thr in (1e-9, 1e-8, 1e-7, 1e-6, 1e-5):
    sel = np.abs(g) > thr
    print(f"|g|>{thr:g}: n={sel.sum()} max rel {rel[sel].max():.3e} median rel {np.median(rel[sel]):.3e}")
rows = np.argsort(-d.max(1))[:10]
deg = np.diff(yelp.norm_adj.tocsr().indptr)
inb = np.zeros(N, bool); inb[u] = True; inb[U + i] = True; inb[U + j] = True
for r in rows:
    c = d[r].argmax()
    print("row", r, "user" if r < U else "item", "deg", deg[r], "in_batch", inb[r], "g", g[r, c], "gpu", g_gpu[r, c], "row |g| max", np.abs(g[r]).max())
# forward check: final embeddings
fin = out["final"]
ue, ie = eng.forward_clean()
