"""Diagnostic: tensor-core InfoNCE vs the float64 oracle, and its launch time (CUDA events)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle
from selfrec_b200 import ops

for n, tau in ((1900, 0.2), (2048, 0.2), (333, 0.15), (64, 0.2), (129, 0.5)):
    d = 64
    rng = np.random.default_rng(n)
    v1 = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
    v2 = (v1 + 0.05 * rng.standard_normal((n, d))).astype(np.float32)
    t1, t2 = torch.from_numpy(v1).cuda(), torch.from_numpy(v2).cuda()
    idx = torch.arange(n, device="cuda", dtype=torch.int32)
    prob = [dict(table1=t1, table2=t2, idx=idx, n=n, weight=1.0)]
    losses, outs = ops.infonce_raw(prob, d, tau)
    torch.cuda.synchronize()
    ref, g1, g2 = oracle.infonce(v1, v2, tau)
    e1 = np.abs(outs[0][0].cpu().numpy() - g1).max() / np.abs(g1).max()
    e2 = np.abs(outs[0][1].cpu().numpy() - g2).max() / np.abs(g2).max()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ops.infonce_raw(prob * 2, d, tau)
    ev[0].record()
    for _ in range(20):
        ops.infonce_raw(prob * 2, d, tau)
    ev[1].record(); torch.cuda.synchronize()
    print(f"n={n} tau={tau} loss {losses[0].item():.7f} ref {ref:.7f}  max err / max|g|: g1 {e1:.2e} g2 {e2:.2e}  "
          f"2 problems: {ev[0].elapsed_time(ev[1]) / 20 * 1000:.1f} us/call (host-inclusive)", flush=True)
