#!/usr/bin/env python
"""Probe of the config-5 path on one GPU (never a bench value by itself; bench.py carries the record):
graph generation + device assembly times, SpMM launch time / algorithmic GB/s, one SimGCL step.

    python tools/config5_probe.py [shape] [--steps K] [--model SimGCL] [--d 128] [--alpha 1.1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fast_batches(data, B, n, seed=0):
    """Batch words from uniformly drawn pairs (device), negatives uniform (no rejection: timing only)."""
    import torch
    from selfrec_b200 import _lib
    pu, pi = data.pairs_dev
    g = torch.Generator(device=pu.device).manual_seed(seed)
    H = _lib.BATCH_HEADER
    out = torch.zeros((n, H + 5 * B), dtype=torch.int32, device=pu.device)
    for k in range(n):
        sel = torch.randint(0, pu.numel(), (B,), generator=g, device=pu.device)
        u, i = pu[sel], pi[sel]
        j = torch.randint(0, data.item_num, (B,), generator=g, device=pu.device, dtype=torch.int32)
        uq, iq = torch.unique(u), torch.unique(i)
        w = out[k]
        w[0], w[1], w[2] = B, uq.numel(), iq.numel()
        w[H:H + B], w[H + B:H + 2 * B], w[H + 2 * B:H + 3 * B] = u, i, j
        w[H + 3 * B:H + 3 * B + uq.numel()] = uq
        w[H + 4 * B:H + 4 * B + iq.numel()] = iq
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="?", default="synthetic-2M")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--model", default="SimGCL")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=1.1)
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    import torch
    from selfrec_b200 import build, ops, synth
    build.build()
    from selfrec_b200.engine import TrainEngine
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rec = {"shape": args.shape, "model": args.model, "d": args.dim, "alpha": args.alpha}
    t0 = time.perf_counter()
    shape = synth.SHAPES[args.shape] if args.shape in synth.SHAPES else tuple(int(x) for x in args.shape.split("x"))
    pu, pi = synth.make_pairs_device(*shape, seed=0, alpha=args.alpha, device=dev)
    torch.cuda.synchronize()
    rec["gen_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    data = synth.DeviceInteraction(pu, pi, shape[0], shape[1])
    torch.cuda.synchronize()
    rec["build_s"] = time.perf_counter() - t0
    adj = data.norm_adj
    deg = (adj.rowptr[1:] - adj.rowptr[:-1])
    rec.update(N=adj.shape[0], nnzA=adj.nnz, n_huge=adj.n_huge, n_work=adj.n_work, n_vlong=adj.n_vlong, n_long=adj.n_long,
               max_deg_user=int(deg[:shape[0]].max()), max_deg_item=int(deg[shape[0]:].max()),
               nnz_in_split_rows=int(deg[deg >= 4096].sum()))
    torch.cuda.empty_cache()
    rec["mem_after_build_gb"] = torch.cuda.memory_allocated() / 1e9
    # ---- SpMM alone ----
    N, d = adj.shape[0], args.dim
    x = torch.randn(N, d, device=dev)
    y = torch.empty_like(x)
    for _ in range(2):
        ops._spmm_raw(adj, x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 6
    e0.record()
    for _ in range(R // 2):
        ops._spmm_raw(adj, x, y)
        ops._spmm_raw(adj, y, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    alg = 8 * adj.nnz + 4 * (N + 1) + 8 * N * d
    rec["spmm"] = {"ms": ms, "alg_bytes": alg, "alg_gbs": alg / ms / 1e6, "gather_bytes": 4 * adj.nnz * d,
                   "gather_gbs": 4 * adj.nnz * d / ms / 1e6}
    del x, y
    torch.cuda.empty_cache()
    # ---- one training step ----
    torch.manual_seed(0)
    eng = TrainEngine(args.model, data, d, args.layers, 2048, 1e-3, 1e-4, eps=0.1, tau=0.2, cl_rate=0.5, layer_cl=1, device=dev)
    rec["mem_engine_gb"] = torch.cuda.memory_allocated() / 1e9
    pool = fast_batches(data, 2048, 8)
    graph = None if args.eager else eng.capture()
    def step(k):
        eng.batch_dev.copy_(pool[k % pool.shape[0]], non_blocking=True)
        if graph is not None:
            graph.replay()
        else:
            eng.step_resident()
    for k in range(2):
        step(k)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    rec["step_ms"] = e0.elapsed_time(e1) / args.steps
    rec["steps_per_s"] = 1e3 / rec["step_ms"]
    rec["loss"] = eng.losses.cpu().tolist()
    rec["mem_peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
