#!/usr/bin/env python
"""Where does a sharded step differ from the single-GPU engine?  (torchrun, any world; diagnostic only)"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from selfrec_b200 import synth
    from selfrec_b200.engine import TrainEngine
    from selfrec_b200.shard_check import device_batches
    from selfrec_b200.sharded import ShardedEngine
    model = sys.argv[1] if len(sys.argv) > 1 else "XSimGCL"
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    data = synth.make_interaction("yelp2018", seed=0)
    d, B = 64, 2048
    kw = dict(eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1) if model != "LightGCN" else dict(l2_div=2048.0)
    U, I = data.user_num, data.item_num
    g = torch.Generator(device=dev).manual_seed(1234)
    iu = torch.empty((U, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    ii = torch.empty((I, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    pool = device_batches(data, B, 2, seed=5, dev=dev)
    sh = ShardedEngine(model, data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
    ref = TrainEngine(model, data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
    deg = torch.from_numpy(__import__("numpy").diff(data.norm_adj.tocsr().indptr)).to(dev)
    for k in range(2):
        ref.batch_dev.copy_(pool[k])
        ref.step_resident()
        sh.step(words_dev=pool[k])
        torch.cuda.synchronize()
        lo, hi = sh.user_lo, sh.user_hi
        ilo, ihi = int(sh.ib[sh.rank]), int(sh.ib[sh.rank + 1])
        for name, a, b, base in (("m_user", sh.mu, ref.m[lo:hi], lo), ("m_item", sh.mi[ilo:ihi], ref.m[U + ilo:U + ihi], U + ilo)):
            diff = (a - b).abs()
            rowmax = diff.max(1).values
            top = torch.topk(rowmax, 5)
            scale = float(b.abs().max())
            bad_rows = int((rowmax > 1e-4 * scale).sum())
            words = pool[k]
            bu = set(words[4:4 + B].tolist())
            bi = set(words[4 + B:4 + 2 * B].tolist()) | set(words[4 + 2 * B:4 + 3 * B].tolist())
            desc = []
            for r, v in zip(top.indices.tolist(), top.values.tolist()):
                gr = base + r
                inb = (gr in bu) if gr < U else ((gr - U) in bi)
                desc.append(f"row {gr} deg {int(deg[gr])} in_batch {inb} diff {v:.2e} ref_rowmax {float(b[r].abs().max()):.2e}")
            print(f"[rank {rank}] step {k} {name}: max|ref| {scale:.2e} rows_off {bad_rows}/{a.shape[0]} | " + " ; ".join(desc), flush=True)
        print(f"[rank {rank}] step {k} losses sh {sh.losses.tolist()} ref {ref.losses.tolist()}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
