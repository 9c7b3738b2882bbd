#!/usr/bin/env python
"""Who is right on the rows where the sharded step and the single-GPU engine disagree?  Both against the float64
oracle (XSimGCL, eps = 0 so that no noise tensor is needed), yelp2018 shape.  torchrun, world 2."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import oracle
    from selfrec_b200 import synth
    from selfrec_b200.engine import TrainEngine
    from selfrec_b200.shard_check import device_batches
    from selfrec_b200.sharded import ShardedEngine
    data = synth.make_interaction("yelp2018", seed=0)
    d, B, L = 64, 2048, 3
    U, I = data.user_num, data.item_num
    g = torch.Generator(device=dev).manual_seed(1234)
    iu = torch.empty((U, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    ii = torch.empty((I, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    pool = device_batches(data, B, 1, seed=5, dev=dev)
    kw = dict(eps=0.0, tau=0.2, cl_rate=0.2, layer_cl=int(os.environ.get("LCL", "0")))
    sh = ShardedEngine("XSimGCL", data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
    ref = TrainEngine("XSimGCL", data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
    ref.batch_dev.copy_(pool[0])
    ref.step_resident()
    sh.step(words_dev=pool[0])
    torch.cuda.synchronize()
    w = pool[0].cpu().numpy()
    b = int(w[0])
    u, i, j = w[4:4 + b], w[4 + B:4 + B + b], w[4 + 2 * B:4 + 2 * B + b]
    E0 = torch.cat([iu, ii]).cpu().numpy()
    out = oracle.train_step("XSimGCL", data.norm_adj.tocsr(), E0, U, u, i, j, n_layers=L, reg=1e-4, batch_size=B, eps=0.0, tau=0.2, cl_rate=0.2,
                            layer_cl=kw["layer_cl"], noise=np.zeros((1, L, U + I, d), np.float32))
    m_or = 0.1 * out["grad"]
    m_ref = ref.m.cpu().numpy()
    scale = np.abs(m_or).max()
    lo, hi = sh.user_lo, sh.user_hi
    ilo, ihi = int(sh.ib[sh.rank]), int(sh.ib[sh.rank + 1])
    m_sh_u, m_sh_i = sh.mu.cpu().numpy(), sh.mi[ilo:ihi].cpu().numpy()
    for name, got, want, base in (("single-GPU users", m_ref[lo:hi], m_or[lo:hi], lo), ("sharded users", m_sh_u, m_or[lo:hi], lo),
                                  ("single-GPU items", m_ref[U + ilo:U + ihi], m_or[U + ilo:U + ihi], ilo), ("sharded items", m_sh_i, m_or[U + ilo:U + ihi], ilo)):
        rowmax = np.abs(got - want).max(1)
        worst = np.argsort(-rowmax)[:4]
        print(f"[rank {rank}] {name}: max err / max|m| = {rowmax.max() / scale:.2e}, rows > 1e-4: {(rowmax > 1e-4 * scale).sum()} | worst " +
              ", ".join(f"id {base + r} err {rowmax[r]:.2e}" for r in worst), flush=True)
    print(f"[rank {rank}] losses single {ref.losses.tolist()} sharded {sh.losses.tolist()} oracle {[out['rec'], out['l2'], out['cl']]}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
