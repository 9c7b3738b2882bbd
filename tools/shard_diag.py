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
    data = synth.make_interaction("yelp2018", seed=0)
    d, B = 64, 2048
    U, I = data.user_num, data.item_num
    g = torch.Generator(device=dev).manual_seed(1234)
    iu = torch.empty((U, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    ii = torch.empty((I, d), device=dev).uniform_(-0.1, 0.1, generator=g)
    pool = device_batches(data, B, 2, seed=5, dev=dev)
    cases = sys.argv[1:] or ["XSimGCL:3:1", "XSimGCL:3:0", "XSimGCL:2:1", "XSimGCL:1:1", "LightGCN:3:0", "LightGCN:2:0", "SimGCL:3:0", "SimGCL:2:0"]
    for case in cases:
        model, L, lcl = case.split(":")
        L, lcl = int(L), int(lcl)
        kw = dict(eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=lcl) if model != "LightGCN" else dict(l2_div=2048.0)
        sh = ShardedEngine(model, data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
        ref = TrainEngine(model, data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, philox_seed=7, device=dev, **kw)
        for k in range(2):
            ref.batch_dev.copy_(pool[k])
            ref.step_resident()
            sh.step(words_dev=pool[k])
            torch.cuda.synchronize()
            lo, hi = sh.user_lo, sh.user_hi
            ilo, ihi = int(sh.ib[sh.rank]), int(sh.ib[sh.rank + 1])
            msg = []
            for name, a, b in (("m_user", sh.mu, ref.m[lo:hi]), ("m_item", sh.mi[ilo:ihi], ref.m[U + ilo:U + ihi])):
                rowmax = (a - b).abs().max(1).values
                scale = float(b.abs().max())
                msg.append(f"{name} rel {float(rowmax.max()) / scale:.1e} rows_off {int((rowmax > 1e-4 * scale).sum())}/{a.shape[0]}")
            if os.environ.get("DIAG_ROWS"):
                w = pool[k].tolist()
                b, nu, ni = w[0], w[1], w[2]
                H = 4
                uq_u, uq_i = w[H + 3 * B:H + 3 * B + nu], w[H + 4 * B:H + 4 * B + ni]
                bu, bi, bj = w[H:H + b], w[H + B:H + B + b], w[H + 2 * B:H + 2 * B + b]
                for name, a, bb, base, uq, secs in (("m_user", sh.mu, ref.m[lo:hi], lo, uq_u, (bu,)), ("m_item", sh.mi[ilo:ihi], ref.m[U + ilo:U + ihi], ilo, uq_i, (bi, bj))):
                    rowmax = (a - bb).abs().max(1).values
                    scale = float(bb.abs().max())
                    off = torch.nonzero(rowmax > 1e-4 * scale).flatten().tolist()[:12]
                    for r in off:
                        gid = base + r
                        pos = uq.index(gid) if gid in uq else -1
                        cnts = [sec.count(gid) for sec in secs]
                        print(f"   [rank {rank}] {name} id {gid} pos_in_unique {pos}/{len(uq)} count_in_batch {cnts} diff {float(rowmax[r]):.2e} refmax {float(bb[r].abs().max()):.2e}", flush=True)
            lr = float(((sh.losses - ref.losses).abs() / ref.losses.abs().clamp_min(1e-12)).max())
            print(f"[rank {rank}/{world}] {case} step {k}: " + " | ".join(msg) + f" | loss rel {lr:.1e}", flush=True)
        del sh, ref
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
