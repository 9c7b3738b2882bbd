#!/usr/bin/env python
"""Per-kernel timeline of the bipartite-sharded step on N ranks (torch.profiler / CUPTI -- ncu must not run a multi-rank
command).  Never a bench value: eager launches under a profiler.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_trace.py [--shape yelp2018|synthetic-2M|synthetic-10M]
        [--model XSimGCL|SimGCL|LightGCN] [--dim 64] [--steps 5]

Rank 0 prints, per kernel name, launches per step, mean duration and share of the step's kernel time, plus the wall
time of the step (CUDA events, eager and graph replay).
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="yelp2018")
    ap.add_argument("--model", default="XSimGCL")
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from selfrec_b200 import build, synth
    build.build()
    from selfrec_b200.shard_check import device_batches
    from selfrec_b200.sharded import ShardedEngine
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.shape.startswith("synthetic"):
        data = synth.make_device_interaction(args.shape, seed=0, alpha=1.1, device=dev)
    else:
        data = synth.make_interaction(args.shape, seed=0)
    kw = dict(eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1) if args.model == "XSimGCL" else (dict(eps=0.1, tau=0.2, cl_rate=0.5) if args.model == "SimGCL" else {})
    sh = ShardedEngine(args.model, data, args.dim, args.layers, 2048, 1e-3, 1e-4, device=dev, philox_seed=7, **kw)
    pool = device_batches(data, 2048, 8, seed=3, dev=dev)

    def step(k):
        sh.batch_dev.copy_(pool[k % 8], non_blocking=True)
        sh.step_resident()

    for k in range(3):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / args.steps
    from torch.profiler import ProfilerActivity, profile
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for k in range(args.steps):
            step(k)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != "CUDA":
            continue
        a = agg.setdefault(ev.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    sh.capture()
    for k in range(3):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for k in range(args.steps * 4):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / (args.steps * 4)
    sh.check_peers()
    # per-rank totals of the kernels that do the work (waits sit in spmm_hub_finish / spmm_csr / reduce_rows when the
    # syncs are folded into the kernels; run with SRB_SHARD_SYNC=barrier to see pure compute next to shard_barrier_kernel)
    mine = {"rank": rank, "users": sh.Ug, "nnz_Ru": int(sh.Ru.nnz), "nnz_Rt": int(sh.Rt.nnz),
            "kernel_us_per_step": sum(v[1] for v in agg.values()) / args.steps,
            "by_kernel_us": {k.split("(")[0].replace("void srb::", "").replace("srb::", ""): round(v[1] / args.steps, 1) for k, v in agg.items()}}
    every = [mine]
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, mine)
    if rank == 0:
        for m in every:
            top = sorted(m["by_kernel_us"].items(), key=lambda kv: -kv[1])[:7]
            print(f"rank {m['rank']}: users {m['users']} nnz Ru {m['nnz_Ru']} Rt {m['nnz_Rt']} kernels {m['kernel_us_per_step']:.0f} us/step  " + "  ".join(f"{k}={v}" for k, v in top))
    if rank == 0:
        tot = sum(v[1] for v in agg.values())
        print(json.dumps({"shape": args.shape, "model": args.model, "world": world, "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
                          "kernel_us_per_step": tot / args.steps, "multicast": sh.use_multicast}))
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{us / args.steps:10.1f} us/step {100 * us / tot:5.1f}%  x{n / args.steps:5.1f}  {us / n:9.1f} us  {name}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
