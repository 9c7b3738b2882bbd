"""Launched by torchrun (one rank per GPU): the row-sharded XSimGCL / LightGCN step must produce
the same parameters as the single-GPU fused engine on the same batches and noise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from selfrec_b200 import synth
    from selfrec_b200.engine import TrainEngine
    from selfrec_b200.sharded import ShardedXSimGCL
    data = synth.make_interaction((3000, 4000, 60000), seed=3)
    N, d, L, B = data.user_num + data.item_num, 64, 3, 512
    rng = np.random.default_rng(0)
    iu = (rng.standard_normal((data.user_num, d)) * 0.05).astype(np.float32)
    ii = (rng.standard_normal((data.item_num, d)) * 0.05).astype(np.float32)
    ok = True
    for model, kw in (("XSimGCL", dict(eps=0.2, tau=0.2, cl_rate=0.2, layer_cl=1)), ("LightGCN", dict(l2_div=float(B)))):
        ref = TrainEngine(model, data, d, L, B, 1e-3, 1e-4, init_user=torch.from_numpy(iu), init_item=torch.from_numpy(ii), **kw)
        sh = ShardedXSimGCL(model, data, d, L, B, 1e-3, 1e-4, init_user=iu, init_item=ii, **kw)
        import random
        random.seed(11)
        batches = [w.copy() for _, w in zip(range(3), ref.batches())]
        for k, w in enumerate(batches):
            if model == "XSimGCL":
                nz = torch.from_numpy(np.random.default_rng(100 + k).random((1, L, N, d), dtype=np.float32)).cuda()
                ref.set_noise_tensor(nz)
                sh.set_noise_tensor(nz)
            ref.step(w)
            sh.step(w)
            torch.cuda.synchronize()
            a, b = ref.params.cpu().numpy(), sh.params.cpu().numpy()
            la, lb = ref.losses.cpu().numpy(), sh.losses.cpu().numpy()
            good = np.allclose(a, b, rtol=1e-4, atol=1e-6) and np.allclose(la, lb, rtol=1e-4, atol=1e-7)
            if not good:
                print(f"rank {rank} {model} step {k}: max param diff {np.abs(a - b).max():.3e} losses {la} vs {lb}", flush=True)
            ok = ok and good
        ue, ie = sh.forward_clean()
        re_u, re_i = ref.forward_clean()
        ok = ok and np.allclose(ue.cpu().numpy(), re_u.cpu().numpy(), rtol=1e-4, atol=1e-6)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_CHECK", "PASS" if t.item() == 1.0 else "FAIL", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
