"""BASELINE configs[4]: Llama-3-70B-shaped solver sweep - n = 8192 (q/k/v/o/gate/up input) and n = 28672 (down_proj
input) Hessians through the whole hot path: SYRK over 65536 synthetic tokens, process_hessian_alt (eigh + pivot order
+ R), gptq_fwrd of one Linear of the group (8192 x 8192 / 8192 x 28672), 4-bit sym g128, eps 1e-4.  Under torchrun
every rank runs its own replica (the single n = 28672 solve does not shard: DESIGN.md 5); rank 0 prints the max.
Usage: python scripts/llama70b_sweep.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gptq_svd_b200 as G
from bench import make_x


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ns = [int(a) for a in sys.argv[1:]] or [8192, 28672]
    for n in ns:
        m = 8192
        X = make_x(torch, 65536, n, 100 + rank, -1.0)

        def hess():
            acc = G.HessianAccumulator(n, X.device)
            acc.add_batch(X.view(-1, 2048, n))
            return acc.get_hessian()
        H, _ = timed(hess)
        H, t_syrk = timed(hess)
        del X
        f, t_solve = timed(lambda: G.spectral_solve(H, 1e-4, "energy"))
        if n <= 8192:
            f, t_solve = timed(lambda: G.spectral_solve(H, 1e-4, "energy"))       # second call (warm)
        del H
        W = (torch.randn(m, n, device="cuda") * 0.02).half()
        _, t_loop = timed(lambda: G.gptq_fwrd(W, f.R, G.Quantizer(4, 128, True), f.perm, block_size=1024, R_x=f.R_x))
        _, t_loop = timed(lambda: G.gptq_fwrd(W, f.R, G.Quantizer(4, 128, True), f.perm, block_size=1024, R_x=f.R_x))
        rec = torch.tensor([t_syrk, t_solve, t_loop], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(rec, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"n": n, "k": int(f.k), "replicas": world, "syrk_ms_65536_tokens": round(float(rec[0]), 2),
                              "solver_ms": round(float(rec[1]), 1), "gptq_fwrd_ms_8192_rows": round(float(rec[2]), 2),
                              "solver_fp64_gflops_7.4n3": round(7.4 * n ** 3 / float(rec[1]) / 1e6, 1)}), flush=True)
        del f, W
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
