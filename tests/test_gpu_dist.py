"""GPU (NCCL, world_size 2): the multi-GPU data plane of north_star on hardware - calibration tokens sharded
across ranks, ONE all-reduce of each fp64 Hessian, the solves of a block on different GPUs, factors handed
point-to-point to the GPUs that own sibling Linears (gptq_svd_b200/dist.py).  Same inputs at 1 and 2 ranks must
give bit-identical k / perm / codes: the SYRK adds each batch's total to H in fp64, so the
two-rank sum c1 + c2 is the one-rank sum (0 + c1) + c2.  Skipped on a box with one GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

GROUPS = [(512, [512, 128, 128]), (512, [512]), (512, [1536, 1536]), (1536, [512])]
TOKENS, SEQ = 8192, 2048


def _inputs(dev):
    Xs, Ws = [], []
    for gi, (n, outs) in enumerate(GROUPS):
        g = torch.Generator(device=dev).manual_seed(100 + gi)
        A = torch.randn(n, n, device=dev, generator=g) * torch.logspace(0, -2, n, device=dev)[None, :]
        Xs.append((torch.randn(TOKENS, n, device=dev, generator=g) @ A.T / n ** 0.5).half())
        Ws.append([(torch.randn(m, n, device=dev, generator=g) * 0.02).half() for m in outs])
    return Xs, Ws


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gptq_svd_b200 import dist as D
        Xs, Ws = _inputs(dev)
        shards = []
        for gi in range(len(GROUPS)):
            b, e = D.shard_range(TOKENS, world, rank, SEQ)
            shards.append([Xs[gi][b:e].view(-1, SEQ, GROUPS[gi][0])])
        # a plan that forces every kind of hand-off: solves alternate between the ranks, loops sit on the OTHER rank
        plan = D.BlockPlan(solve_owner=[0, 1, 0, 1], loop_owner=[[1, 0, 1], [0], [1, 1], [1]])
        timers = {}
        res = D.quantize_block_parallel(shards, Ws, GROUPS, bits=4, sym=False, eps=1e-4, block_size=1024,
                                        timers=timers, plan=plan)
        torch.cuda.synchronize()
        assert len(timers["allreduce"]) == len(GROUPS)
        auto = D.quantize_block_parallel(shards, Ws, GROUPS, bits=4, sym=False, eps=1e-4, block_size=1024)   # LPT plan
        torch.cuda.synchronize()
        torch.save({"forced": {k: (v.codes.cpu(), v.rank, v.final_W.cpu()) for k, v in res.items()},
                    "auto": {k: (v.codes.cpu(), v.rank) for k, v in auto.items()}},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_block_parallel_two_ranks_equals_one_rank(tmp_path):
    import gptq_svd_b200 as G
    world = 2
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got_forced, got_auto = {}, {}
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"rank{r}.pt"))
        got_forced.update(d["forced"])
        got_auto.update(d["auto"])
    want_keys = {(gi, li) for gi, (_, outs) in enumerate(GROUPS) for li in range(len(outs))}
    assert set(got_forced) == want_keys and set(got_auto) == want_keys        # every Linear quantised exactly once
    # one rank, same batches in the same order
    dev = torch.device("cuda", 0)
    Xs, Ws = _inputs(dev)
    for gi, (n, outs) in enumerate(GROUPS):
        acc = G.HessianAccumulator(n, dev)
        for r in range(world):
            b, e = TOKENS * r // world, TOKENS * (r + 1) // world
            acc.add_batch(Xs[gi][b:e].view(-1, SEQ, n))
        f = G.spectral_solve(acc.get_hessian(), 1e-4, "energy")
        for li in range(len(outs)):
            q = G.gptq_quantize(Ws[gi][li], f.R, G.Quantizer(4, 128, False), f.perm, 1024, True, f.R_x)
            codes, k, fw = got_forced[(gi, li)]
            assert k == f.k == got_auto[(gi, li)][1]
            assert torch.equal(codes, q.codes.cpu()), f"codes of Linear {(gi, li)} differ between 1 and 2 ranks"
            assert torch.equal(fw, q.final_W.cpu())
            assert torch.equal(got_auto[(gi, li)][0], q.codes.cpu())
