"""CPU (gloo, world_size 2): host-side logic of the multi-GPU path - token sharding, LPT
placement, the Hessian all-reduce and the factor broadcast."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gptq_svd_b200 import dist as D

QWEN3_8B = [(4096, [4096, 1024, 1024]), (4096, [4096]), (4096, [12288, 12288]), (12288, [4096])]


def test_shard_range_covers_everything():
    for total, world, mult in [(262144, 8, 2048), (1000, 3, 64), (5, 8, 1), (128 * 2048, 4, 2048)]:
        got = []
        for r in range(world):
            b, e = D.shard_range(total, world, r, mult)
            assert 0 <= b <= e <= total and (b % mult == 0 or b == total)
            got += list(range(b, e))
        assert got == list(range(total))


def test_lpt_and_block_plan():
    assert D.lpt_assign([5, 3, 3, 1], 2) == [0, 1, 1, 0]
    assert D.lpt_assign([1, 1, 1], 1) == [0, 0, 0]
    for world in (1, 2, 4, 8):
        p = D.plan_block(QWEN3_8B, world)
        assert len(p.solve_owner) == 4 and all(0 <= r < world for r in p.solve_owner)
        assert [len(x) for x in p.loop_owner] == [3, 1, 2, 1]
        if world >= 2:      # the n=12288 solve (27x an n=4096 solve) gets a rank of its own
            big = p.solve_owner[3]
            assert all(r != big for r in p.solve_owner[:3])


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, T = 32, 64
        rng = np.random.RandomState(0)
        X = rng.standard_normal((T, n))
        b, e = D.shard_range(T, world, rank, 8)
        Xl = torch.from_numpy(X[b:e])
        H = Xl.T @ Xl
        total = D.allreduce_hessian(H, e - b)
        assert total == T
        assert torch.allclose(H, torch.from_numpy(X.T @ X), rtol=1e-13, atol=1e-13)
        gathered = [torch.empty_like(H) for _ in range(world)]
        dist.all_gather(gathered, H)
        assert all(torch.equal(g, gathered[0]) for g in gathered)          # bit-identical on every rank
        f = None
        if rank == 1:
            f = D.SpectralFactors(R=torch.arange(3 * n, dtype=torch.float64).reshape(3, n), R_x=torch.ones(3, n, dtype=torch.float64),
                                  perm=torch.arange(n), eigvals=torch.ones(n, dtype=torch.float64), k=3)
        g = D.broadcast_factors(f, n, 1, torch.device("cpu"))
        assert g.k == 3 and g.R.shape == (3, n) and float(g.R[2, 5]) == 2 * n + 5 and int(g.perm[7]) == 7
        # point-to-point hand-offs in the global exchange order (what quantize_block_parallel does with factors)
        groups = [(n, [8, 8]), (n, [8]), (2 * n, [8])]
        plan = D.BlockPlan(solve_owner=[0, 1, 0], loop_owner=[[1, 0], [0], [1]])
        order = D.exchange_order(groups, plan)
        assert order == [(0, 0, 1), (1, 1, 0), (2, 0, 1)]            # narrow groups first, every (src, dst) once
        mine = {gi: D.SpectralFactors(R=torch.full((2, nn), float(gi)), R_x=torch.full((2, nn), 10.0 + gi),
                                      perm=torch.arange(nn), eigvals=torch.ones(nn, dtype=torch.float64), k=2)
                for gi, (nn, _) in enumerate(groups) if plan.solve_owner[gi] == rank}
        mine = {gi: D.SpectralFactors(R=f.R.double(), R_x=f.R_x.double(), perm=f.perm, eigvals=f.eigvals, k=f.k)
                for gi, f in mine.items()}
        got = {}
        for gi, src, dst in order:
            if rank == src:
                D.send_factors(mine[gi], dst)
            elif rank == dst:
                got[gi] = D.recv_factors(groups[gi][0], src, torch.device("cpu"))
        want = {0: [1, 2], 1: [0]}[1 - rank] if False else ([1] if rank == 0 else [0, 2])
        assert sorted(got) == want
        for gi, f in got.items():
            assert f.k == 2 and float(f.R[1, 3]) == float(gi) and float(f.R_x[0, 0]) == 10.0 + gi
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_allreduce_and_broadcast_gloo(tmp_path):
    world = 2
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
