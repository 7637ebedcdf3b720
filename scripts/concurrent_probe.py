"""Several spectral solves in flight at once (gptq_svd_b200.concurrent.SolverPool) vs one after another.
Usage: python scripts/concurrent_probe.py [n] [count]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G
from gptq_svd_b200.concurrent import SolverPool
from scripts.solver_sweep import make_h

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cnt = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Hs = [make_h(n, seed=s) for s in range(cnt)]
G.spectral_solve(Hs[0], 1e-4, "energy")
torch.cuda.synchronize()
t0 = time.perf_counter()
seq = [G.spectral_solve(H, 1e-4, "energy") for H in Hs]
torch.cuda.synchronize()
t_seq = time.perf_counter() - t0
pool = SolverPool(workers=cnt)
for budget in (None, 148, 74, 49, 36):
    pool.spectral_solve_many(Hs, 1e-4, "energy", sm_budget=budget)      # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = pool.spectral_solve_many(Hs, 1e-4, "energy", sm_budget=budget)
    torch.cuda.synchronize()
    t_con = time.perf_counter() - t0
    same_k = all(r.k == q.k for r, q in zip(res, seq))
    same_p = all(torch.equal(r.perm[:r.k], q.perm[:q.k]) for r, q in zip(res, seq))
    dR = max(float((r.R - q.R).abs().max() / q.R.abs().max()) for r, q in zip(res, seq)) if same_k else -1
    dRx = max(float((r.R_x - q.R_x).abs().max() / q.R_x.abs().max()) for r, q in zip(res, seq)) if same_k else -1
    de = max(float((r.eigvals - q.eigvals).abs().max() / q.eigvals.abs().max()) for r, q in zip(res, seq))
    print(f"n={n} x{cnt}: sequential {t_seq*1e3:.1f} ms, concurrent (budget {budget}) {t_con*1e3:.1f} ms, "
          f"k equal {same_k}, perm[:k] equal {same_p}, max rel dR {dR:.2e} dRx {dRx:.2e} deig {de:.2e}", flush=True)
pool.close()

# one wide solve next to the narrow ones: does the bandwidth-bound solve hide the latency-bound ones?
if len(sys.argv) > 3:
    nb = int(sys.argv[3])
    Hb = make_h(nb, seed=9)
    G.spectral_solve(Hb, 1e-4, "energy")
    torch.cuda.synchronize()
    t0 = time.perf_counter(); G.spectral_solve(Hb, 1e-4, "energy"); torch.cuda.synchronize()
    t_big = time.perf_counter() - t0
    pool = SolverPool(workers=cnt + 1)
    for small_b, big_b in ((16, 100), (12, 112), (24, 76), (8, 124)):
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pool.spectral_solve_many([Hb] + Hs, 1e-4, "energy", sm_budget=[big_b] + [small_b] * cnt)
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
        print(f"wide n={nb} alone {t_big*1e3:.0f} ms; wide ({big_b} SMs) + {cnt} x n={n} ({small_b} SMs each) together {t_all*1e3:.0f} ms", flush=True)
    pool.close()
