"""ncu target: ONE tq_spectral_solve on an LLM-like Hessian of order n (default 12288: two-stage reduction,
eigenvector subset, pivoted Cholesky, R from R_x).  Usage: python scripts/ncu_solver_target.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G
from scripts.solver_sweep import make_h

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
H = make_h(n)
f = G.spectral_solve(H, 1e-4, "energy")
torch.cuda.synchronize()
print("ok n=", n, "k=", f.k)
