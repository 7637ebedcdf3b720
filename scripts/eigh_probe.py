"""eigh alone (tq_eigh) on a synthetic Hessian, for ncu captures of the sytrd panel kernel.
Usage: python scripts/eigh_probe.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptq_svd_b200 import stages as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n + 64, n, device="cuda", dtype=torch.float64, generator=g)
H = X.T @ X / n
del X
w, V = S.eigh(H)
torch.cuda.synchronize()
print("ok", float(w[-1]))
