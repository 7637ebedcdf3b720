"""Smallest case that goes through the symmetric TMA sytrd panel (n even, >= 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptq_svd_b200 import stages as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(2 * n, n, device="cuda", dtype=torch.float64, generator=g)
H = X.T @ X / (2 * n)
w, V = S.eigh(H)
torch.cuda.synchronize()
wr = torch.linalg.eigvalsh(H)
print("max rel eig err", float((w - wr).abs().max() / wr.abs().max()))
print("residual", float(torch.linalg.norm(H @ V - V * w[None, :]) / torch.linalg.norm(H)))
