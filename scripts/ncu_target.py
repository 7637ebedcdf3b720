"""Small single-GPU target for ncu: one n=4096 group of the hot path
(one 65536-token add_batch, spectral solve, one 4096x4096 gptq_fwrd)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G
from bench import make_x

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = make_x(torch, 65536, n, 0, -1.0)
acc = G.HessianAccumulator(n, "cuda")
acc.add_batch(X)
H = acc.get_hessian()
R, R_x, perm = G.process_hessian_alt(H, 1e-4, "energy")
W = (torch.randn(n, n, device="cuda") * 0.02).half()
fw, k = G.gptq_fwrd(W, R, G.Quantizer(4, 128, True), perm, block_size=1024, use_triton=True, R_x=R_x)
torch.cuda.synchronize()
print("ok k=", k)
