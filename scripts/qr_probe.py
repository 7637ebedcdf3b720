"""Unpivoted QR stage alone (tq_qr_r) on a k x n matrix, for an ncu launch list.
Usage: python scripts/qr_probe.py [k n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptq_svd_b200 import stages as S
k = int(sys.argv[1]) if len(sys.argv) > 1 else 11030
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12288
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(k, n, device="cuda", dtype=torch.float64, generator=g)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = S.qr_r(A)
torch.cuda.synchronize()
e0.record(); R = S.qr_r(A); e1.record(); torch.cuda.synchronize()
print(f"qr_r k={k} n={n}: {e0.elapsed_time(e1):.1f} ms (incl. layout conversion)")
