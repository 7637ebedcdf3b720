"""Debug: dump (A22 u) of the first sytrd column (TQ_SYM_DEBUG=30) and compare with torch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptq_svd_b200 import _lib, stages as S
from gptq_svd_b200.gptq_utils import _ptr, _stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(2 * n, n, device="cuda", dtype=torch.float64, generator=g)
H = (X.T @ X / (2 * n)).contiguous()
lib = _lib.load()
ws = S._ws(n, H.device)
w = torch.zeros(n, dtype=torch.float64, device="cuda")
V = torch.empty((n, n), dtype=torch.float64, device="cuda")
st = lib.tq_eigh(_ptr(H), n, n, _ptr(w), _ptr(V), n, _ptr(ws), ws.numel(), _stream(H))
torch.cuda.synchronize()
ref = H[1:, 1:] @ H[1:, 0]
got = w[1:]
err = (got - ref).abs()
print("status", st, "max abs err", float(err.max()), "ref max", float(ref.abs().max()))
bad = torch.nonzero(err > 1e-10 * ref.abs().max()).flatten()
print("bad rows:", bad.numel(), bad[:40].tolist())
if bad.numel():
    i = int(bad[0]); print("first bad: got", float(got[i]), "ref", float(ref[i]))
