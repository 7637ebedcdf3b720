"""Debug: d / e / tau of sytrd (TQ_SYM_DEBUG=32), saved to a file for old-vs-new comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptq_svd_b200 import _lib, stages as S
from gptq_svd_b200.gptq_utils import _ptr, _stream
n = int(sys.argv[1]); out = sys.argv[2]
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(2 * n, n, device="cuda", dtype=torch.float64, generator=g)
H = (X.T @ X / (2 * n)).contiguous()
lib = _lib.load()
ws = S._ws(n, H.device)
w = torch.zeros(n, dtype=torch.float64, device="cuda")
V = torch.zeros((n, n), dtype=torch.float64, device="cuda")
st = lib.tq_eigh(_ptr(H), n, n, _ptr(w), _ptr(V), n, _ptr(ws), ws.numel(), _stream(H))
torch.cuda.synchronize()
torch.save({"d": w.cpu(), "e": V[0].cpu(), "tau": V[1].cpu()}, out)
if len(sys.argv) > 3:
    a = torch.load(sys.argv[3]); b = torch.load(out)
    for k in ("d", "e", "tau"):
        diff = (a[k] - b[k]).abs() / a[k].abs().max()
        bad = torch.nonzero(diff > 1e-10).flatten()
        print(k, "max rel diff", float(diff.max()), "first bad idx", bad[:8].tolist())
