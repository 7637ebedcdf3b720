"""gptq_fwrd alone on one Linear shape (for an ncu launch list of the loop kernels).
Usage: python scripts/loop_probe.py [m n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12288
k = int(n * 0.9)
g = torch.Generator(device="cuda").manual_seed(0)
R = torch.triu(torch.randn(k, n, device="cuda", dtype=torch.float64, generator=g)) * 0.05
R += torch.eye(k, n, device="cuda", dtype=torch.float64) * 2
Rx = R.clone()
perm = torch.randperm(n, device="cuda", generator=g)
W = (torch.randn(m, n, device="cuda", generator=g) * 0.02).half()
def run():
    q = G.Quantizer(4, 128, True)
    return G.gptq_fwrd(W, R, q, perm, block_size=1024, use_triton=True, R_x=Rx)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"gptq_fwrd m={m} n={n} k={k}: {e0.elapsed_time(e1):.2f} ms")
# per-kernel breakdown of one more call (sampled timing API of the library, every launch)
import ctypes as C
from gptq_svd_b200 import _lib
lib = _lib.load()
lib.tq_profile_begin(1)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); q = G.Quantizer(4, 128, True); G.gptq_fwrd(W, R, q, perm, block_size=1024, use_triton=True); t1.record()
torch.cuda.synchronize()
print(f"  without the error metric: {t0.elapsed_time(t1):.2f} ms")
W32 = W.float(); Q32 = W32 * 1.01
G.log_quantization_error(W32, Q32, Rx, perm); torch.cuda.synchronize()
t0.record(); err = G.log_quantization_error(W32, Q32, Rx, perm); t1.record()
torch.cuda.synchronize()
print(f"  error metric alone (cast + gather + tcgen05 GEMM + read-back): {t0.elapsed_time(t1):.2f} ms")
for kind, name in enumerate(_lib.PROF_KINDS):
    w, ms_, sa, to, wa, sm = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0), C.c_double(0), C.c_double(0)
    lib.tq_profile_kernel(kind, C.byref(w), C.byref(ms_), C.byref(sa), C.byref(to), C.byref(wa), C.byref(sm))
    if to.value:
        unit = _lib.PROF_UNIT[kind]
        rate = w.value / (ms_.value / 1e3) / (1e9 if unit == "B" else 1e12)
        print(f"  {name:26s} {to.value:4d} launches {ms_.value:8.3f} ms  {rate:9.1f} {'GB/s' if unit == 'B' else 'TFLOP/s'}")
lib.tq_profile_end(None, None, None, None)
