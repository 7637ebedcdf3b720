"""cuBLAS DGEMM rates on this B200 for the shapes the solver uses (torch.matmul fp64 -> cublasDgemm)."""
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for n in (4096, 12288):
    C = torch.randn(n, n, device=dev, dtype=torch.float64)
    for K in (32, 64, 128, 256, 512, 1024):
        A = torch.randn(n, K, device=dev, dtype=torch.float64)
        B = torch.randn(K, n, device=dev, dtype=torch.float64)
        ms = t(lambda: C.addmm_(A, B, alpha=-1.0))
        print(f"n={n} rank-{K} update C-=A B: {ms:.3f} ms  {2*n*n*K/ms/1e9:.1f} TF/s  ({2*n*n*8/ms/1e6:.0f} GB/s C traffic)")
        At = torch.randn(n, K, device=dev, dtype=torch.float64)
        ms = t(lambda: torch.mm(At.T, C[:, :n]))
        print(f"n={n} K^T: (Kxn)(nxn) V^T C K={K}: {ms:.3f} ms  {2*n*n*K/ms/1e9:.1f} TF/s")
    ms = t(lambda: torch.mm(C, C), it=2)
    print(f"n={n} square: {ms:.2f} ms {2*n**3/ms/1e9:.1f} TF/s")
    del C
