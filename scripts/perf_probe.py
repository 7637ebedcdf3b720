"""Stage timings on one B200 (CUDA events): SYRK TFLOP/s vs kc, solver stages, loop.
Usage: python scripts/perf_probe.py [n ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G
from gptq_svd_b200 import stages as S


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def make_x(rows, n, seed=0, decay=-1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(n, n, device="cuda", generator=g) * torch.logspace(0, decay, n, device="cuda")[None, :]
    out = torch.empty(rows, n, device="cuda", dtype=torch.float16)
    for c in range(0, rows, 8192):
        z = torch.randn(min(8192, rows - c), n, device="cuda", generator=g)
        x = z @ A.T
        out[c:c + 8192] = (x / (n ** 0.5) * 3).half()
    out[:, :8] *= 30
    return out


def main():
    ns = [int(a) for a in sys.argv[1:]] or [4096]
    for n in ns:
        rows = 65536
        X = make_x(rows, n)
        print(f"== n={n}")
        ref = None
        for kc in (256, 512, 1024, 4096):
            acc = G.HessianAccumulator(n, "cuda", kc_tokens=kc)
            acc.add_batch(X)      # warm
            acc = G.HessianAccumulator(n, "cuda", kc_tokens=kc)
            ms, _ = timed(lambda: acc.add_batch(X), reps=3)
            if ref is None:
                ref = torch.zeros(n, n, device="cuda", dtype=torch.float64)
                for c in range(0, rows, 8192):
                    xb = X[c:c + 8192].double()
                    ref += xb.T @ xb
                ref *= 3
            err = float(torch.linalg.norm(acc.H - ref) / torch.linalg.norm(ref))
            print(f"syrk kc={kc}: {ms:.3f} ms/batch  {rows * n * n / ms / 1e9:.1f} TFLOP/s (algorithmic T*n^2)  rel_fro={err:.2e}")
        H = ref / (3 * rows)
        ms, (w, V) = timed(lambda: S.eigh(H))
        print(f"eigh: {ms:.1f} ms")
        ms2, wr = timed(lambda: torch.linalg.eigvalsh(H))
        print(f"  (torch/cuSOLVER eigvalsh for scale: {ms2:.1f} ms)  max|dw|/|w|max={float((w - wr).abs().max() / wr.abs().max()):.2e}")
        for eps in (1e-4,):
            ms, f = timed(lambda: G.spectral_solve(H, eps, "energy"))
            print(f"spectral_solve eps={eps}: {ms:.1f} ms  k={f.k}")
        Sm = (f.eigvals[:f.k].sqrt()[:, None] * V.T.flip(0)[:f.k]).contiguous()
        ms, (Rx, perm) = timed(lambda: S.qrcp(Sm))
        print(f"qrcp k={f.k}: {ms:.1f} ms  perm_equal={bool((perm[:f.k] == f.perm[:f.k]).all())}")
        ms, _ = timed(lambda: S.qr_r(Sm))
        print(f"qr_r k={f.k}: {ms:.1f} ms")
        for m in (4096, 12288) if n == 4096 else (4096,):
            W = (torch.randn(m, n, device="cuda") * 0.02).half().float()
            q = G.Quantizer(4, 128, True)
            G.gptq_fwrd(W, f.R, q, f.perm, block_size=1024)
            ms, _ = timed(lambda: G.gptq_fwrd(W, f.R, G.Quantizer(4, 128, True), f.perm, block_size=1024, R_x=f.R_x))
            print(f"gptq_fwrd m={m} n={n} (+error metric): {ms:.1f} ms")
        del X, ref, H


if __name__ == "__main__":
    main()
