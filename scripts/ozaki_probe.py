"""Prototype for VERDICT r1 #6: fp64-grade GEMM from exact int8 tensor-core products (Ozaki splitting), to decide
whether a hand-written tcgen05 `kind::i8` version is worth building for the solver's BLAS-3.

Not product code: the int8 products go through `torch._int_mm` (cuBLASLt), which is the UPPER bound a hand-written
kernel of the same structure would be measured against.  For every shape the solver's DGEMMs have and every slice
count s the script reports
    err      max |C - C64| / (|A| |B|)_ij      (C64 = cuBLAS DGEMM; the bar for the solver is ~1e-15 .. 1e-14)
    ms       slicing + s (s + 1) / 2 int8 GEMMs + fp64 recombination, CUDA events, median of 5
    TF/s     2 m n k / ms, against the DGEMM's
One JSON line per (shape, s) on stdout.

Splitting: rows of A / columns of B are scaled by a power of two so that |x| < 1, then peeled 6 bits at a time
(slices in [-64, 64] as int8, so every int32 dot product over k <= 2^19 terms is exact); products A_p B_q with
p + q < s are kept (the triangle), each added to the fp64 result with its own power-of-two weight.
"""
import json
import sys

import torch

BITS = 6


def slices(X, s, dim):
    """X (fp64) -> (list of s int8 tensors, fp64 scale per row (dim=1) or column (dim=0))."""
    amax = X.abs().amax(dim=dim, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax)) + 1            # |X| / 2^e < 1/2
    scale = torch.exp2(e)
    r = X / scale
    out = []
    for _ in range(s):
        r = r * (1 << BITS)
        q = torch.round(r)
        out.append(q.to(torch.int8))
        r = r - q
    return out, scale


def ozaki_mm(A, B, s):
    As, sa = slices(A, s, 1)
    Bs, sb = slices(B, s, 0)
    C = torch.zeros(A.shape[0], B.shape[1], device=A.device, dtype=torch.float64)
    for p in range(s):
        for q in range(s - p):
            P = torch._int_mm(As[p], Bs[q])
            C.add_(P.double(), alpha=2.0 ** (-BITS * (p + q + 2)))
    return C * sa * sb


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [("rank-128 trailing update (sy2sb / pchol)", 8192, 8192, 128),
              ("back-transform (Q1/Q2, t = 1200 columns)", 12288, 1200, 4096),
              ("D&C merge / R-factor products", 6144, 6144, 6144)]
    for name, m, n, k in shapes:
        A = torch.randn(m, k, device=dev, dtype=torch.float64) * torch.logspace(0, -3, k, device=dev, dtype=torch.float64)
        B = torch.randn(k, n, device=dev, dtype=torch.float64)
        C64 = A @ B
        bound = A.abs() @ B.abs()
        ms64 = timed(lambda: A @ B)
        print(json.dumps({"shape": name, "m": m, "n": n, "k": k, "impl": "cuBLAS DGEMM", "ms": round(ms64, 3),
                          "TF/s": round(2 * m * n * k / ms64 / 1e9, 1)}))
        sys.stdout.flush()
        for s in (4, 6, 8, 9, 10):
            try:
                C = ozaki_mm(A, B, s)
            except Exception as ex:                      # int8 GEMM not available for this shape / build
                print(json.dumps({"shape": name, "slices": s, "error": repr(ex)[:200]}))
                break
            err = float(((C - C64).abs() / bound).max())
            ms = timed(lambda: ozaki_mm(A, B, s))
            As, _ = slices(A, s, 1)
            Bs, _ = slices(B, s, 0)
            ms_mm = timed(lambda: [torch._int_mm(As[p], Bs[q]) for p in range(s) for q in range(s - p)])
            print(json.dumps({"shape": name, "slices": s, "int8_gemms": s * (s + 1) // 2, "err_vs_|A||B|": err,
                              "ms_total": round(ms, 3), "ms_int8_gemms_only": round(ms_mm, 3),
                              "TF/s_total": round(2 * m * n * k / ms / 1e9, 1),
                              "TF/s_gemms_only": round(2 * m * n * k / ms_mm / 1e9, 1),
                              "int8_TOPS_gemms_only": round(s * (s + 1) / 2 * 2 * m * n * k / ms_mm / 1e9, 1),
                              "vs_dgemm_total": round(ms64 / ms, 2), "vs_dgemm_gemms_only": round(ms64 / ms_mm, 2)}))
            sys.stdout.flush()
            del C, As, Bs


if __name__ == "__main__":
    main()
