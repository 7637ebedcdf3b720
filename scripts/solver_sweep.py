"""Solver ms per n x n Hessian (the second half of BASELINE.json's metric): tq_spectral_solve
(eigh + rank rule + pivot order / R_x + R) on LLM-like synthetic Hessians, with oracle-free
invariants checked on the device:  R_x^T R_x = P^T H_k P,  (R^T R)(P^T H_k P) = projector.
Usage: python scripts/solver_sweep.py [n ...]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gptq_svd_b200 as G


DECAY = float(os.environ.get('SWEEP_DECAY', '-1.0'))
EPS = float(os.environ.get('SWEEP_EPS', '1e-4'))


def make_h(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(n, n, device="cuda", generator=g) * torch.logspace(0, DECAY, n, device="cuda")[None, :]
    H = torch.zeros(n, n, device="cuda", dtype=torch.float64)
    rows = max(2 * n, 8192)
    for c in range(0, rows, 8192):
        z = torch.randn(min(8192, rows - c), n, device="cuda", generator=g)
        x = (z @ A.T / n ** 0.5 * 3).half()
        x[:, :8] *= 30
        xd = x.double()
        H += xd.T @ xd
    return H / rows


def main():
    ns = [int(a) for a in sys.argv[1:]] or [1024, 2048, 3072, 4096, 8192, 12288]
    out = []
    for n in ns:
        H = make_h(n)
        torch.cuda.synchronize()
        if not os.environ.get("SWEEP_NO_WARM"):
            f = G.spectral_solve(H, EPS, "energy")      # warm-up (workspace allocation, attributes)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f = G.spectral_solve(H, EPS, "energy")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k = f.k
        P = f.perm
        # invariants in fp64 on the device (chunked to bound memory)
        Hp = H[P][:, P]
        RtR = f.R.T @ f.R                                 # P^T H_k^+ P
        RxtRx = f.R_x.T @ f.R_x                           # P^T H_k P
        w = f.eigvals.flip(0)      # our own spectrum (clamped at 1e-12, ascending); cuSOLVER syevd rejects n = 28672
        tail = float(w[: n - k].clamp(min=0).sum()) if k < n else 0.0
        e_hk = float(torch.linalg.norm(RxtRx - Hp)) / float(torch.linalg.norm(Hp))      # = ||H - H_k|| / ||H||
        M = RtR @ RxtRx                                   # projector onto the retained subspace
        e_proj = float(torch.linalg.norm(M @ M - M)) / float(torch.linalg.norm(M))
        tr = float(torch.trace(M))
        rec = {"n": n, "k": k, "solver_ms": round(ms, 2), "fp64_gflops_7.4n3": round(7.4 * n ** 3 / ms / 1e6, 1),
               "rel_Hk_vs_H": e_hk, "discarded_energy_frac": tail / float(w.clamp(min=0).sum()),
               "projector_err": e_proj, "trace_projector_minus_k": tr - k}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del H, f, Hp, RtR, RxtRx, M
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
