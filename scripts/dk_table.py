"""SURVEY H1 protocol (ii): end-to-end retained rank from an fp32-accumulated Hessian.  The tcgen05 SYRK
accumulates kc tokens in fp32 TMEM and adds chunk sums in fp32 / batches in fp64; the reference accumulates in fp64
(/root/reference/src/TruncGPTQ/gptq_utils.py:215-222).  For eps in 1e-4 .. 1e-7 this reports
  k_ref  = oracle solver on the oracle's fp64 H,
  k_gpu  = CUDA solver on the CUDA SYRK's H (end to end),
  k_mix  = CUDA solver on the oracle's fp64 H (stage-wise; must equal k_ref),
and the energy the differing directions carry.   Usage: python scripts/dk_table.py [n=1024] [T=262144] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

EPS = (1e-4, 1e-5, 1e-6, 1e-7)


def dk_table(n=1024, T=262144, seed=5):
    import gptq_svd_b200 as G
    from oracle import truncgptq_oracle as O
    X = O.make_activations(T, n, seed=seed, dist="llm")
    Xg = torch.from_numpy(X).cuda()
    Hd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    acc = G.HessianAccumulator(n, "cuda")
    for c in range(0, T, 65536):
        xb = Xg[c:c + 65536]
        Hd += xb.double().T @ xb.double()
        acc.add_batch(xb)
    H_ref = (Hd / T).cpu().numpy()
    H_ref = (H_ref + H_ref.T) / 2
    H_gpu = acc.get_hessian()
    relH = float(np.linalg.norm(H_gpu.cpu().numpy() - H_ref) / np.linalg.norm(H_ref))
    e = np.maximum(np.linalg.eigvalsh(H_ref), 1e-12)[::-1]
    rows = []
    for eps in EPS:
        k_ref = O.rank_rule(np.sqrt(e) ** 2, eps, "energy")
        k_gpu = G.spectral_solve(H_gpu, eps, "energy").k
        k_mix = G.spectral_solve(torch.from_numpy(H_ref).cuda(), eps, "energy").k
        lo, hi = min(k_ref, k_gpu), max(k_ref, k_gpu)
        rows.append({"eps": eps, "k_ref": int(k_ref), "k_gpu_end_to_end": int(k_gpu), "k_gpu_on_fp64_H": int(k_mix),
                     "dk": int(k_gpu - k_ref), "energy_frac_of_differing_directions": float(e[lo:hi].sum() / e.sum()),
                     "lambda_k_over_lambda_1": float(e[k_ref - 1] / e[0])})
    return {"n": n, "T": T, "dist": "llm (cond ~1e10)", "rel_fro_H": relH, "rows": rows}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    out = dk_table(n, T)
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            json.dump(out, f, indent=1)
