"""SYRK accumulation length kc (tokens per TMEM chunk): speed, H error and end-to-end k per eps."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gptq_svd_b200 as G
from oracle import truncgptq_oracle as O

def main():
    n, T = 1024, 262144
    X = O.make_activations(T, n, seed=5, dist="llm")
    Xg = torch.from_numpy(X).cuda()
    Hd = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for c in range(0, T, 65536):
        xb = Xg[c:c + 65536].double(); Hd += xb.T @ xb
    H_ref = (Hd / T).cpu().numpy(); H_ref = (H_ref + H_ref.T) / 2
    e = np.maximum(np.linalg.eigvalsh(H_ref), 1e-12)[::-1]
    kref = [O.rank_rule(np.sqrt(e) ** 2, eps, "energy") for eps in (1e-4, 1e-5, 1e-6, 1e-7)]
    print("k_ref", kref)
    Xbig = torch.randn(65536, 4096, device="cuda", dtype=torch.float16)
    for kc in (128, 256, 512, 1024):
        acc = G.HessianAccumulator(n, "cuda", kc_tokens=kc, verify=False)
        for c in range(0, T, 65536):
            acc.add_batch(Xg[c:c + 65536])
        Hg = acc.get_hessian()
        rel = float(np.linalg.norm(Hg.cpu().numpy() - H_ref) / np.linalg.norm(H_ref))
        ks = [G.spectral_solve(Hg, eps, "energy").k for eps in (1e-4, 1e-5, 1e-6, 1e-7)]
        a2 = G.HessianAccumulator(4096, "cuda", kc_tokens=kc, verify=False)
        a2.add_batch(Xbig); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): a2.add_batch(Xbig)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        a3 = G.HessianAccumulator(4096, "cuda", kc_tokens=kc, verify=True)
        a3.add_batch(Xbig); torch.cuda.synchronize()
        e0.record()
        for _ in range(10): a3.add_batch(Xbig)
        e1.record(); torch.cuda.synchronize()
        ms_v = e0.elapsed_time(e1) / 10
        print(json.dumps({"kc": kc, "rel_fro_H": rel, "k_gpu": ks, "dk": [a - b for a, b in zip(ks, kref)],
                          "syrk_ms_n4096_65536tok": ms, "tflops": 65536 * 4096 * 4096 / ms / 1e9, "ms_with_probe": ms_v}), flush=True)

if __name__ == "__main__":
    main()
