"""
Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src/TruncGPTQ/gptq_utils.py) on CPU through oracle/ref_stub.py.

Run in the build container only:  python oracle/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

Each fixture holds the inputs (X fp16 or H fp64, W fp32, knobs) and every stage
output of the reference: H, eigenvalues (torch.linalg.eigh, the call at
gptq_utils.py:93), k, perm, R_x, R, scale, zero, final_W for both loop variants
(Triton kernel under TRITON_INTERPRET=1 and the torch fallback loop) and the
relative error the reference logs (gptq_utils.py:291).
"""
from __future__ import annotations

import logging
import os
import sys

os.environ["TRITON_INTERPRET"] = "1"
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import truncgptq_oracle as O          # noqa: E402
from oracle.ref_stub import load_reference        # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [
    # name,         n,   m,  T,    dist,  eps,  method,   bits, group, sym, block
    ("llm_n128_w4a", 128, 64, 4096, "llm", 1e-4, "energy", 4, 128, False, 1024),
    ("llm_n128_w4s", 128, 64, 4096, "llm", 1e-2, "energy", 4, 128, True, 1024),
    ("llm_n256_w3a", 256, 96, 8192, "llm", 1e-4, "energy", 3, 128, False, 1024),
    ("llm_n256_w2a", 256, 40, 8192, "llm", 1e-5, "energy", 2, 128, False, 64),
    ("flat_n128_w4a", 128, 64, 4096, "flat", 1e-6, "energy", 4, 128, False, 1024),
    ("llm_n256_w4a_pc", 256, 32, 8192, "llm", 1e-3, "energy", 4, -1, False, 128),
    ("llm_n128_mt", 128, 16, 4096, "llm", 5e-4, "mean_trimmed", 4, 128, False, 1024),
    ("llm_n384_w4s", 384, 72, 8192, "llm", 1e-4, "energy", 4, 128, True, 1024),
]


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.rel = None

    def emit(self, record):
        msg = record.getMessage()
        if "Relative prediction error" in msg:
            self.rel = float(msg.split(":")[-1])


def main():
    G = load_reference(triton_interpret=True)
    os.makedirs(OUT, exist_ok=True)
    cap = _Capture()
    logging.getLogger().addHandler(cap)
    logging.getLogger().setLevel(logging.INFO)
    for i, (name, n, m, T, dist, eps, method, bits, group, sym, block) in enumerate(CASES):
        X = O.make_activations(T, n, seed=1000 + i, dist=dist)
        W = O.make_weight(m, n, seed=2000 + i)
        acc = G.HessianAccumulator(n, "cpu")
        Xt = torch.from_numpy(X)
        for c in range(0, T, 1024):                       # several add_batch calls, one 3-D
            xb = Xt[c:c + 1024]
            acc.add_batch(xb.reshape(2, -1, n) if c == 0 else xb)
        H = acc.get_hessian()
        L = torch.linalg.eigh(H.double())[0]
        eig = torch.sqrt(L.clamp(min=1e-12)).flip(0) ** 2
        R, R_x, perm = G.process_hessian_alt(H, eps, method)
        out = {}
        for tag, use_triton in (("triton", True), ("torch", False)):
            q = G.Quantizer(bits, group, sym)
            cap.rel = None
            fw, k = G.gptq_fwrd(torch.from_numpy(W), R, q, perm, block_size=block,
                                use_triton=use_triton, R_x=R_x)
            out[f"final_W_{tag}"] = fw.numpy()
            out[f"rel_err_{tag}"] = np.float64(cap.rel)
            out["scale"] = q.scale.numpy()
            out["zero"] = q.zero.numpy()
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            X=X if name in ("llm_n128_w4a", "flat_n128_w4a") else X[:0], W=W, H=H.numpy(), eig=eig.numpy(),
            k=np.int64(k), perm=perm.numpy(), R=R.numpy(), R_x=R_x.numpy(),
            eps=np.float64(eps), method=np.array(method), bits=np.int64(bits),
            group=np.int64(group), sym=np.bool_(sym), block=np.int64(block),
            n_tokens=np.int64(T), **out)
        print(f"{name}: n={n} m={m} k={k} rel_triton={out['rel_err_triton']:.5f} "
              f"rel_torch={out['rel_err_torch']:.5f}")


if __name__ == "__main__":
    main()
