"""First GPU run of the experimental two-stage tridiagonal reduction: tq_eigh with the one-stage and
the two-stage reduction side by side on the same Hessian - wall time (CUDA events), eigenvalue agreement, residual
and orthogonality of the two-stage eigenvectors.  With TQ_TRACE=1 the library prints its stage timers (sy2sb, sb2st
and the chase kernel's per-task cycle split, apply_q2, apply_q1) to stderr.
Usage: [TQ_TRACE=1] [TQ_SY2SB_GEMM=1] [TQ_Q2_UNBATCHED=1] [TQ_CHASE_GRID=g] python scripts/two_stage_probe.py [n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gptq_svd_b200 import _lib
from gptq_svd_b200 import stages as S
from scripts.solver_sweep import make_h


def timed_eigh(H):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w, V = S.eigh(H)
    e1.record()
    torch.cuda.synchronize()
    return w, V, e0.elapsed_time(e1)


def main():
    lib = _lib.load()
    ns = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 12288]
    for n in ns:
        H = make_h(n)
        rec = {"n": n}
        ref_w = None
        for name, flag in (("one_stage", 0), ("two_stage", 1)):
            lib.tq_set_eigh_two_stage(flag)
            timed_eigh(H)                                  # warm-up: attributes, cuBLAS heuristics, allocator
            w, V, ms = timed_eigh(H)
            rec[name + "_ms"] = round(ms, 2)
            if ref_w is None:
                ref_w = w
            else:
                rec["max_rel_dw"] = float((w - ref_w).abs().max() / ref_w.abs().max())
                nrm = float(torch.linalg.norm(H))
                rec["residual"] = float(torch.linalg.norm(H @ V - V * w[None, :])) / nrm
                rec["orthogonality"] = float(torch.linalg.norm(V.T @ V - torch.eye(n, device="cuda", dtype=torch.float64)))
            del w, V
        lib.tq_set_eigh_two_stage(-1)
        print(json.dumps(rec), flush=True)
        del H
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
