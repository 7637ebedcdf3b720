"""
Import the UNMODIFIED reference (`/root/reference/src/TruncGPTQ/gptq_utils.py`)
on a CPU-only box.  TEST INFRASTRUCTURE ONLY (golden-vector generation and
oracle validation in the build container; /root/reference does not exist on the
GPU box, so nothing under `-m gpu`, smoke() or bench.py calls this).

The reference imports `jax` for one call, `jax.scipy.linalg.qr(pivoting=True)`
(gptq_utils.py:114), which dispatches to MAGMA's dgeqp3 on GPU.  jax is not in
this image; the stub below provides that one entry point through
scipy.linalg.qr(pivoting=True) = LAPACK dgeqp3 (the algorithm MAGMA implements),
plus the no-op config hooks the module touches at import time
(gptq_utils.py:26-29).  SURVEY.md Appendix A.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = os.environ.get("TQ_REFERENCE_SRC", "/root/reference/src/TruncGPTQ")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "gptq_utils.py"))


def load_reference(triton_interpret: bool = False):
    """Returns the reference's gptq_utils module (imported once)."""
    if "gptq_utils" in sys.modules and getattr(sys.modules["gptq_utils"], "__tq_ref__", False):
        return sys.modules["gptq_utils"]
    if not reference_available():
        raise RuntimeError(f"reference sources not found at {REFERENCE_SRC}")
    if triton_interpret:
        os.environ["TRITON_INTERPRET"] = "1"
    import numpy as np
    import scipy.linalg as sla
    import torch

    jax = types.ModuleType("jax")
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax.clear_caches = lambda: None
    dl = types.ModuleType("jax.dlpack")
    dl.from_dlpack = lambda t: t
    jsp = types.ModuleType("jax.scipy")
    jl = types.ModuleType("jax.scipy.linalg")

    def qr(a, pivoting=False, mode="economic"):
        q, r, p = sla.qr(a.numpy(), mode="economic", pivoting=True)
        return torch.from_numpy(q), torch.from_numpy(r), torch.from_numpy(p.astype(np.int64))

    jl.qr = qr
    jsp.linalg = jl
    jax.scipy = jsp
    jax.dlpack = dl
    sys.modules.update({"jax": jax, "jax.dlpack": dl, "jax.scipy": jsp, "jax.scipy.linalg": jl})
    sys.path.insert(0, REFERENCE_SRC)
    import gptq_utils as G  # noqa: E402  (the reference module)

    G.__tq_ref__ = True
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None   # gptq_utils.py:558 is unconditional
    return G
