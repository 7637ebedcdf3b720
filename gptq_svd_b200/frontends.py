"""
The two other front ends of the reference that end in the same `gptq_fwrd` loop
(SURVEY.md section 8, rows f3 / f4), with the reference's names and signatures:

    process_hessian(H, actorder=False, damp_percent=0.01) -> (H_inv_chol, perm)   gptq_utils.py:129-165
    Sketcher(layer, rank, device)  .hook_fn / .get_scaled_sketch()                gptq_utils.py:171-211
    process_sketch(sketch, threshold=1e-2, threshold_method="mean_trimmed")       gptq_utils.py:33-84

Numerics run in libtruncgptq.so (`tq_cholesky_solve`, `tq_sketch_accum`, `tq_sketch_solve`);
PyTorch allocates, draws the Gaussian block (same `torch.randn` call as the reference, so a
seeded run reproduces it) and sorts the diagonal for act-order.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check
from .gptq_utils import _DTYPE_CODE, _METHOD_CODE, _ptr, _require_cuda, _stream, _workspace


def process_hessian(H: torch.Tensor, actorder: bool = False,
                    damp_percent: float = 0.01) -> Tuple[torch.Tensor, torch.Tensor]:
    """Upper Cholesky factor of the damped inverse Hessian and the (act-order) permutation.

    Same ladder as the reference: damp = 10^e * damp_percent for e = 0..4, first success wins
    and a higher rung is logged (gptq_utils.py:148-160).  When every rung fails the reference
    means to fall back to the identity (:161-163, unreachable there because of a misspelt
    variable); here the identity is returned."""
    _require_cuda(H, "process_hessian")
    lib = _lib.load()
    n = H.shape[0]
    if H.dim() != 2 or H.shape[1] != n:
        raise RuntimeError("process_hessian: H must be square")
    Hd = H.to(dtype=torch.float64)
    if Hd.stride(1) != 1:
        Hd = Hd.contiguous()
    dev = H.device
    if actorder:
        perm = torch.argsort(torch.diag(Hd), descending=True)        # gptq_utils.py:138
    else:
        perm = torch.arange(n, device=dev)
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.tq_cholesky_workspace(n, C.byref(nbytes)), "tq_cholesky_workspace")
        ws = _workspace(nbytes.value, dev)
        out = torch.empty((n, n), dtype=torch.float64, device=dev)
        e = C.c_int(0)
        check(lib.tq_cholesky_solve(_ptr(Hd), Hd.stride(0), n, _ptr(perm) if actorder else C.c_void_p(0),
                                    float(damp_percent), _ptr(out), n, C.byref(e), _ptr(ws), ws.numel(),
                                    _stream(Hd)), "tq_cholesky_solve")
    if e.value > 0:
        logging.info(f"  Ref-GPTQ required high damping: {10 ** e.value * damp_percent}")
    elif e.value < 0:
        logging.warning(" Hessian is singular. Using Identity fallback.")
    return out, perm


class Sketcher:
    """Y (rank x n fp32) += R_batch @ x per forward pass (gptq_utils.py:171-211)."""

    def __init__(self, layer: nn.Module, rank: int, device="cuda"):
        self.layer = layer
        self.rank = rank
        self.device = device
        self.in_features = layer.in_features
        self.Y = torch.zeros((rank, self.in_features), device=device, dtype=torch.float32).contiguous()
        self.n_samples = 0
        _require_cuda(self.Y, "Sketcher")

    def add_batch(self, x: torch.Tensor, R_batch: Optional[torch.Tensor] = None):
        """The body of hook_fn; `R_batch` lets a test supply the Gaussian block."""
        if x.dim() > 2:
            x = x.reshape(-1, x.shape[-1])
        batch_count = x.shape[0]
        if batch_count == 0:
            return
        _require_cuda(x, "Sketcher.hook_fn")
        self.n_samples += batch_count
        if x.dtype not in _DTYPE_CODE:
            x = x.to(torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        if R_batch is None:
            R_batch = torch.randn((self.rank, batch_count), device=self.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(self.Y.device):
            nbytes = C.c_size_t(0)
            check(lib.tq_sketch_accum_workspace(self.rank, batch_count, self.in_features, C.byref(nbytes)),
                  "tq_sketch_accum_workspace")
            ws = _workspace(nbytes.value, self.Y.device)
            check(lib.tq_sketch_accum(_ptr(self.Y), self.Y.stride(0), _ptr(R_batch), R_batch.stride(0), _ptr(x),
                                      _DTYPE_CODE[x.dtype], x.stride(0), self.rank, batch_count, self.in_features,
                                      _ptr(ws), ws.numel(), _stream(self.Y)),
                  "tq_sketch_accum")

    def hook_fn(self, module: nn.Module, input_args, output):
        self.add_batch(input_args[0])

    def get_scaled_sketch(self):
        if self.n_samples == 0:
            return None, None, 0                                   # gptq_utils.py:205-206
        self.Y.mul_(1.0 / math.sqrt(self.n_samples * self.rank))    # :208-210
        return self.Y


def process_sketch(sketch: torch.Tensor, threshold: float = 1e-2,
                   threshold_method: str = "mean_trimmed") -> Tuple[torch.Tensor, torch.Tensor]:
    """(R, perm) from the scaled sketch (gptq_utils.py:33-84); see `tq_sketch_solve`."""
    _require_cuda(sketch, "process_sketch")
    lib = _lib.load()
    rank, n = sketch.shape
    Y = sketch.to(torch.float32)
    if Y.stride(1) != 1:
        Y = Y.contiguous()
    dev = sketch.device
    method = _METHOD_CODE.get(threshold_method)
    if method is None:
        # the reference leaves current_rank undefined for any other method (:49-58)
        raise UnboundLocalError("process_sketch: threshold_method must be 'energy' or 'mean_trimmed'")
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.tq_sketch_workspace(rank, n, C.byref(nbytes)), "tq_sketch_workspace")
        ws = _workspace(nbytes.value, dev)
        R = torch.empty((n, n), dtype=torch.float64, device=dev)
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        k = C.c_int64(0)
        check(lib.tq_sketch_solve(_ptr(Y), Y.stride(0), rank, n, float(threshold), method, _ptr(R), _ptr(perm),
                                  C.byref(k), _ptr(ws), ws.numel(), _stream(Y)), "tq_sketch_solve")
    return R[:int(k.value)], perm
