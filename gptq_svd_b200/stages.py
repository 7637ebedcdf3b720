"""Stage-level entry points of the spectral solver (tq_eigh, tq_rank_select, tq_qrcp,
tq_qr_r) for stage-wise parity tests and profiling.  `process_hessian_alt` runs the same
stages inside one `tq_spectral_solve` call."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check
from .gptq_utils import _METHOD_CODE, _ptr, _require_cuda, _stream, _workspace


def _ws(n, dev):
    lib = _lib.load()
    nbytes = C.c_size_t(0)
    check(lib.tq_solver_workspace(n, C.byref(nbytes)), "tq_solver_workspace")
    return _workspace(nbytes.value, dev)


def eigh(H: torch.Tensor):
    """(w ascending, V) with eigenvector i in COLUMN i, like torch.linalg.eigh (gptq_utils.py:93)."""
    _require_cuda(H, "eigh")
    lib = _lib.load()
    n = H.shape[0]
    Hd = H.to(torch.float64).contiguous()
    with torch.cuda.device(H.device):
        ws = _ws(n, H.device)
        w = torch.empty(n, dtype=torch.float64, device=H.device)
        Vrows = torch.empty((n, n), dtype=torch.float64, device=H.device)
        check(lib.tq_eigh(_ptr(Hd), Hd.stride(0), n, _ptr(w), _ptr(Vrows), n, _ptr(ws), ws.numel(), _stream(Hd)),
              "tq_eigh")
    return w, Vrows.T


def rank_select(w_asc: torch.Tensor, threshold: float, method: str):
    """(eigvals clamped descending, k) by the rule of gptq_utils.py:94-108."""
    _require_cuda(w_asc, "rank_select")
    lib = _lib.load()
    n = w_asc.shape[0]
    w = w_asc.to(torch.float64).contiguous()
    with torch.cuda.device(w.device):
        ws = _workspace(4096, w.device)
        eig = torch.empty(n, dtype=torch.float64, device=w.device)
        k = C.c_int64(0)
        check(lib.tq_rank_select(_ptr(w), n, float(threshold), _METHOD_CODE.get(method, _lib.TQ_RANK_FULL),
                                 _ptr(eig), C.byref(k), _ptr(ws), ws.numel(), _stream(w)), "tq_rank_select")
    return eig, int(k.value)


def qrcp(A: torch.Tensor):
    """(R_x sign-normalised k x n, perm) of the column-pivoted QR of A (k x n, k <= n)."""
    _require_cuda(A, "qrcp")
    lib = _lib.load()
    k, n = A.shape
    Ad = A.to(torch.float64).contiguous()
    with torch.cuda.device(A.device):
        ws = _ws(n, A.device)
        R = torch.empty((k, n), dtype=torch.float64, device=A.device)
        perm = torch.empty(n, dtype=torch.int64, device=A.device)
        check(lib.tq_qrcp(_ptr(Ad), Ad.stride(0), k, n, _ptr(R), n, _ptr(perm), _ptr(ws), ws.numel(), _stream(Ad)),
              "tq_qrcp")
    return R, perm


def qr_r(A: torch.Tensor):
    """Sign-normalised R (k x n) of the unpivoted QR of A (k x n, k <= n)."""
    _require_cuda(A, "qr_r")
    lib = _lib.load()
    k, n = A.shape
    Ad = A.to(torch.float64).contiguous()
    with torch.cuda.device(A.device):
        ws = _ws(n, A.device)
        R = torch.empty((k, n), dtype=torch.float64, device=A.device)
        check(lib.tq_qr_r(_ptr(Ad), Ad.stride(0), k, n, _ptr(R), n, _ptr(ws), ws.numel(), _stream(Ad)), "tq_qr_r")
    return R


def two_stage_debug(H: torch.Tensor):
    """Experimental two-stage tridiagonal reduction, stage by stage (tq_two_stage_debug): returns
    (band, d, e) - band[c, i - c] = B[i, c] of the band matrix after stage 1 (bandwidth 64), (d, e) the tridiagonal
    matrix after the bulge chase.  H, B and T share their eigenvalues."""
    _require_cuda(H, "two_stage_debug")
    lib = _lib.load()
    n = H.shape[0]
    Hd = H.to(torch.float64).contiguous()
    with torch.cuda.device(H.device):
        lib.tq_set_eigh_two_stage(1)
        try:
            ws = _ws(n, H.device)
        finally:
            lib.tq_set_eigh_two_stage(-1)
        band = torch.empty((n, 128), dtype=torch.float64, device=H.device)
        d = torch.empty(n, dtype=torch.float64, device=H.device)
        e = torch.empty(n, dtype=torch.float64, device=H.device)
        check(lib.tq_two_stage_debug(_ptr(Hd), Hd.stride(0), n, _ptr(band), _ptr(d), _ptr(e), _ptr(ws), ws.numel(),
                                     _stream(Hd)), "tq_two_stage_debug")
    return band, d, e[: n - 1]
