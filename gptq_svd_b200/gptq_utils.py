"""
Host-side mirror of the reference's quantizer core (`src/TruncGPTQ/gptq_utils.py`),
backed by the sm_100a CUDA library `libtruncgptq.so` through its C ABI.

Drop-in surface (same names, argument meaning, return types and error behaviour as
the reference module imported at `src/TruncGPTQ/quantize.py:14`):

    HessianAccumulator(in_features, device, dtype=torch.float64)   gptq_utils.py:213-228
    process_hessian_alt(H, threshold, threshold_method)            gptq_utils.py:87-126
    Quantizer(w_bits, group_size, sym)                             gptq_utils.py:230-272
    gptq_fwrd(weight_mat, H_inv_sqrt, quantizer, perm, ...)        gptq_utils.py:459-565
    log_quantization_error(W_orig, W_quant, R_x, perm)             gptq_utils.py:275-291

The sibling front ends (process_hessian, Sketcher, process_sketch) are in frontends.py, the caller's
loop (quantize.main) in pipeline.py, several solves in flight on one GPU in concurrent.py.

New, additive (the reference has no integer output, README.md:133):
    gptq_quantize(...) -> QuantizedLinear(final_W, codes, scale, zero, rank)
    pack_codes(codes, bits)

PyTorch is used for device memory, streams and dtype plumbing only.  Every numerical
stage runs in hand-written CUDA; there is no CPU path - tensors must live on a B200.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check

_DTYPE_CODE = {torch.float16: _lib.TQ_F16, torch.bfloat16: _lib.TQ_BF16,
               torch.float32: _lib.TQ_F32, torch.float64: _lib.TQ_F64}
_METHOD_CODE = {"energy": _lib.TQ_RANK_ENERGY, "mean_trimmed": _lib.TQ_RANK_MEAN_TRIMMED}


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; the TruncGPTQ hot path runs on a B200 "
                           "(sm_100a) only and has no CPU fallback")


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------
class HessianAccumulator:
    """H (n x n fp64, on device) += X^T X per calibration batch (gptq_utils.py:213-228).

    `add_batch` runs the tcgen05 SYRK (`tq_syrk_accum_scaled`); `.H` is the un-normalised,
    fully symmetric fp64 sum after every call, `.n_samples` the token count,
    `get_hessian()` returns `H / n_samples` (or H itself when empty).

    fp16 / bf16 activations feed the tensor cores as they are (their products are exact in
    fp32).  fp32 / fp64 activations - which the reference widens to fp64 (:221) - are scaled
    by a per-batch power of two chosen on the device so that the largest magnitude lands in
    [2^13, 2^14), rounded to fp16 and the batch's contribution multiplied by 1 / scale^2: no
    overflow for any finite input; each entry keeps 11 significant bits (the rounding errors
    of T tokens average out: H stays within the 1e-5 bar, tests/test_gpu_syrk.py).  NaN /
    infinity in such a batch raises in `get_hessian()`.

    `verify=True` (default) keeps a one-number probe of the accumulation (v^T H v = sum of
    ||X_b v||^2 for a fixed sign vector v: one extra pass over X per batch) and
    `get_hessian()` raises RuntimeError when H disagrees with it - a guard against a
    silently corrupted Hessian, which would poison k, perm, R and every code of the group.
    """

    PROBE_TOL = 1e-4

    def __init__(self, in_features, device, dtype=torch.float64, kc_tokens: int = 0, verify: bool = True):
        if dtype != torch.float64:
            raise ValueError("HessianAccumulator: the accumulator is fp64 (reference default)")
        self.H = torch.zeros((in_features, in_features), device=device, dtype=dtype)
        _require_cuda(self.H, "HessianAccumulator")
        self.n_samples = 0
        self.kc_tokens = kc_tokens
        self.verify = bool(verify)
        # device scalars: [0] probe (f64), [1] v^T H v (f64), [2..5] cast scratch (32 B), [6] status (i32)
        self._dev = torch.zeros(8, dtype=torch.float64, device=self.H.device)

    def add_batch(self, x: torch.Tensor):
        if x.dim() == 3:
            x = x.reshape(-1, x.shape[-1])
        _require_cuda(x, "HessianAccumulator.add_batch")
        lib = _lib.load()
        n = self.H.shape[0]
        if x.shape[1] != n:
            raise RuntimeError(f"add_batch: expected {n} features, got {x.shape[1]}")
        rows = x.shape[0]
        if rows == 0:
            return
        if x.device != self.H.device:
            raise RuntimeError(f"add_batch: activations on {x.device}, accumulator on {self.H.device}")
        base = self._dev.data_ptr()
        alpha = C.c_void_p(0)
        with torch.cuda.device(x.device):
            if x.dtype not in (torch.float16, torch.bfloat16):
                src = x if x.stride(1) == 1 else x.contiguous()
                ldd = (n + 7) // 8 * 8
                x16 = torch.zeros((rows, ldd), dtype=torch.float16, device=x.device)
                check(lib.tq_cast_to_f16_scaled(_ptr(src), _DTYPE_CODE[src.dtype], rows, n, src.stride(0),
                                                _ptr(x16), ldd, C.c_void_p(base + 16), C.c_void_p(base + 48),
                                                _stream(x)), "tq_cast_to_f16_scaled")
                x, ldx = x16, ldd
                alpha = C.c_void_p(base + 32)
            else:
                if x.stride(1) != 1 or x.stride(0) % 8 != 0 or x.data_ptr() % 16 != 0:
                    ldd = (n + 7) // 8 * 8
                    xp = torch.zeros((rows, ldd), dtype=x.dtype, device=x.device)
                    xp[:, :n].copy_(x)
                    x, ldx = xp, ldd
                else:
                    ldx = x.stride(0)
            check(lib.tq_syrk_accum_scaled(_ptr(self.H), self.H.stride(0), _ptr(x), _DTYPE_CODE[x.dtype], rows, n,
                                           ldx, self.kc_tokens, alpha, _stream(x)), "tq_syrk_accum")
            if self.verify:
                check(lib.tq_hessian_probe_accum(_ptr(x), _DTYPE_CODE[x.dtype], rows, n, ldx, alpha,
                                                 C.c_void_p(base), _stream(x)), "tq_hessian_probe_accum")
        self.n_samples += rows

    def check(self):
        """Raise RuntimeError when a batch held NaN / infinity or when H fails the probe identity
        (one pass over H and a 4-byte read-back; called by get_hessian)."""
        lib = _lib.load()
        base = self._dev.data_ptr()
        n = self.H.shape[0]
        with torch.cuda.device(self.H.device):
            if self.verify and self.n_samples > 0:
                check(lib.tq_hessian_probe_check(_ptr(self.H), self.H.stride(0), n, C.c_void_p(base),
                                                 self.PROBE_TOL, C.c_void_p(base + 8), C.c_void_p(base + 48),
                                                 _stream(self.H)), "tq_hessian_probe_check")
            status = int(self._dev[6:7].view(torch.int32)[0].item())
        if status & 1:
            raise RuntimeError("HessianAccumulator: a calibration batch contains NaN or infinity")
        if status & 2:
            probe, vhv = self._dev[0].item(), self._dev[1].item()
            raise RuntimeError(f"HessianAccumulator: H fails the probe identity v^T H v = sum ||X v||^2 "
                               f"({vhv:.9e} vs {probe:.9e}): the accumulation is corrupted")

    def get_hessian(self) -> torch.Tensor:
        if self.n_samples == 0:
            return self.H
        self.check()
        lib = _lib.load()
        out = torch.empty_like(self.H)
        n = self.H.shape[0]
        with torch.cuda.device(self.H.device):
            check(lib.tq_hessian_scale(_ptr(self.H), self.H.stride(0), n, self.n_samples, _ptr(out),
                                       out.stride(0), _stream(out)), "tq_hessian_scale")
        return out


# ----------------------------------------------------------------------------
@dataclass
class SpectralFactors:
    R: torch.Tensor        # k x n fp64
    R_x: torch.Tensor      # k x n fp64
    perm: torch.Tensor     # n int64
    eigvals: torch.Tensor  # n fp64, clamped at 1e-12, descending
    k: int


def spectral_solve(H: torch.Tensor, threshold: float = 0.0005, threshold_method: str = "mean_trimmed",
                   householder_qrcp: bool = False) -> SpectralFactors:
    """process_hessian_alt plus the eigenvalues and k (one `tq_spectral_solve` call).

    perm / R_x come from the diagonally pivoted Cholesky of S^T S (same pivots and factor as
    the column-pivoted QR of S, BLAS-3 bound); `householder_qrcp=True` runs the LAPACK-style
    Householder QRCP of S instead (BLAS-2 bound)."""
    _require_cuda(H, "process_hessian_alt")
    lib = _lib.load()
    n = H.shape[0]
    if H.dim() != 2 or H.shape[1] != n:
        raise RuntimeError("process_hessian_alt: H must be square")
    Hd = H.to(dtype=torch.float64)
    if Hd.stride(1) != 1:
        Hd = Hd.contiguous()
    method = _METHOD_CODE.get(threshold_method, _lib.TQ_RANK_FULL)
    if householder_qrcp:
        method |= _lib.TQ_SOLVE_HOUSEHOLDER_QRCP
    dev = H.device
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.tq_solver_workspace(n, C.byref(nbytes)), "tq_solver_workspace")
        ws = _workspace(nbytes.value, dev)
        R = torch.empty((n, n), dtype=torch.float64, device=dev)
        Rx = torch.empty((n, n), dtype=torch.float64, device=dev)
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        eig = torch.empty(n, dtype=torch.float64, device=dev)
        k = C.c_int64(0)
        check(lib.tq_spectral_solve(_ptr(Hd), Hd.stride(0), n, float(threshold), method, _ptr(R), _ptr(Rx),
                                    _ptr(perm), _ptr(eig), C.byref(k), _ptr(ws), ws.numel(), _stream(Hd)),
              "tq_spectral_solve")
    kk = int(k.value)
    return SpectralFactors(R=R[:kk], R_x=Rx[:kk], perm=perm, eigvals=eig, k=kk)


def process_hessian_alt(H: torch.Tensor, threshold: float = 0.0005,
                        threshold_method: str = "mean_trimmed") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(R, R_x, perm) with R^T R = P^T H_k^+ P and R_x^T R_x = P^T H_k P (gptq_utils.py:87-126)."""
    f = spectral_solve(H, threshold, threshold_method)
    return f.R, f.R_x, f.perm


# ----------------------------------------------------------------------------
class Quantizer:
    """Static per-(row, group) quantisation grid (gptq_utils.py:230-272)."""

    def __init__(self, w_bits: int = 4, group_size: int = 128, sym: bool = False):
        self.w_bits = w_bits
        self.group_size = group_size
        self.sym = sym
        if self.sym:
            half_range = 2 ** (w_bits - 1) - 1
            self.max_q = half_range
            self.min_q = -half_range
        else:
            self.max_q = 2 ** w_bits - 1
            self.min_q = 0
        self.scale = None
        self.zero = None

    def find_params(self, weights: torch.Tensor):
        _require_cuda(weights, "Quantizer.find_params")
        m, n = weights.shape
        g_size = self.group_size if self.group_size > 0 else n
        assert n % g_size == 0          # gptq_utils.py:253
        lib = _lib.load()
        W = weights.to(torch.float32)
        if W.stride(1) != 1:
            W = W.contiguous()
        ng = n // g_size
        self.scale = torch.empty((m, ng, 1), dtype=torch.float32, device=W.device)
        self.zero = torch.empty((m, ng, 1), dtype=torch.float32, device=W.device)
        with torch.cuda.device(W.device):
            check(lib.tq_find_params(_ptr(W), W.stride(0), m, n, self.w_bits, self.group_size, int(self.sym),
                                     _ptr(self.scale), _ptr(self.zero), _stream(W)), "tq_find_params")

    def get_expanded_params(self, m, n):
        g = self.group_size if self.group_size > 0 else n
        s_expanded = torch.repeat_interleave(self.scale, g, dim=1)
        z_expanded = torch.repeat_interleave(self.zero, g, dim=1)
        return s_expanded[:, :n].squeeze(-1), z_expanded[:, :n].squeeze(-1)


# ----------------------------------------------------------------------------
def _relative_error(W32, Wq32, R_x, perm) -> float:
    lib = _lib.load()
    m, n = W32.shape
    k = R_x.shape[0]
    Rx = R_x if R_x.dtype in (torch.float32, torch.float64) else R_x.to(torch.float32)
    if Rx.stride(1) != 1:
        Rx = Rx.contiguous()
    dev = W32.device
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.tq_quant_error_workspace(m, n, k, C.byref(nbytes)), "tq_quant_error_workspace")
        ws = _workspace(nbytes.value, dev)
        out2 = torch.zeros(2, dtype=torch.float64, device=dev)
        check(lib.tq_quant_error(_ptr(W32), W32.stride(0), _ptr(Wq32), Wq32.stride(0), _ptr(Rx),
                                 _DTYPE_CODE[Rx.dtype], Rx.stride(0), k, _ptr(perm), m, n, _ptr(out2), _ptr(ws),
                                 ws.numel(), _stream(W32)), "tq_quant_error")
    num, den = out2.tolist()
    return float((num ** 0.5) / (den ** 0.5))


def log_quantization_error(W_orig: torch.Tensor, W_quant: torch.Tensor, R_x: torch.Tensor, perm: torch.Tensor):
    """Logs and returns ||(W-Q)[:,perm] R_x^T|| / ||W[:,perm] R_x^T|| (gptq_utils.py:275-291)."""
    if R_x is None or perm is None:
        return None
    W32 = W_orig.to(torch.float32).contiguous()
    Q32 = W_quant.to(torch.float32).contiguous()
    rel = _relative_error(W32, Q32, R_x, perm.to(torch.int64).contiguous())
    logging.info(f"   [Metric] Relative prediction error: {rel:.6f}")
    return rel


@dataclass
class QuantizedLinear:
    final_W: torch.Tensor           # m x n dequantised weights, input dtype, original column order
    codes: torch.Tensor             # m x n uint8, code - min_q
    scale: torch.Tensor             # m x n/g fp32
    zero: torch.Tensor              # m x n/g fp32
    rank: int
    min_q: int
    bits: int
    rel_error: Optional[float] = None


# Lazy trailing update W[:, i2:] -= E @ U: tcgen05 3xTF32 GEMM by default; set to True for the
# strict-fp32 SIMT GEMM (the reference runs this product with TF32 disabled, gptq_utils.py:474-475).
TRAILING_STRICT_FP32 = False


def gptq_quantize(weight_mat: torch.Tensor, H_inv_sqrt: torch.Tensor, quantizer: Quantizer, perm: torch.Tensor,
                  block_size: int = 128, use_triton: bool = True, R_x: Optional[torch.Tensor] = None,
                  want_codes: bool = True, strict_fp32: Optional[bool] = None) -> QuantizedLinear:
    """gptq_fwrd plus the integer codes and grid parameters."""
    if strict_fp32 is None:
        strict_fp32 = TRAILING_STRICT_FP32
    _require_cuda(weight_mat, "gptq_fwrd")
    lib = _lib.load()
    m, n = weight_mat.shape
    dev = weight_mat.device
    orig_dtype = weight_mat.dtype
    W32 = weight_mat.to(device=dev, dtype=torch.float32)
    if W32.stride(1) != 1:
        W32 = W32.contiguous()
    R = H_inv_sqrt.to(device=dev)
    if R.dtype not in (torch.float32, torch.float64):
        R = R.to(torch.float32)
    if R.dim() != 2 or R.shape[1] != n:
        raise RuntimeError(f"gptq_fwrd: H_inv_sqrt must be k x {n}, got {tuple(R.shape)}")
    if R.stride(1) != 1:
        R = R.contiguous()
    k = R.shape[0]
    if k < n:
        logging.info(f"   Rank percent used: {float(k) / n:.2%}")     # gptq_utils.py:487-488
    p64 = perm.to(device=dev, dtype=torch.int64).contiguous()
    quantizer.find_params(W32)
    out = torch.empty((m, n), dtype=torch.float32, device=dev)
    codes = torch.empty((m, n), dtype=torch.uint8, device=dev) if want_codes else None
    with torch.cuda.device(dev):
        nbytes = C.c_size_t(0)
        check(lib.tq_gptq_loop_workspace(m, n, k, C.byref(nbytes)), "tq_gptq_loop_workspace")
        ws = _workspace(nbytes.value, dev)
        check(lib.tq_gptq_loop(_ptr(W32), W32.stride(0), _ptr(R), _DTYPE_CODE[R.dtype], R.stride(0) if k else n,
                               k, _ptr(p64), _ptr(quantizer.scale), _ptr(quantizer.zero), m, n, quantizer.w_bits,
                               quantizer.group_size, int(quantizer.sym), int(block_size),
                               (_lib.TQ_LOOP_TRITON if use_triton else _lib.TQ_LOOP_TORCH)
                               | (_lib.TQ_LOOP_STRICT_FP32 if strict_fp32 else 0), _ptr(out), out.stride(0),
                               _ptr(codes), n, _ptr(ws), ws.numel(), _stream(W32)), "tq_gptq_loop")
    rel = None
    if R_x is not None:
        rel = _relative_error(W32, out, R_x.to(dev), p64)
        logging.info(f"   [Metric] Relative prediction error: {rel:.6f}")
    return QuantizedLinear(final_W=out.to(dtype=orig_dtype), codes=codes, scale=quantizer.scale.squeeze(-1),
                           zero=quantizer.zero.squeeze(-1), rank=k, min_q=quantizer.min_q, bits=quantizer.w_bits,
                           rel_error=rel)


def gptq_fwrd(weight_mat: torch.Tensor, H_inv_sqrt: torch.Tensor, quantizer: Quantizer, perm: torch.Tensor,
              block_size: int = 128, use_triton: bool = True,
              R_x: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
    """Blocked GPTQ loop; returns (dequantised W in the input dtype, rank) (gptq_utils.py:459-565).

    `block_size` keeps the reference's numerical meaning (which (c, j) pairs are scaled by
    reciprocal-multiply vs division); `use_triton` selects the Triton-kernel arithmetic
    (half-up, :345-386) or the torch-loop arithmetic (half-even, :516-534) - both run in
    the same CUDA kernels.
    """
    q = gptq_quantize(weight_mat, H_inv_sqrt, quantizer, perm, block_size, use_triton, R_x, want_codes=False)
    return q.final_W, q.rank


def pack_codes(codes: torch.Tensor, bits: int) -> torch.Tensor:
    """Pack m x n uint8 biased codes LSB-first along n into uint32 words (int32 storage)."""
    _require_cuda(codes, "pack_codes")
    lib = _lib.load()
    m, n = codes.shape
    codes = codes.contiguous()
    nwords = (n * bits + 31) // 32
    out = torch.empty((m, nwords), dtype=torch.int32, device=codes.device)
    with torch.cuda.device(codes.device):
        check(lib.tq_pack_codes(_ptr(codes), codes.stride(0), m, n, bits, _ptr(out), nwords, _stream(codes)),
              "tq_pack_codes")
    return out


def export_gptq(q: QuantizedLinear, group_size: int, scales_dtype: torch.dtype = torch.float16, v1_zero_offset: bool = True):
    """Checkpoint tensors of one Linear in the GPTQ / AutoGPTQ layout that vLLM's `gptq` loader reads (the
    packed-weight output the reference lists as a roadmap item, README.md:133):

        qweight  int32 [in_features * bits / 32, out_features]   codes packed along the INPUT dimension
        qzeros   int32 [n_groups, out_features * bits / 32]       zero points packed along the OUTPUT dimension
        scales   fp16  [n_groups, out_features]
        g_idx    int32 [in_features] = i // group_size            sequential groups (static groups: README.md:43)

    Stored values are unsigned: code' = code - min_q, zero' = zero - min_q (symmetric grids: min_q = -(2^(b-1) - 1),
    zero = 0), so W[j, i] = scales[g, j] * (code'[i, j] - zero'[g, j]).  `v1_zero_offset=True` stores zero' - 1
    (modulo 2^bits) as the v1 format does; readers add the 1 back.  2 / 4 / 8 bits put 32 / bits values into a
    word, 3 bits 32 values into 3 words (AutoGPTQ's scheme = the LSB-first bitstream)."""
    _require_cuda(q.codes, "export_gptq")
    lib = _lib.load()
    m, n = q.codes.shape
    bits = q.bits
    g = group_size if group_size > 0 else n
    if n % g != 0 or (n * bits) % 32 != 0 or (m * bits) % 32 != 0:
        raise ValueError(f"export_gptq: in_features {n} / out_features {m} do not pack into whole 32-bit words at {bits} bits")
    ng = n // g
    dev = q.codes.device
    codes = q.codes.contiguous()
    qweight = torch.empty((n * bits // 32, m), dtype=torch.int32, device=dev)
    zeros_u8 = (q.zero.to(torch.float32) - float(q.min_q)).round().to(torch.uint8).t().contiguous()      # [ng, m]
    qz_t = torch.empty((m * bits // 32, ng), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.tq_pack_gptq(_ptr(codes), n, m, n, bits, 0, _ptr(qweight), m, _stream(codes)), "tq_pack_gptq")
        check(lib.tq_pack_gptq(_ptr(zeros_u8), m, ng, m, bits, -1 if v1_zero_offset else 0, _ptr(qz_t),
                               ng, _stream(codes)), "tq_pack_gptq")
    return {"qweight": qweight, "qzeros": qz_t.t().contiguous(), "scales": q.scale.t().contiguous().to(scales_dtype),
            "g_idx": (torch.arange(n, device=dev, dtype=torch.int32) // g), "bits": bits, "group_size": g,
            "sym": q.min_q < 0, "zero_offset": 1 if v1_zero_offset else 0}
