"""ctypes binding of libtruncgptq.so (the C ABI declared in include/truncgptq.h).

The library is built in-tree (gptq_svd_b200/libtruncgptq.so) by
`__graft_entry__.build()` / `make -C gptq_svd_b200/csrc`.  There is no fallback:
if the shared object is missing, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtruncgptq.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "truncgptq.h")

TQ_F16, TQ_BF16, TQ_F32, TQ_F64 = 0, 1, 2, 3
TQ_RANK_ENERGY, TQ_RANK_MEAN_TRIMMED, TQ_RANK_FULL = 0, 1, 2
TQ_SOLVE_HOUSEHOLDER_QRCP = 0x100
TQ_LOOP_TRITON, TQ_LOOP_TORCH = 0, 1
TQ_LOOP_STRICT_FP32 = 0x100

_i64, _i32, _vp, _sz, _dbl = C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_double

PROF_KINDS = ["sytrd_panel_sym_kernel", "sytrd_panel_kernel", "qrcp_panel_kernel", "sb2st_chase_kernel",
              "pchol_panel_kernel", "qr_cluster_panel_kernel", "gptq_block_kernel", "trailing_tc_kernel",
              "trailing_update_kernel", "syrk_tcgen05_kernel", "metric_tc_kernel"]
PROF_UNIT = ["B", "B", "B", "B", "B", "B", "B", "flop", "flop", "flop", "flop"]

STAGE_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_void_p)
TQ_STAGE_SYTRD_DONE = 1
TQ_STAGE_BAND_DONE = 2

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "tq_version": [],
    "tq_launch_count": [],
    "tq_set_sm_budget": [_i32],
    "tq_set_eigh_two_stage": [_i32],
    "tq_two_stage_debug": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp],
    "tq_set_stage_callback": [STAGE_CALLBACK, _vp],
    "tq_profile_begin": [_i32],
    "tq_profile_kernel": [_i32, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_dbl),
                          C.POINTER(_dbl)],
    "tq_profile_end": [C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_i64), C.POINTER(_i64)],
    "tq_last_error": [],
    "tq_syrk_accum": [_vp, _i64, _vp, _i32, _i64, _i64, _i64, _i32, _vp],
    "tq_syrk_accum_scaled": [_vp, _i64, _vp, _i32, _i64, _i64, _i64, _i32, _vp, _vp],
    "tq_cast_to_f16_scaled": [_vp, _i32, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp],
    "tq_hessian_probe_accum": [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp],
    "tq_hessian_probe_check": [_vp, _i64, _i64, _vp, _dbl, _vp, _vp, _vp],
    "tq_hessian_scale": [_vp, _i64, _i64, _i64, _vp, _i64, _vp],
    "tq_cast_to_f16": [_vp, _i32, _i64, _i64, _i64, _vp, _i64, _vp],
    "tq_solver_workspace": [_i64, C.POINTER(_sz)],
    "tq_spectral_solve": [_vp, _i64, _i64, _dbl, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _sz, _vp],
    "tq_eigh": [_vp, _i64, _i64, _vp, _vp, _i64, _vp, _sz, _vp],
    "tq_rank_select": [_vp, _i64, _dbl, _i32, _vp, C.POINTER(_i64), _vp, _sz, _vp],
    "tq_qrcp": [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _sz, _vp],
    "tq_qr_r": [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _sz, _vp],
    "tq_find_params": [_vp, _i64, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp],
    "tq_gptq_loop_workspace": [_i64, _i64, _i64, C.POINTER(_sz)],
    "tq_gptq_loop": [_vp, _i64, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32,
                     _i32, _vp, _i64, _vp, _i64, _vp, _sz, _vp],
    "tq_pack_codes": [_vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp],
    "tq_pack_gptq": [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _i64, _vp],
    "tq_quant_error_workspace": [_i64, _i64, _i64, C.POINTER(_sz)],
    "tq_cholesky_workspace": [_i64, C.POINTER(_sz)],
    "tq_cholesky_solve": [_vp, _i64, _i64, _vp, _dbl, _vp, _i64, C.POINTER(_i32), _vp, _sz, _vp],
    "tq_sketch_accum_workspace": [_i64, _i64, _i64, C.POINTER(_sz)],
    "tq_sketch_accum": [_vp, _i64, _vp, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _sz, _vp],
    "tq_sketch_workspace": [_i64, _i64, C.POINTER(_sz)],
    "tq_sketch_solve": [_vp, _i64, _i64, _i64, _dbl, _i32, _vp, _vp, C.POINTER(_i64), _vp, _sz, _vp],
    "tq_quant_error": [_vp, _i64, _vp, _i64, _vp, _i32, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _sz, _vp],
}
_RESTYPES = {"tq_last_error": C.c_char_p, "tq_launch_count": C.c_int64}

_lib = None


def header_symbols():
    """Every function name declared in include/truncgptq.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tq_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load libtruncgptq.so; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the TruncGPTQ hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


class TruncGPTQError(RuntimeError):
    pass


def check(status: int, what: str):
    if status != 0:
        msg = load().tq_last_error()
        raise TruncGPTQError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
