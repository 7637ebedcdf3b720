"""CPU: the C-ABI library loads, exports every symbol declared in include/truncgptq.h,
the ctypes table covers the header, and compute entry points fail loudly without a GPU
(no CPU fallback anywhere on the product path)."""
import ctypes as C
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from gptq_svd_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        entry.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libtruncgptq.so does not export {s}"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.tq_version() == 100


def test_header_cites_reference_lines():
    src = open(_lib.HEADER_PATH).read()
    assert len(re.findall(r"gptq_utils\.py:\d+", src)) >= 15
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)     # no torch types in signatures


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_calls_fail_loudly_without_gpu(lib):
    st = lib.tq_find_params(None, 0, 0, 0, 4, 128, 0, None, None, None)
    assert st == -2 and b"no CPU fallback" in lib.tq_last_error()
    nbytes = C.c_size_t(0)
    assert lib.tq_solver_workspace(4096, C.byref(nbytes)) == 0 and nbytes.value > 4 * 4096 * 4096 * 8
    assert lib.tq_gptq_loop_workspace(128, 256, 64, C.byref(nbytes)) == 0 and nbytes.value > 128 * 256 * 4
    import gptq_svd_b200 as G
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.HessianAccumulator(64, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.process_hessian_alt(torch.eye(8, dtype=torch.float64), 1e-4, "energy")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.gptq_fwrd(torch.zeros(4, 8), torch.eye(8), G.Quantizer(4, -1, False), torch.arange(8))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(os.path.dirname(_lib.LIB_PATH))
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_quantizer_api_mirrors_reference():
    import gptq_svd_b200 as G
    q = G.Quantizer(4, 128, True)
    assert (q.max_q, q.min_q) == (7, -7)                # gptq_utils.py:239-242: 15 levels at 4-bit sym
    q = G.Quantizer(2, 128, True)
    assert (q.max_q, q.min_q) == (1, -1)
    q = G.Quantizer(3, -1, False)
    assert (q.max_q, q.min_q) == (7, 0) and q.scale is None and q.zero is None
    import inspect
    sig = inspect.signature(G.gptq_fwrd)
    assert list(sig.parameters) == ["weight_mat", "H_inv_sqrt", "quantizer", "perm", "block_size", "use_triton", "R_x"]
    assert sig.parameters["block_size"].default == 128 and sig.parameters["use_triton"].default is True
    sig = inspect.signature(G.process_hessian_alt)
    assert sig.parameters["threshold"].default == 0.0005 and sig.parameters["threshold_method"].default == "mean_trimmed"


def test_frontend_entry_points_fail_loudly_and_mirror_reference(lib):
    """Rows f3 / f4: the new compute entry points refuse to run without a B200, the workspace queries
    work on the host, and the Python signatures are the reference's (gptq_utils.py:33-36,129-133,176)."""
    import inspect
    import gptq_svd_b200 as G
    e = C.c_int(0)
    assert lib.tq_cholesky_solve(None, 0, 0, None, 0.01, None, 0, C.byref(e), None, 0, None) == -2
    assert b"no CPU fallback" in lib.tq_last_error()
    assert lib.tq_sketch_accum(None, 0, None, 0, None, 0, 0, 0, 0, 0, None, 0, None) == -2
    k = C.c_int64(0)
    assert lib.tq_sketch_solve(None, 0, 0, 0, 1e-2, 1, None, None, C.byref(k), None, 0, None) == -2
    nbytes = C.c_size_t(0)
    assert lib.tq_cholesky_workspace(1024, C.byref(nbytes)) == 0 and nbytes.value >= 3 * 1024 * 1024 * 8
    assert lib.tq_sketch_workspace(512, 1024, C.byref(nbytes)) == 0 and nbytes.value >= 6 * 1024 * 1024 * 8
    assert lib.tq_set_sm_budget(37) == 0 and lib.tq_set_sm_budget(0) == 0
    assert lib.tq_set_stage_callback(_lib.STAGE_CALLBACK(0), None) == 0
    sig = inspect.signature(G.process_hessian)
    assert list(sig.parameters) == ["H", "actorder", "damp_percent"]
    assert sig.parameters["actorder"].default is False and sig.parameters["damp_percent"].default == 0.01
    sig = inspect.signature(G.process_sketch)
    assert list(sig.parameters) == ["sketch", "threshold", "threshold_method"]
    assert sig.parameters["threshold"].default == 1e-2 and sig.parameters["threshold_method"].default == "mean_trimmed"
    assert list(inspect.signature(G.Sketcher.__init__).parameters) == ["self", "layer", "rank", "device"]
    assert hasattr(G.Sketcher, "hook_fn") and hasattr(G.Sketcher, "get_scaled_sketch")


def test_committed_bench_line_keeps_the_contract():
    """The bench line committed under profiles/ (the last run of the round) carries every key the driver reads:
    the base contract, the tier's `roofline` / `cpu_baseline` objects, `e2e`, `gpu_launches`, `clocks`."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_bench_final_n1.json")
    j = json.loads(open(path).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["higher_is_better"] is False and j["n_gpus"] == 1 and j["gpu_launches"] > 0
    assert "workload" in j["config"] and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"])
    e = j["e2e"]
    assert e["value"] >= j["value"] * 0.95 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(j["clocks"]["reasons"]))
