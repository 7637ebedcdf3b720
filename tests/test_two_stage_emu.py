"""CPU checks of the experimental two-stage tridiagonal reduction (gptq_svd_b200/csrc/two_stage.cu):

* the numpy model of its data layout (scripts/prototypes/sb2st_band.py) against numpy.linalg.eigh and against the
  full-storage prototype;
* the KERNEL SOURCE of two_stage_kernels.cuh compiled for the host (tests/emu/two_stage_emu.cpp: one OS thread per
  CUDA thread, CTAs running concurrently) against that model - band extraction, the persistent bulge-chase kernel
  with its acquire / release progress counters, the staircase copies of the Q2 back-transformation;
* the WHOLE two-stage path - host driver of two_stage.cu compiled unchanged with g++, its kernels on the emulation
  runtime, cuBLAS / CUDA runtime entry points replaced by reference loops (tests/emu/two_stage_host_emu.cpp) -
  against numpy.linalg.eigh;
* the wavefront schedule of the Q2 back-transformation at the Qwen3-8B sizes.
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from scripts.prototypes import sb2st_band as M
from scripts.prototypes import two_stage_tridiag as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 64
# The emulation runs one OS thread per CUDA thread: a reduction at n = 256 costs ~15 s.  The default run keeps one
# case of everything; TQ_TEST_EMU_FULL=1 adds the larger / redundant ones (all of them passed when they were written).
FULL = os.environ.get("TQ_TEST_EMU_FULL") == "1"
slow = pytest.mark.skipif(not FULL, reason="larger emulation case: set TQ_TEST_EMU_FULL=1")


def _spd(n, seed):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n, n)) * np.logspace(0, -2, n)[None, :]
    return X @ X.T


@pytest.mark.parametrize("n,b", [(64, 8), (96, 16), (128, 32)])
def test_band_model_is_an_eigendecomposition(n, b):
    A = _spd(n, n)
    w, Z, d, e = M.eigh_two_stage_model(A, b, ncta=5, rng=np.random.RandomState(3))
    assert np.abs(w - np.linalg.eigvalsh(A)).max() <= 1e-13 * w.max()
    assert np.linalg.norm(A @ Z - Z * w) <= 1e-13 * np.linalg.norm(A) * n
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= 1e-12


def test_band_model_matches_full_storage_prototype():
    n, b = 96, 16
    A = _spd(n, 7)
    band, _ = P.sy2sb(A, b)
    d0, e0, refl, off = P.sb2st(band, b)
    Bd, ldb = M.extract_band(band, b)
    d1, e1, Vs, tau2, _ = M.sb2st_band(Bd, ldb, n, b, ncta=6, rng=np.random.RandomState(1))
    sc = np.abs(A).max()
    assert np.abs(d0 - d1).max() <= 1e-12 * sc and np.abs(e0 - e1).max() <= 1e-12 * sc
    for r0, v, tau, s, k in refl:                      # same reflectors, stored where the kernel stores them
        assert np.abs(Vs[r0 + s * n:r0 + len(v) + s * n] - v).max() <= 1e-9
        assert abs(tau2[s + k * n] - tau) <= 1e-9


def test_progress_protocol_distance_three_is_valid_and_minimal():
    """The chase kernel publishes prog[s] = k + 1 as soon as task k has written its G block (its D / E blocks are
    still in registers) and releases task (s + 1, k) at prog[s] >= k + 3.  With CTAs stepped half a task at a time
    in random order the result must not depend on the schedule - and a distance of 2 must break it (the test
    would otherwise prove nothing)."""
    n, b = 96, 8
    A = _spd(n, 21)
    band, _ = P.sy2sb(A, b)
    d0, e0, _, _ = P.sb2st(band, b)
    Bd, ldb = M.extract_band(band, b)
    sc = np.abs(A).max()
    broken = 0
    for trial in range(5):
        d, e, *_ = M.sb2st_band(Bd, ldb, n, b, ncta=3 + trial, rng=np.random.RandomState(trial), lag=3)
        assert np.abs(d - d0).max() <= 1e-12 * sc and np.abs(np.abs(e) - np.abs(e0)).max() <= 1e-12 * sc
        d, e, *_ = M.sb2st_band(Bd, ldb, n, b, ncta=3 + trial, rng=np.random.RandomState(trial), lag=2)
        broken += not (np.abs(d - d0).max() <= 1e-9 * sc and np.abs(np.abs(e) - np.abs(e0)).max() <= 1e-9 * sc)
    assert broken > 0


def test_late_load_protocol_is_valid_and_minimal():
    """TQ_CHASE_LATE: no wait in front of the reflector step of a task k >= 1, prog >= 2 in front of task 0,
    prog >= k + 3 in front of the D / E loads - valid under half-task interleavings, and neither wait can be weaker"""
    n, b = 96, 8
    A = _spd(n, 22)
    band, _ = P.sy2sb(A, b)
    d0, e0, _, _ = P.sb2st(band, b)
    Bd, ldb = M.extract_band(band, b)
    sc = np.abs(A).max()

    def same(lag, lag_a, trial):
        d, e, *_ = M.sb2st_band(Bd, ldb, n, b, ncta=3 + trial, rng=np.random.RandomState(trial), lag=lag,
                                late_loads=True, lag_a=lag_a)
        return np.abs(d - d0).max() <= 1e-10 * sc and np.abs(np.abs(e) - np.abs(e0)).max() <= 1e-10 * sc

    assert all(same(3, 2, t) for t in range(6))
    assert not all(same(3, 1, t) for t in range(6))
    assert not all(same(2, 2, t) for t in range(6))


@pytest.mark.parametrize("n", list(range(256, 1600, 64)) + [4096, 8192, 12288, 14336, 28672])
def test_q2_wavefronts_at_model_sizes(n):
    """What apply_q2 (two_stage.cu) relies on: the groups of a wavefront are consecutive sweep blocks, start 3 b
    rows apart, only the lowest one can be clipped, and a wavefront never exceeds the scratch it sizes."""
    waves = M.q2_groups(n, B)                          # asserts stride and clipping itself
    maxb = n // (3 * B) + 2
    seen = set()
    for w, grp in waves:
        assert len(grp) <= maxb
        sbs = [g[0] for g in grp]
        assert sbs == list(range(sbs[0], sbs[0] + len(sbs)))
        for sb, k, rlo, hg, m in grp:
            assert rlo + hg <= n and 1 <= m <= B
            seen.add((sb, k))
    nsweeps = n - 2
    want = {(s // B, k) for s in range(0, nsweeps, B) for k in range(M.num_tasks(s, n, B))}
    assert seen == want                                # every group exactly once


# --------------------------------------------------------------------------------------------- kernel emulation
def _build(name, extra=()):
    """g++ build of tests/emu/<name>.cpp -> tests/emu/build/<name>.so (rebuilt when a source is newer)"""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = os.path.join(ROOT, "tests", "emu", name + ".cpp")
    out_dir = os.path.join(ROOT, "tests", "emu", "build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, name + ".so")
    csrc = os.path.join(ROOT, "gptq_svd_b200", "csrc")
    deps = [src, os.path.join(ROOT, "tests", "emu", "emu_runtime.h")] + [
        os.path.join(csrc, f) for f in ("two_stage_kernels.cuh", "two_stage.cu", "solver_kernels.cuh", "common.cuh")]
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        # -Bsymbolic: the emulation's own cudaMemsetAsync / cublasDgemm_v2 ... must win over a libcudart / libcublas
        # that another test of the same process has already loaded
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wl,-Bsymbolic", *extra,
                        "-o", so, src], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="module")
def emu():
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not available")
    lib = _build("two_stage_emu", extra=("-I" + cuda_inc,))
    dp = np.ctypeslib.ndpointer(np.float64, flags="C")
    ip = np.ctypeslib.ndpointer(np.int32, flags="C")
    lib.emu_constants.argtypes = [ip]
    lib.emu_band_extract.argtypes = [dp, C.c_int64, C.c_int, dp]
    lib.emu_band_diag.argtypes = [dp, C.c_int, dp, dp]
    lib.emu_chase.argtypes = [dp, C.c_int, dp, C.c_int64, dp, ip, C.c_int, C.c_int]
    lib.emu_copy_staircase.argtypes = [dp, C.c_int64, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
    k = np.zeros(8, np.int32)
    lib.emu_constants(k)
    assert list(k[:4]) == [B, 2 * B, 2 * B - 1, 2 * B]
    return lib


def _emu_reduce(emu, Astore, n, grid, helper=0):
    """band extraction + bulge chase + (d, e) through the emulated kernels; Astore: n x n, lower triangle valid"""
    Acm = np.ascontiguousarray(Astore.T).reshape(-1)          # column-major
    Bd = np.full(n * 2 * B, np.nan)
    emu.emu_band_extract(Acm, n, n, Bd)
    Bd0 = Bd.copy()
    Vs = np.zeros(n * n)
    tau2 = np.zeros(n * (n // B + 2))
    prog = np.zeros(n, np.int32)
    emu.emu_chase(Bd, n, Vs, n, tau2, prog, grid, helper)
    assert np.all(prog[:n - 2] == (1 << 30))
    d, e = np.zeros(n), np.zeros(n)
    emu.emu_band_diag(Bd, n, d, e)
    return Bd0, Bd, Vs, tau2, d, e[:n - 1]


@pytest.mark.parametrize("n,grid,helper", [(136, 2, 0), (136, 3, 1), (136, 3, 2), (136, 4, 3),
                                           pytest.param(256, 3, 0, marks=slow), pytest.param(200, 4, 3, marks=slow)])
def test_emulated_chase_kernel_matches_model(emu, n, grid, helper):
    """helper bit 0: the ninth warp owns the progress counters (TQ_CHASE_HELPER=1); bit 1: second wait in front of the
    D / E loads (TQ_CHASE_LATE=1)"""
    A = _spd(n, 100 + n)
    band = np.where(np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= B, A, 0.0)
    Bd_model, ldb = M.extract_band(band, B)
    Bd0, Bd, Vs, tau2, d, e = _emu_reduce(emu, np.tril(band), n, grid, helper)
    assert np.array_equal(Bd0, Bd_model)
    d1, e1, Vs1, tau21, Bd1 = M.sb2st_band(Bd_model, ldb, n, B, ncta=4, rng=np.random.RandomState(5))
    sc = np.abs(A).max()
    assert np.abs(Bd - Bd1).max() <= 1e-10 * sc               # whole band array, bulge room included
    assert np.abs(d - d1).max() <= 1e-10 * sc and np.abs(e - e1).max() <= 1e-10 * sc
    assert np.abs(Vs - Vs1).max() <= 1e-8 and np.abs(tau2 - tau21).max() <= 1e-8
    # and, independently of the model: T has the band matrix's eigenvalues, nothing is left off the tridiagonal
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(band)).max() <= 1e-13 * sc * n
    full = M.band_to_full(Bd, 2 * B, n)
    assert np.abs(full - T).max() <= 1e-13 * sc * n


def test_emulated_kernels_give_an_eigendecomposition(emu):
    """stage 1 as the host driver runs it (numpy), stages 2 and the Q2 staircase copies through the emulated kernels"""
    n = 192
    A = _spd(n, 9)
    Ast, tau1 = M.sy2sb_wy(A, B)
    _, _, Vs, tau2, d, e = _emu_reduce(emu, Ast, n, 3)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, Z = np.linalg.eigh(T)
    for _, grp in M.q2_groups(n, B):
        sb0, k0 = grp[0][0], grp[0][1]
        cnt = len(grp)
        Vc = np.full(cnt * 2 * B * B, np.nan)
        taub = np.full(cnt * B, np.nan)
        emu.emu_copy_staircase(Vs, n, tau2, n, sb0, k0, cnt, Vc, taub)
        for i, (sb, k, rlo, hg, m) in enumerate(grp):
            V = Vc[i * 2 * B * B:(i + 1) * 2 * B * B].reshape(B, 2 * B).T      # (2b x b), ld 2b, column-major
            Vm, taum = M.staircase(Vs, n, tau2, n, B, sb, k)
            assert np.all(V[0] == 0.0) and np.array_equal(V[1:], Vm)           # shifted down by one row
            assert np.array_equal(taub[i * B:(i + 1) * B], taum)
            assert (rlo - 1) % B == 0                                          # aligned window start
            Vh = V[:hg + 1]
            Z[rlo - 1:rlo + hg] -= Vh @ (M.larft(V, taum) @ (Vh.T @ Z[rlo - 1:rlo + hg]))
    Z = M.apply_q1(Ast, tau1, B, Z)
    assert np.abs(w - np.linalg.eigvalsh(A)).max() <= 1e-13 * w.max() * n
    assert np.linalg.norm(A @ Z - Z * w) <= 1e-13 * np.linalg.norm(A) * n
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= 1e-11


# --------------------------------------------------------------------------------------------- whole path
def _load_host_emu():
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isfile(os.path.join(cuda_inc, "cublas_v2.h")):
        pytest.skip("CUDA headers not available")
    lib = _build("two_stage_host_emu", extra=("-I" + cuda_inc,))
    dp = np.ctypeslib.ndpointer(np.float64, flags="C")
    lib.emu_two_stage_reduce.argtypes = [dp, C.c_int64, dp, dp, C.c_int]
    lib.emu_two_stage_back.argtypes = [dp, C.c_int64, dp, C.c_int64]
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_q2_schedule.argtypes = [C.c_int64, np.ctypeslib.ndpointer(np.int32, flags="C"), C.c_int]
    lib.emu_two_stage_debug.argtypes = [dp, C.c_int64, dp, dp, dp, C.c_int]
    return lib


def _check_whole_path(lib, n, sms, ncols):
    """two_stage_reduce + two_stage_back exactly as tq_eigh calls them (eigh.cu), D&C replaced by numpy"""
    A = _spd(n, 31 + n)
    Acm = np.ascontiguousarray(A.T).reshape(-1).copy()
    d, e = np.zeros(n), np.zeros(n)
    assert lib.emu_two_stage_reduce(Acm, n, d, e, sms) == 0, lib.emu_last_error()
    T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    w, ZT = np.linalg.eigh(T)
    dw = np.abs(w - np.linalg.eigvalsh(A)).max() / w.max()
    assert dw <= 1e-13, f"eigenvalues off by {dw:.3e} (relative to the largest)"
    w, ZT = w[n - ncols:], ZT[:, n - ncols:]                   # back-transform only the leading eigenvectors
    Zcm = np.ascontiguousarray(ZT.T).reshape(-1).copy()
    assert lib.emu_two_stage_back(Acm, n, Zcm, ncols) == 0, lib.emu_last_error()
    Z = Zcm.reshape(ncols, n).T
    res = np.linalg.norm(A @ Z - Z * w) / np.linalg.norm(A)
    orth = np.linalg.norm(Z.T @ Z - np.eye(ncols))
    assert res <= 1e-13 * n ** 0.5 and orth <= 1e-12 * n ** 0.5, f"residual {res:.3e}, orthogonality {orth:.3e}"


@pytest.fixture(scope="module")
def host_emu():
    return _load_host_emu()


@pytest.mark.parametrize("n,sms,ncols", [(256, 3, 200), pytest.param(320, 2, 320, marks=slow)])
def test_whole_two_stage_path_on_the_host(host_emu, n, sms, ncols):
    _check_whole_path(host_emu, n, sms, ncols)


@pytest.mark.parametrize("n", [256, 448, 1024, 4096, 12288, 28672])
def test_host_q2_schedule_equals_model(host_emu, n):
    """The wavefront enumeration of apply_q2 (C++, two_stage.cu) issues exactly the model's groups, in an order
    that keeps every dependency: (sb, k) after (sb, k-1) and after (sb+1, k-2 .. k)."""
    cap = 1 << 16
    out = np.zeros(4 * cap, np.int32)
    cnt = host_emu.emu_q2_schedule(n, out, cap)
    assert 0 < cnt <= cap, host_emu.emu_last_error()
    batches = out[:4 * cnt].reshape(cnt, 4)
    order = {}
    for bi, (sb0, k0, count, hg) in enumerate(batches):
        assert 1 <= count <= n // (3 * B) + 2
        for i in range(count):
            sb, k = int(sb0 + i), int(k0 + 2 * i)
            rlo = sb * B + 1 + k * B
            assert (sb, k) not in order and hg == min(2 * B - 1, n - rlo)
            order[(sb, k)] = bi
    want = {(g[0], g[1]) for _, grp in M.q2_groups(n, B) for g in grp}
    assert set(order) == want
    for (sb, k), pos in order.items():
        for dep in ((sb, k - 1), (sb + 1, k - 2), (sb + 1, k - 1), (sb + 1, k)):
            if dep in order:
                assert order[dep] < pos, ((sb, k), dep)


def test_debug_entry_point_returns_band_and_tridiagonal(host_emu):
    """tq_two_stage_debug (the C ABI's stage-by-stage view, used by tests/test_gpu_two_stage.py): H, the band matrix
    after stage 1 and the tridiagonal matrix after stage 2 share their eigenvalues."""
    n = 256
    A = _spd(n, 77)
    band = np.full(n * 2 * B, np.nan)
    d, e = np.zeros(n), np.zeros(n)
    assert host_emu.emu_two_stage_debug(np.ascontiguousarray(A).reshape(-1), n, band, d, e, 3) == 0, \
        host_emu.emu_last_error()
    Bf = M.band_to_full(band, 2 * B, n)
    assert np.abs(np.tril(Bf, -(B + 1))).max() == 0.0
    wr = np.linalg.eigvalsh(A)
    assert np.abs(np.linalg.eigvalsh(Bf) - wr).max() <= 1e-13 * wr.max()
    T = np.diag(d) + np.diag(e[:n - 1], 1) + np.diag(e[:n - 1], -1)
    assert np.abs(np.linalg.eigvalsh(T) - wr).max() <= 1e-13 * wr.max()
