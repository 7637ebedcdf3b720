"""
Round-2 preparation (CPU prototype, numpy): two-stage tridiagonal reduction.

Why.  The one-stage reduction streams the lower triangle of the trailing matrix once per COLUMN
(4 n^3 / 3 bytes; sytrd_panel_sym_kernel already runs that stream at ~0.93 of the HBM peak, so the
0.82 s it costs at n = 12288 can only shrink by moving fewer bytes).  Reducing first to a BAND of width
b (stage 1, all BLAS-3: the trailing matrix is read once per b columns) and then chasing the band down
to tridiagonal form (stage 2, O(n^2 b) flops on an L2-resident (b+1) x n array) removes the BLAS-2 half.
Price: the eigenvectors need two back-transformations, Z = Q1 (Q2 Z_T), 4 n^3 flops instead of 2 n^3.
Estimate for n = 12288 on a B200 (36 TF/s DGEMM): stage 1 ~0.17 s, stage 2 ~0.10-0.15 s (pipelined
sweeps, ~8 us per sweep start), Q2 ~0.25 s (grouped reflectors), Q1 ~0.13 s  =>  eigh ~0.75 s vs 1.04 s.

This file pins the algorithm (index ranges, reflector bookkeeping, order of the back-transformation) on
the CPU so that the CUDA version has an oracle:
    band, Q1 = sy2sb(A, b)            A = Q1 band Q1^T,   band has bandwidth b
    d, e, refl = sb2st(band, b)       band = Q2 T Q2^T,   refl = [(row0, v, tau), ...] in generation order
    Z = apply_q2(refl, Z_T)           Q2 Z_T
Run:  python scripts/prototypes/two_stage_tridiag.py   (self-check against numpy.linalg.eigh)
"""
import numpy as np


def house(x):
    """LAPACK dlarfg: (v, tau, beta) with v[0] = 1 and (I - tau v v^T) x = beta e1."""
    alpha, xn = x[0], np.linalg.norm(x[1:])
    if xn == 0.0:
        return np.concatenate([[1.0], np.zeros(len(x) - 1)]), 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, (beta - alpha) / beta, beta


def sy2sb(A, b):
    """Stage 1: A (symmetric) -> band of width b by QR factorizations of the sub-diagonal block columns and
    two-sided block updates.  Returns the band matrix (full storage) and Q1 (explicit, for the prototype;
    the CUDA version keeps compact-WY panels and applies them like ormtr)."""
    A = A.copy()
    n = A.shape[0]
    Q1 = np.eye(n)
    for j in range(0, n - b - 1, b):
        r0 = j + b                                   # rows below the band in block column [j, j+b)
        P = A[r0:, j:j + b]
        Q, R = np.linalg.qr(P, mode="complete")      # CUDA: cluster QR panel + compact WY
        A[r0:, j:j + b] = Q.T @ P                    # = [R; 0]
        A[j:j + b, r0:] = A[r0:, j:j + b].T
        A[r0:, r0:] = Q.T @ A[r0:, r0:] @ Q          # CUDA: W = A22 V T, rank-2b update (DGEMM / DSYR2K)
        Q1[:, r0:] = Q1[:, r0:] @ Q
    return A, Q1


def sb2st(B, b):
    """Stage 2: bulge chasing (Schwarz / Bischof-Lang-Sun).  Sweep s annihilates column s below its first
    sub-diagonal; the fill-in (bulge) it creates one block further down is chased off the matrix with one
    more reflector per block.  Works on full storage here; the CUDA kernel holds the (b+1) x n band in L2 and
    one task (reflector + the <= 3b x b window it touches) in shared memory.  Sweep s+1 may start task k as
    soon as sweep s has finished task k+2, i.e. sweeps are released 3 task-times apart (sb2st_wavefront below
    checks it: a distance of 1 or 2 changes the result, 3 does not) - the pipeline a persistent kernel runs."""
    B = B.copy()
    n = B.shape[0]
    refl = []
    for s in range(n - 2):
        c = s                                        # column whose sub-band part is annihilated
        r0 = s + 1
        k = 0                                        # chase index: reflector (s, k) acts on rows s+1+k b ...
        while r0 < n - 1:
            r1 = min(r0 + b, n)                      # reflector acts on rows [r0, r1)
            if r1 - r0 < 2:
                break
            v, tau, beta = house(B[r0:r1, c].copy())
            if tau != 0.0:
                H = np.eye(r1 - r0) - tau * np.outer(v, v)
                lo, hi = max(0, r0 - b), min(n, r1 + b)          # band window the reflector touches
                B[r0:r1, lo:hi] = H @ B[r0:r1, lo:hi]
                B[lo:hi, r0:r1] = B[lo:hi, r0:r1] @ H
            refl.append((r0, v, tau, s, k))
            # the bulge appears in block (rows [r1, r1+b), columns [r0, r1)); its first column is column r0
            c = r0
            r0 = r1
            k += 1
    d = np.diag(B).copy()
    e = np.diag(B, -1).copy()
    off = B - np.diag(d) - np.diag(e, -1) - np.diag(e, 1)
    return d, e, refl, float(np.abs(off).max())


def sb2st_wavefront(B, b, lag, rng=None):
    """The same chase, executed as the persistent CUDA kernel will run it: task (s, k) is released at time
    lag * s + k; tasks with the same time stamp run concurrently (here: in a shuffled order).  If `lag` is a
    valid pipeline distance the result equals the sweep-by-sweep order exactly."""
    B = B.copy()
    n = B.shape[0]
    tasks = []
    for s in range(n - 2):
        r0, k = s + 1, 0
        while r0 < n - 1 and min(r0 + b, n) - r0 >= 2:
            tasks.append((lag * s + k, s, k))
            r0, k = min(r0 + b, n), k + 1
    if rng is not None:
        tasks = [tasks[i] for i in rng.permutation(len(tasks))]
    tasks.sort(key=lambda t: t[0])                   # stable: equal stamps keep the shuffled order
    for _, s, k in tasks:
        r0 = s + 1 + k * b
        r1 = min(r0 + b, n)
        c = s if k == 0 else r0 - b
        v, tau, beta = house(B[r0:r1, c].copy())
        if tau != 0.0:
            H = np.eye(r1 - r0) - tau * np.outer(v, v)
            lo, hi = max(0, r0 - b), min(n, r1 + b)
            B[r0:r1, lo:hi] = H @ B[r0:r1, lo:hi]
            B[lo:hi, r0:r1] = B[lo:hi, r0:r1] @ H
    return np.diag(B).copy(), np.diag(B, -1).copy()


def apply_q2(refl, Z):
    """Z <- Q2 Z with Q2 = H_1 H_2 ... H_m in generation order (band = Q2 T Q2^T): apply in REVERSE order.
    CUDA: reflectors of consecutive sweeps at the same chase position overlap by one row - groups of 64 sweeps
    form a (b + 64) x 64 parallelogram that is applied as one compact-WY block (GEMM-shaped)."""
    Z = Z.copy()
    for r0, v, tau, _, _ in reversed(refl):
        if tau != 0.0:
            Z[r0:r0 + len(v), :] -= tau * np.outer(v, v @ Z[r0:r0 + len(v), :])
    return Z


def apply_q2_grouped(refl, Z, b, nb):
    """The GEMM-shaped version the CUDA kernel needs.  Reflector (s, k) acts on rows [s+1+kb, s+1+(k+1)b); two
    reflectors overlap iff |(s-s') + (k-k') b| < b, and of two overlapping ones the later generated (larger
    (s, k) lexicographically) must be applied first.  For blocks of nb <= b consecutive sweeps that order is
    kept by:  sweep blocks DESCENDING, inside a block chase index k ASCENDING, inside (block, k) sweeps
    descending - and the nb reflectors of one (block, k) form a (b + nb - 1) x nb staircase V that is applied
    as ONE compact-WY block  Z[rows] -= V (T (V^T Z[rows]))  (three GEMMs with K = nb)."""
    assert nb <= b
    Z = Z.copy()
    by = {}
    for r0, v, tau, s, k in refl:
        by[(s // nb, k)] = by.get((s // nb, k), []) + [(s, r0, v, tau)]
    blocks = sorted({sb for sb, _ in by}, reverse=True)
    for sb in blocks:
        for k in sorted(kk for (s2, kk) in by if s2 == sb):
            grp = sorted(by[(sb, k)])                 # sweeps ascending: G = H_s0 H_s0+1 ... (apply last one first)
            rlo = grp[0][1]
            rhi = max(r0 + len(v) for _, r0, v, _ in grp)
            V = np.zeros((rhi - rlo, len(grp)))
            taus = np.zeros(len(grp))
            for j, (_, r0, v, tau) in enumerate(grp):
                V[r0 - rlo:r0 - rlo + len(v), j] = v
                taus[j] = tau
            T = np.zeros((len(grp), len(grp)))        # forward columnwise larft: H_0 H_1 ... H_{m-1} = I - V T V^T
            for j in range(len(grp)):
                T[j, j] = taus[j]
                if j:
                    T[:j, j] = -taus[j] * (T[:j, :j] @ (V[:, :j].T @ V[:, j]))
            Z[rlo:rhi, :] -= V @ (T @ (V.T @ Z[rlo:rhi, :]))
    return Z


def eigh_two_stage(A, b):
    band, Q1 = sy2sb(A, b)
    d, e, refl, off = sb2st(band, b)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, ZT = np.linalg.eigh(T)                         # CUDA: the existing divide & conquer (stedc)
    Z2 = apply_q2(refl, ZT)
    Z2g = apply_q2_grouped(refl, ZT, b, max(1, b // 2))
    assert np.abs(Z2 - Z2g).max() <= 1e-12, np.abs(Z2 - Z2g).max()      # the grouped order is the same operator
    Z = Q1 @ Z2g
    return w, Z, off, len(refl)


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    for n, b in ((40, 4), (97, 8), (200, 16), (257, 32)):
        M = rng.standard_normal((n, n)) * np.logspace(0, -3, n)[None, :]
        A = M @ M.T
        band, Q1 = sy2sb(A, b)
        bw = max((abs(i - j) for i in range(n) for j in range(n) if abs(band[i, j]) > 1e-13 * np.abs(A).max()), default=0)
        w, Z, off, nref = eigh_two_stage(A, b)
        wr = np.linalg.eigvalsh(A)
        res = np.linalg.norm(A @ Z - Z * w[None, :]) / np.linalg.norm(A)
        orth = np.linalg.norm(Z.T @ Z - np.eye(n))
        d0, e0, _, _ = sb2st(band, b)
        ok = {}
        for lag in (1, 2, 3):
            d1, e1 = sb2st_wavefront(band, b, lag, rng)
            ok[lag] = bool(np.allclose(d0, d1, rtol=0, atol=1e-12 * np.abs(A).max()) and
                           np.allclose(np.abs(e0), np.abs(e1), rtol=0, atol=1e-12 * np.abs(A).max()))
        print(f"   pipeline distance between consecutive sweeps (tasks): {ok}")
        assert ok[3]
        print(f"n={n} b={b}: bandwidth after stage 1 = {bw}, off-tridiagonal after stage 2 = {off:.1e}, "
              f"{nref} reflectors (n^2/2b = {n * n // (2 * b)}), max|dw|/|w|max = {np.abs(w - wr).max() / wr.max():.1e}, "
              f"residual {res:.1e}, orthogonality {orth:.1e}")
        assert bw <= b and off < 1e-12 * np.abs(A).max() and res < 1e-12 and orth < 1e-11
    print("ok")
