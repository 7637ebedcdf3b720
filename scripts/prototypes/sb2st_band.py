"""
CPU model (numpy) of the CUDA two-stage tridiagonal reduction in `csrc/two_stage.cu`, written with the SAME data
layout and index arithmetic as the kernels so that every offset can be checked here before the first GPU run
(the method that found the sytrd tile-indexing bugs in round 1, tests/test_sytrd_tiling.py):

  * band storage  Bd[(i - c) + c * ldb],  0 <= i - c < ldb = 2 b   (lower band of width b plus the bulge room);
    a dense block with rows R0.. and columns C0.. is the sub-array  base = (R0 - C0) + C0 * ldb,  ld = ldb - 1;
  * task (s, k) of the bulge chase: reflector rows [r0, r1), r0 = s + 1 + k b; blocks  G (bulge, b x b, arrives in
    shared memory from the previous task of the sweep), D (diagonal, lower triangle), E (below, becomes the next G);
  * progress counters: prog[s] = k + 1 is published as soon as task k of sweep s has written its G block back (all
    earlier tasks are then complete; D_k and E_k are already in registers, E travels in shared memory to the next
    task); sweep s + 1 may run task k once  prog[s] >= k + 3;
  * reflector store  Vs[r0 + i + s * ldv]  (column s = all reflectors of sweep s, stacked), tau2[s + k * n];
  * Q2 back-transformation in WAVEFRONTS  w = k + 2 (M - sb): groups (sweep block sb of nb = b sweeps, chase index
    k) with equal w sit 3 b rows apart in Z - one strided-batched DGEMM triple per wavefront.

`scripts/prototypes/two_stage_tridiag.py` (full storage, sweep by sweep) is the oracle for this file.
"""
import numpy as np


def house(x):
    alpha, xn = x[0], np.linalg.norm(x[1:])
    if xn == 0.0:
        v = np.zeros(len(x))
        v[0] = 1.0
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, (beta - alpha) / beta, beta


def extract_band(A, b):
    """band_extract_kernel: Bd[(i - c) + c ldb] = A[i, c] for 0 <= i - c <= b, zero elsewhere."""
    n = A.shape[0]
    ldb = 2 * b
    Bd = np.zeros(n * ldb)
    for c in range(n):
        for d in range(0, min(b, n - 1 - c) + 1):
            Bd[d + c * ldb] = A[c + d, c]
    return Bd, ldb


def band_to_full(Bd, ldb, n):
    B = np.zeros((n, n))
    for c in range(n):
        for d in range(min(ldb, n - c)):
            B[c + d, c] = Bd[d + c * ldb]
            B[c, c + d] = Bd[d + c * ldb]
    return B


def num_tasks(s, n, b):
    """tasks of sweep s: k = 0 .. K-1 with r0 = s + 1 + k b <= n - 2"""
    if s > n - 3:
        return 0
    return (n - 3 - s) // b + 1


class Blk:
    """view of a dense block inside the band array (what the kernel loads into shared memory)"""

    def __init__(self, Bd, ldb, R0, C0, nr, nc):
        self.Bd, self.base, self.ld, self.nr, self.nc = Bd, (R0 - C0) + C0 * ldb, ldb - 1, nr, nc
        li, lc = np.arange(nr)[:, None], np.arange(nc)[None, :]
        self.I = self.base + li + lc * self.ld                 # flat index of element (li, lc)
        self.off = (R0 - C0) + li - lc                          # row inside the band array: must be in [0, ldb)
        self.low = li >= lc
        self.ldb = ldb

    def _mask(self, lower_only):
        m = self.low if lower_only else np.ones_like(self.low)
        assert np.all((self.off[m] >= 0) & (self.off[m] < self.ldb)), "block leaves the band array"
        return m

    def load(self, lower_only=False):
        m = self._mask(lower_only)
        M = np.zeros((self.nr, self.nc))
        M[m] = self.Bd[self.I[m]]
        return M

    def store(self, M, lower_only=False):
        m = self._mask(lower_only)
        self.Bd[self.I[m]] = M[m]


def task_phase_a(Bd, ldb, n, b, s, k, carry, Vs, ldv, tau2, late_loads=False):
    """First half of a task as the CTA executes it: load D and E (registers), form the reflector from the first
    column of G (the E block of the previous task of this sweep, held in shared memory; column s for k = 0), apply it
    to the rest of G from the left and write G back.  After this the sweep's progress counter is published: the
    next sweep only ever needs the G block of a task, never its D or E block (see sb2st_band)."""
    r0 = s + 1 + k * b
    r1 = min(r0 + b, n)
    ln = r1 - r0
    assert ln >= 2
    hi = min(n, r1 + b)
    ne = hi - r1 if ln == b else 0
    Dv = Blk(Bd, ldb, r0, r0, ln, ln)
    Ev = Blk(Bd, ldb, r1, r0, ne, b) if ne > 0 else None
    D = E = None
    if not late_loads:                                        # variant 1: D and E loaded before the reflector step
        D = Dv.load(lower_only=True)
        E = Ev.load() if ne > 0 else None
    if k == 0:
        base = 1 + s * ldb                                  # column s, rows s+1..: contiguous
        x = Bd[base:base + ln].copy()
        v, tau, beta = house(x)
        Bd[base] = beta
        Bd[base + 1:base + ln] = 0.0
    else:
        G = carry                                            # ln x b, rows r0:r1, columns r0-b:r0
        assert G.shape == (ln, b)
        v, tau, beta = house(G[:, 0].copy())
        G[0, 0] = beta
        G[1:, 0] = 0.0
        if tau != 0.0:
            w = v @ G[:, 1:]
            G[:, 1:] -= tau * np.outer(v, w)
        Blk(Bd, ldb, r0, r0 - b, ln, b).store(G)
    Vs[r0 + s * ldv:r0 + ln + s * ldv] = v
    tau2[s + k * n] = tau
    return dict(v=v, tau=tau, D=D, Dv=Dv, E=E, Ev=Ev, ne=ne, r1=r1)


def task_phase_b(n, st):
    """Second half: two-sided update of D (lower triangle, written back) and right update of E, which stays in
    shared memory as the next task's G (or is written back when the sweep ends).  Uses the copies loaded in phase A."""
    v, tau, D, E = st["v"], st["tau"], st["D"], st["E"]
    if D is None:                                             # variant 2: loaded only now, after the second wait
        D = st["Dv"].load(lower_only=True)
        E = st["Ev"].load() if st["ne"] > 0 else None
    if tau != 0.0:
        Dfull = D + np.tril(D, -1).T
        p = tau * (Dfull @ v)
        w = p - 0.5 * tau * (p @ v) * v
        D -= np.tril(np.outer(v, w) + np.outer(w, v))
    st["Dv"].store(D, lower_only=True)
    if st["ne"] <= 0:
        return None
    if tau != 0.0:
        u = E @ v
        E -= tau * np.outer(u, v)
    if st["ne"] >= 2 and st["r1"] <= n - 2:
        return E                                             # next task exists: carried in shared memory
    st["Ev"].store(E)
    return None


def run_task(Bd, ldb, n, b, s, k, carry, Vs, ldv, tau2):
    return task_phase_b(n, task_phase_a(Bd, ldb, n, b, s, k, carry, Vs, ldv, tau2))


def sb2st_band(Bd, ldb, n, b, ncta=5, rng=None, lag=3, late_loads=False, lag_a=2):
    """The persistent kernel: CTA g owns sweeps g, g + ncta, ...  prog[s] = k + 1 is published when task k of
    sweep s has written its G block back (phase A) - tasks < k are then complete, D_k / E_k are still in
    registers - and a CTA may start task (s, k) when prog[s-1] >= k + lag or sweep s-1 is finished.  lag = 3 is the
    smallest valid distance: (s+1, k) touches the first entry of G_{k+2} of sweep s and nothing of D_{k+2} / E_{k+2}.
    CTAs are stepped in random order, HALF a task at a time, so that other sweeps do run between the two halves."""
    Bd = Bd.copy()
    ldv = n
    Vs = np.zeros(n * n)
    tau2 = np.zeros(n * (n // b + 2))
    BIG = 1 << 30
    prog = np.zeros(max(n, 1), dtype=np.int64)
    state = [{"s": g, "k": 0, "carry": None, "half": None} for g in range(ncta)]
    nsweeps = max(0, n - 2)
    rng = rng or np.random.RandomState(0)
    live = True
    while live:
        live = False
        for g in rng.permutation(ncta):
            stt = state[g]
            s = stt["s"]
            if s >= nsweeps:
                continue
            live = True
            K = num_tasks(s, n, b)
            k = stt["k"]
            if stt["half"] is None:
                if late_loads:
                    # variant 2: the reflector step of a task k >= 1 works on the block carried in shared memory and
                    # waits for nothing; task 0 reads column s from the band array and needs prog[s-1] >= lag_a
                    if s > 0 and k == 0 and prog[s - 1] < lag_a:
                        continue
                elif s > 0 and prog[s - 1] < k + lag:
                    continue                                 # spin
                stt["half"] = task_phase_a(Bd, ldb, n, b, s, k, stt["carry"], Vs, ldv, tau2, late_loads)
                if k + 1 < K:
                    prog[s] = k + 1                          # early publish: G_k is final
                continue
            if late_loads and s > 0 and prog[s - 1] < k + lag:
                continue                                     # variant 2: second wait, in front of the D / E loads
            stt["carry"] = task_phase_b(n, stt["half"])
            stt["half"] = None
            k += 1
            if k == K:
                assert stt["carry"] is None
                prog[s] = BIG
                stt["s"], stt["k"] = s + ncta, 0
            else:
                stt["k"] = k
    d = np.array([Bd[c * ldb] for c in range(n)])
    e = np.array([Bd[1 + c * ldb] for c in range(n - 1)])
    return d, e, Vs, tau2, Bd


def q2_groups(n, b):
    """Wavefront schedule of the Q2 back-transformation (nb = b).  Returns [(w, [(sb, k, rlo, hg, m), ...])] with the
    groups of a wavefront ordered by ascending rlo; rlo advances by 3 b from one to the next."""
    nb = b
    nsweeps = n - 2
    if nsweeps <= 0:
        return []
    M = (nsweeps - 1) // nb
    out = []
    kmax0 = num_tasks(0, n, b) - 1
    for w in range(0, kmax0 + 2 * M + 1):
        grp = []
        for sb in range(M + 1):
            k = w - 2 * (M - sb)
            s0 = sb * nb
            if k < 0 or k > num_tasks(s0, n, b) - 1:
                continue
            rlo = s0 + 1 + k * b
            m = max(0, min(s0 + nb, nsweeps, n - 2 - k * b) - s0)       # sweeps of the block that have a task k
            hg = min(b + nb - 1, n - rlo)
            grp.append((sb, k, rlo, hg, m))
        if grp:
            for a, c in zip(grp[:-1], grp[1:]):
                assert c[2] - a[2] == 3 * b                 # uniform stride: one strided-batched GEMM
                assert a[3] == b + nb - 1 and a[4] == nb    # only the last group of a wavefront can be clipped
            out.append((w, grp))
    return out


def staircase(Vs, ldv, tau2, n, b, sb, k):
    """copy_staircase_kernel: clean (b + nb - 1) x nb block of group (sb, k) and its tau vector."""
    nb = b
    s0 = sb * nb
    rlo = s0 + 1 + k * b
    H = b + nb - 1
    Vc = np.zeros((H, nb))
    tau = np.zeros(nb)
    for j in range(nb):
        s = s0 + j
        r0 = s + 1 + k * b
        if s > n - 3 or r0 > n - 2:
            continue
        ln = min(b, n - r0)
        for r in range(j, j + ln):
            Vc[r, j] = Vs[rlo + r + s * ldv]
        tau[j] = tau2[s + k * n]
    return Vc, tau


def larft(V, tau):
    m = len(tau)
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = tau[j]
        if j:
            T[:j, j] = -tau[j] * (T[:j, :j] @ (V[:, :j].T @ V[:, j]))
    return T


def apply_q2_wavefront(Vs, ldv, tau2, n, b, Z):
    Z = Z.copy()
    for _, grp in q2_groups(n, b):
        for sb, k, rlo, hg, m in grp:                        # independent: disjoint rows
            Vc, tau = staircase(Vs, ldv, tau2, n, b, sb, k)
            T = larft(Vc, tau)
            Vh = Vc[:hg]                                     # rows beyond n are structurally zero
            assert np.all(Vc[hg:] == 0.0)
            Z[rlo:rlo + hg] -= Vh @ (T @ (Vh.T @ Z[rlo:rlo + hg]))
    return Z



# ----------------------------------------------------------------------------------------------- stage 1 (sy2sb)
def qr_inplace(P):
    """LAPACK dgeqr2 layout (what qr_r_colmajor leaves): R on / above the diagonal, reflector tails below, tau."""
    P = P.copy()
    s, nb = P.shape
    tau = np.zeros(nb)
    for j in range(min(s, nb)):
        v, t, beta = house(P[j:, j].copy())
        tau[j] = t
        if t != 0.0 and j + 1 < nb:
            w = v @ P[j:, j + 1:]
            P[j:, j + 1:] -= t * np.outer(v, w)
        P[j, j] = beta
        P[j + 1:, j] = v[1:]
    return P, tau


def clean_reflectors(Astore, row0, col0, s, jb):
    """copy_reflectors_kernel: unit diagonal, zeros above, stored entries below; column t has its unit at row t."""
    Vc = np.zeros((s, jb))
    for t in range(jb):
        for r in range(s):
            Vc[r, t] = 0.0 if r < t else (1.0 if r == t else Astore[row0 + r, col0 + t])
    return Vc


def sy2sb_wy(A, b):
    """Stage 1 as the CUDA driver runs it (lower triangle only is referenced / updated):
         panel P = A[r0:, j:j+b]  -> QR in place, tau1[j:j+b]
         V = clean(P), T = larft(V, tau);  X = A22 V T (DSYMM, DGEMM);  M = T^T (V^T X);  W = X - V M / 2
         A22 -= V W^T + W V^T   (DSYR2K, lower)."""
    A = np.tril(A).copy()                                     # nothing above the diagonal may be read
    n = A.shape[0]
    assert n % b == 0 and n >= 2 * b
    tau1 = np.zeros(n)
    for j in range(0, n - b, b):
        r0 = j + b
        s = n - r0
        P, tau = qr_inplace(A[r0:, j:j + b])
        A[r0:, j:j + b] = P
        tau1[j:j + b] = tau
        V = clean_reflectors(A, r0, j, s, b)
        T = larft(V, tau)
        A22 = np.tril(A[r0:, r0:]) + np.tril(A[r0:, r0:], -1).T
        X = (A22 @ V) @ T
        Mm = T.T @ (V.T @ X)
        W = X - 0.5 * V @ Mm
        A[r0:, r0:] -= np.tril(V @ W.T + W @ V.T)
    return A, tau1


def apply_q1(Astore, tau1, b, Z, ob=None):
    """Z <- Q1 Z, ormtr style: blocks of `ob` reflector columns, last block first.  Reflector c has its unit at row
    c + b, so a block starting at column j0 is a clean staircase from row j0 + b."""
    n = Astore.shape[0]
    ob = ob or 2 * b
    Z = Z.copy()
    nref = n - b
    nblk = (nref + ob - 1) // ob
    for blk in range(nblk - 1, -1, -1):
        j0 = blk * ob
        jb = min(ob, nref - j0)
        s = n - j0 - b
        V = clean_reflectors(Astore, j0 + b, j0, s, jb)
        T = larft(V, tau1[j0:j0 + jb])
        Z[j0 + b:] -= V @ (T @ (V.T @ Z[j0 + b:]))
    return Z


def eigh_two_stage_model(A, b, ncta=7, rng=None):
    n = A.shape[0]
    Ast, tau1 = sy2sb_wy(A, b)
    Bd, ldb = extract_band(Ast, b)
    d, e, Vs, tau2, _ = sb2st_band(Bd, ldb, n, b, ncta=ncta, rng=rng)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, ZT = np.linalg.eigh(T)
    Z = apply_q1(Ast, tau1, b, apply_q2_wavefront(Vs, n, tau2, n, b, ZT))
    return w, Z, d, e


if __name__ == "__main__":
    rng = np.random.RandomState(1)
    for n, b in ((64, 8), (96, 16), (128, 32), (136, 8)):
        Mx = rng.standard_normal((n, n))
        A = Mx @ Mx.T
        w, Z, d, e = eigh_two_stage_model(A, b, rng=rng)
        res = np.linalg.norm(A @ Z - Z * w) / np.linalg.norm(A)
        print(n, b, "residual", res, "orth", np.linalg.norm(Z.T @ Z - np.eye(n)),
              "dw", np.abs(w - np.linalg.eigvalsh(A)).max() / w.max())
