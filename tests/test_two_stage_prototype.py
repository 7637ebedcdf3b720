"""The CPU prototype of the two-stage tridiagonal reduction (round-2 preparation,
scripts/prototypes/two_stage_tridiag.py) stays an oracle only if it keeps passing."""
import numpy as np

from scripts.prototypes.two_stage_tridiag import eigh_two_stage, sy2sb


def test_two_stage_reduction_matches_eigh():
    rng = np.random.RandomState(1)
    n, b = 97, 8
    M = rng.standard_normal((n, n)) * np.logspace(0, -3, n)[None, :]
    A = M @ M.T
    band, Q1 = sy2sb(A, b)
    assert np.abs(np.tril(band, -(b + 1))).max() <= 1e-13 * np.abs(A).max()
    assert np.linalg.norm(Q1 @ band @ Q1.T - A) <= 1e-12 * np.linalg.norm(A)
    w, Z, off, nref = eigh_two_stage(A, b)
    assert off <= 1e-12 * np.abs(A).max()
    assert np.abs(w - np.linalg.eigvalsh(A)).max() <= 1e-13 * w.max()
    assert np.linalg.norm(A @ Z - Z * w[None, :]) <= 1e-12 * np.linalg.norm(A)
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) <= 1e-11
