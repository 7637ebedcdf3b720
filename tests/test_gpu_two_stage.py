"""GPU parity of the two-stage tridiagonal reduction (csrc/two_stage.cu; the default for n >= 8192, forced on here
with tq_set_eigh_two_stage(1) so that the small orders exercise it too).  Its kernels and host driver were first
checked on the CPU (tests/test_two_stage_emu.py); this file has passed on a B200 in full (profiles/r02_*).
Bars are those of the one-stage path (tests/test_gpu_solver.py)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture()
def two_stage():
    from gptq_svd_b200 import _lib, stages
    lib = _lib.load()
    lib.tq_set_eigh_two_stage(1)
    yield stages
    lib.tq_set_eigh_two_stage(-1)


def _spd(n, seed, decay=-3.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(2 * n, n, device="cuda", dtype=torch.float64, generator=g)
    X = X * torch.logspace(0, decay, n, device="cuda", dtype=torch.float64)[None, :]
    return X.T @ X / (2 * n)


@pytest.mark.parametrize("n", [256, 1024, 4096])
def test_each_stage_keeps_the_spectrum(two_stage, n):
    """Run this one first when something is off: it tells the band reduction from the bulge chase."""
    H = _spd(n, 5 + n)
    band, d, e = two_stage.two_stage_debug(H)
    wr = torch.linalg.eigvalsh(H)
    scale = float(wr.abs().max())
    i = torch.arange(n, device="cuda")
    Bf = torch.zeros(n, n, device="cuda", dtype=torch.float64)
    for off in range(0, 65):                       # band[c, off] = B[c + off, c]
        m = n - off
        Bf[i[:m] + off, i[:m]] = band[:m, off]
    Bf = torch.tril(Bf) + torch.tril(Bf, -1).T
    assert float(band[:, 65:].abs().max()) == 0.0, "bulge room is not empty after stage 1"
    assert float((torch.linalg.eigvalsh(Bf) - wr).abs().max()) <= 1e-12 * scale * n ** 0.5, "stage 1 (sy2sb)"
    T = torch.diag(d) + torch.diag(e, 1) + torch.diag(e, -1)
    assert float((torch.linalg.eigvalsh(T) - wr).abs().max()) <= 1e-12 * scale * n ** 0.5, "stage 2 (sb2st)"


@pytest.mark.parametrize("n", [256, 320, 1024, 4096])
def test_two_stage_eigh_is_an_eigendecomposition(two_stage, n):
    H = _spd(n, n)
    w, V = two_stage.eigh(H)
    wr = torch.linalg.eigvalsh(H)
    scale = float(wr.abs().max())
    assert bool((w[1:] >= w[:-1]).all())
    assert float((w - wr).abs().max()) <= 1e-12 * scale * n ** 0.5
    assert float(torch.linalg.norm(H @ V - V * w[None, :])) <= 1e-12 * float(torch.linalg.norm(H)) * n ** 0.5
    assert float(torch.linalg.norm(V.T @ V - torch.eye(n, device="cuda", dtype=torch.float64))) <= 1e-12 * n


def test_two_stage_matches_one_stage_through_the_solver(two_stage):
    """Same retained rank, pivots and factors as the default path (the eigenvectors of separated eigenvalues are
    unique up to sign, and R / R_x do not depend on that sign)."""
    import gptq_svd_b200 as G
    from gptq_svd_b200 import _lib
    n = 1024
    H = _spd(n, 11, decay=-2.0)
    f2 = G.spectral_solve(H, 1e-4, "energy")
    _lib.load().tq_set_eigh_two_stage(0)
    f1 = G.spectral_solve(H, 1e-4, "energy")
    assert f1.k == f2.k
    assert torch.equal(f1.perm[:f1.k], f2.perm[:f2.k])
    assert float((f1.R[:f1.k] - f2.R[:f2.k]).abs().max()) <= 1e-8 * float(f1.R[:f1.k].abs().max())
    assert float((f1.R_x[:f1.k] - f2.R_x[:f2.k]).abs().max()) <= 1e-8 * float(f1.R_x[:f1.k].abs().max())


def test_two_stage_falls_back_when_n_is_not_a_multiple_of_64(two_stage):
    H = _spd(300, 3)
    w, V = two_stage.eigh(H)
    assert float((w - torch.linalg.eigvalsh(H)).abs().max()) <= 1e-12 * float(w.abs().max()) * 300 ** 0.5
