"""Pin the CPU oracle: against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), against LAPACK dgeqp3 (scipy), and against the
oracle-free invariants of SURVEY.md 8(c6).  CPU only."""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import truncgptq_oracle as O


@pytest.mark.parametrize("name", golden_cases())
def test_solver_matches_reference(name, golden):
    g = golden(name)
    f = O.process_hessian_alt(g["H"], float(g["eps"]), str(g["method"]))
    assert f.k == int(g["k"])
    np.testing.assert_allclose(f.eigvals, g["eig"], rtol=1e-9, atol=1e-12)
    k = f.k
    assert np.array_equal(f.perm[:k], g["perm"][:k])
    # R = qr(Lambda^-1/2 V^T P) is only reproducible to eps_mach * cond(H_k) between two
    # fp64 LAPACK builds (numpy's vs torch's eigh): measured 5e-17 * cond.
    cond = f.eigvals[0] / f.eigvals[k - 1]
    scale = np.abs(g["R"]).max()
    assert np.abs(f.R - g["R"]).max() <= (2e-15 * cond + 1e-13) * scale
    scale_x = np.abs(g["R_x"]).max()
    assert np.abs(f.R_x - g["R_x"]).max() <= 1e-11 * scale_x
    assert np.all(np.diagonal(f.R) > 0)


@pytest.mark.parametrize("name", ["llm_n128_w4a", "flat_n128_w4a"])
def test_hessian_matches_reference(name, golden):
    g = golden(name)
    X = g["X"]
    acc = O.HessianAccumulator(X.shape[1])
    for c in range(0, X.shape[0], 1024):
        xb = X[c:c + 1024]
        acc.add_batch(xb.reshape(2, -1, X.shape[1]) if c == 0 else xb)
    H = acc.get_hessian()
    assert acc.n_samples == int(g["n_tokens"])
    assert np.linalg.norm(H - g["H"]) <= 1e-13 * np.linalg.norm(g["H"])


@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("tag", ["triton", "torch"])
def test_loop_matches_reference(name, tag, golden):
    g = golden(name)
    q = O.Quantizer(int(g["bits"]), int(g["group"]), bool(g["sym"]))
    fw, k = O.gptq_fwrd(g["W"], g["R"], q, g["perm"], block_size=int(g["block"]),
                        use_triton=(tag == "triton"))
    assert k == int(g["k"])
    # grid parameters are bit-exact
    assert np.array_equal(q.scale, g["scale"])
    assert np.array_equal(q.zero, g["zero"])
    ref = g[f"final_W_{tag}"]
    # dequantised values: identical codes almost everywhere (BLAS summation order only)
    same = np.mean(fw == ref)
    assert same >= 0.999, same
    err = O.quantization_error(g["W"], fw, g["R_x"], g["perm"])
    assert abs(err - float(g[f"rel_err_{tag}"])) <= 1e-2 * float(g[f"rel_err_{tag}"])
    codes = O.recover_codes(fw, q)
    assert codes.min() >= q.min_q and codes.max() <= q.max_q


@pytest.mark.parametrize("shape", [(40, 64), (96, 200), (300, 300), (200, 520), (260, 150)])
def test_dgeqp3_matches_lapack(shape):
    import scipy.linalg as sla

    rng = np.random.RandomState(shape[0] * 1000 + shape[1])
    m, n = shape
    A = rng.standard_normal((m, n)) * np.logspace(0, -6, n)[rng.permutation(n)][None, :]
    R, p = O.dgeqp3(A)
    _, R2, p2 = sla.qr(A, mode="economic", pivoting=True)
    kk = min(m, n)
    assert np.array_equal(p[:kk], p2[:kk])
    sgn = np.sign(np.diagonal(R)) * np.sign(np.diagonal(R2))
    assert np.abs(R - sgn[:, None] * R2).max() <= 1e-12 * np.abs(R2).max()


def test_dgeqp3_ties_first_index():
    # identical column norms: idamax picks the first
    A = np.eye(6)[:, [2, 0, 1, 5, 4, 3]].copy()
    R, p = O.dgeqp3(A)
    assert p[0] == 0


@pytest.mark.parametrize("name", ["llm_n256_w3a", "flat_n128_w4a"])
def test_invariants(name, golden):
    g = golden(name)
    H = g["H"]
    f = O.process_hessian_alt(H, float(g["eps"]), "energy")
    L, V = np.linalg.eigh(H)
    L = np.maximum(L, 1e-12)[::-1]
    V = V[:, ::-1]
    k = f.k
    Hk = (V[:, :k] * L[:k]) @ V[:, :k].T
    Hkp = (V[:, :k] / L[:k]) @ V[:, :k].T
    P = f.perm
    assert np.linalg.norm(f.R_x.T @ f.R_x - Hk[np.ix_(P, P)]) <= 1e-12 * np.linalg.norm(H)
    assert np.linalg.norm(f.R.T @ f.R - Hkp[np.ix_(P, P)]) <= 1e-9 * np.linalg.norm(Hkp)
    # GPTQ error under H_k is no worse than RTN's
    q = O.Quantizer(int(g["bits"]), int(g["group"]), bool(g["sym"]))
    fw, _ = O.gptq_fwrd(g["W"], f.R, q, f.perm, block_size=64)
    s, z = q.get_expanded_params(*g["W"].shape)
    rtn = (np.clip(np.rint(g["W"] / s + z), q.min_q, q.max_q) - z) * s
    e_g = O.quantization_error(g["W"], fw, f.R_x, f.perm)
    e_r = O.quantization_error(g["W"], rtn, f.R_x, f.perm)
    assert e_g <= e_r
    # block-size independence up to fp32 noise
    fw2, _ = O.gptq_fwrd(g["W"], f.R, O.Quantizer(int(g["bits"]), int(g["group"]), bool(g["sym"])),
                         f.perm, block_size=1024)
    assert np.mean(fw == fw2) >= 0.999


def test_rank_rule_edges():
    e = np.array([4.0, 3.0, 2.0, 1.0])
    assert O.rank_rule(e, 0.0, "energy") == 4           # cumsum <= total everywhere
    assert O.rank_rule(e, 0.5, "energy") == 2           # 4 <= 5 -> 1, +1
    assert O.rank_rule(e, 1.0, "energy") == 1           # nothing <= 0 -> 0, +1
    assert O.rank_rule(e, 0.3, "none") == 4
    assert O.rank_rule(e, 0.9, "mean_trimmed") == int(np.sum(np.sqrt(e) > 0.9 * np.mean(np.sqrt(e)[1:])))


@pytest.mark.parametrize("bits,min_q", [(4, 0), (3, 0), (2, 0), (8, 0), (4, -7), (3, -3)])
def test_pack_roundtrip(bits, min_q):
    rng = np.random.RandomState(bits)
    n = 256
    codes = rng.randint(min_q, min_q + (2 ** bits if min_q == 0 else 2 ** bits - 1), size=(5, n))
    words = O.pack_codes(codes, bits, min_q)
    assert words.shape == (5, (n * bits + 31) // 32)
    bitsarr = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, :n * bits]
    vals = (bitsarr.reshape(5, n, bits) * (1 << np.arange(bits))).sum(-1) + min_q
    assert np.array_equal(vals, codes)


def test_quantizer_asserts_on_bad_group():
    q = O.Quantizer(4, 128, False)
    with pytest.raises(AssertionError):
        q.find_params(np.zeros((4, 100), dtype=np.float32))
