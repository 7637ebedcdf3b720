"""GPU parity: quantisation grid + blocked GPTQ loop (tq_find_params, tq_gptq_loop,
tq_pack_codes, tq_quant_error) against the golden vectors of the unmodified reference
and against the CPU oracle.  Tolerances are north_star's: scales/zeros bit-equal,
integer codes >= 99.9 % identical, ||WX-QX|| within 1 %."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from oracle import truncgptq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gptq_svd_b200 as G
    return G


def _run(G, W, R, perm, bits, group, sym, block, use_triton, R_x=None, strict=False):
    from gpu_common import to_gpu
    q = G.Quantizer(bits, group, sym)
    res = G.gptq_quantize(to_gpu(W), to_gpu(R), q, to_gpu(perm), block_size=block,
                          use_triton=use_triton, R_x=None if R_x is None else to_gpu(R_x), strict_fp32=strict)
    torch.cuda.synchronize()
    return q, res


# trailing update: tcgen05 3xTF32 GEMM (default product path) and the strict-fp32 SIMT GEMM
TRAILING = [pytest.param(False, id="tcgen05_3xtf32"), pytest.param(True, id="simt_fp32")]


@pytest.mark.parametrize("strict", TRAILING)
@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("tag", ["triton", "torch"])
def test_loop_vs_reference_golden(G, name, tag, strict, golden):
    g = golden(name)
    bits, group, sym, block = int(g["bits"]), int(g["group"]), bool(g["sym"]), int(g["block"])
    q, res = _run(G, g["W"], g["R"], g["perm"], bits, group, sym, block, tag == "triton", g["R_x"], strict)
    assert res.rank == int(g["k"])
    assert np.array_equal(q.scale.cpu().numpy(), g["scale"])      # bit-exact grid
    assert np.array_equal(q.zero.cpu().numpy(), g["zero"])
    fw = res.final_W.cpu().numpy()
    ref = g[f"final_W_{tag}"]
    assert np.mean(fw == ref) >= 0.999
    oq = O.Quantizer(bits, group, sym)
    oq.find_params(g["W"])
    assert np.mean((res.codes.cpu().numpy().astype(np.int32) + res.min_q) == O.recover_codes(ref, oq)) >= 0.999
    ref_err = float(g[f"rel_err_{tag}"])
    assert abs(res.rel_error - ref_err) <= 0.01 * ref_err


@pytest.mark.parametrize("m,n,bits,group,sym,eps", [
    (2048, 1024, 4, 128, False, 1e-4),     # Qwen3-0.6B q_proj shape (BASELINE configs[0])
    (1024, 1024, 4, 128, True, 1e-4),
    (520, 384, 3, 128, False, 1e-6),       # ragged rows
    (257, 256, 2, -1, False, 1e-3),        # per-channel groups, odd row count
    (2048, 1024, 3, 128, False, 1e-4),     # BASELINE configs[2]: 3-bit asym g128 (run_benchmark.py:51-77)
    (2048, 1024, 2, 128, False, 1e-5),     # BASELINE configs[3]: 2-bit asym g128, published eps 1e-5
])
@pytest.mark.parametrize("strict", TRAILING)
def test_loop_vs_oracle(G, m, n, bits, group, sym, eps, strict):
    X = O.make_activations(8192, n, seed=7 * n + m, dist="llm").astype(np.float64)
    H = X.T @ X / X.shape[0]
    f = O.process_hessian_alt(H, eps, "energy")
    W = O.make_weight(m, n, seed=m + n)
    oq = O.Quantizer(bits, group, sym)
    fw_o, k, codes_o = O.gptq_fwrd(W, f.R, oq, f.perm, block_size=1024, use_triton=True, fma=True,
                                   return_codes=True)
    q, res = _run(G, W, f.R, f.perm, bits, group, sym, 1024, True, f.R_x, strict)
    print(f"codes identical to the oracle: {np.mean((res.codes.cpu().numpy().astype(np.int32) + res.min_q) == codes_o):.5%}")
    assert res.rank == k
    assert np.array_equal(q.scale.cpu().numpy(), oq.scale)
    assert np.array_equal(q.zero.cpu().numpy(), oq.zero)
    codes = res.codes.cpu().numpy().astype(np.int32) + res.min_q
    assert np.mean(codes == codes_o) >= 0.999
    assert codes.min() >= oq.min_q and codes.max() <= oq.max_q
    err_o = O.quantization_error(W, fw_o, f.R_x, f.perm)
    assert abs(res.rel_error - err_o) <= 0.01 * err_o
    # codes are exactly round(final_W / s + z)
    assert np.array_equal(O.recover_codes(res.final_W.cpu().numpy(), oq), codes)


@pytest.mark.parametrize("bits,sym", [(4, False), (3, False), (2, False), (8, False), (4, True), (3, True)])
def test_pack_matches_oracle(G, bits, sym):
    from gpu_common import to_gpu
    rng = np.random.RandomState(bits)
    nlev = 2 ** bits - (1 if sym else 0)
    codes = rng.randint(0, nlev, size=(37, 384)).astype(np.uint8)
    packed = G.pack_codes(to_gpu(codes), bits).cpu().numpy().view(np.uint32)
    assert np.array_equal(packed, O.pack_codes(codes.astype(np.int64), bits, 0))


def test_edge_cases(G):
    from gpu_common import to_gpu
    n, m = 256, 8
    W = O.make_weight(m, n, 3)
    perm = np.random.RandomState(0).permutation(n)
    # k = 0: pure RTN (half-even) of every column
    q = G.Quantizer(4, 128, False)
    fw, k = G.gptq_fwrd(to_gpu(W), torch.zeros((0, n), dtype=torch.float64, device="cuda"), q, to_gpu(perm))
    oq = O.Quantizer(4, 128, False)
    oq.find_params(W)
    s, z = oq.get_expanded_params(m, n)
    rtn = (np.clip(np.rint(W / s + z), 0, 15) - z) * s
    assert k == 0 and np.array_equal(fw.cpu().numpy(), rtn)
    # fp16 weights come back as fp16
    fw16, _ = G.gptq_fwrd(to_gpu(W).half(), torch.zeros((0, n), dtype=torch.float64, device="cuda"), q, to_gpu(perm))
    assert fw16.dtype == torch.float16
    # group size that does not divide n: AssertionError like the reference (gptq_utils.py:253)
    with pytest.raises(AssertionError):
        G.Quantizer(4, 100, False).find_params(to_gpu(W))
    # CPU tensors are refused loudly (no CPU fallback)
    with pytest.raises(RuntimeError):
        G.Quantizer(4, 128, False).find_params(torch.from_numpy(W))
    # a perm that is not a permutation (duplicate / out of range) fails before anything is gathered
    bad = perm.copy()
    bad[3] = bad[4]
    with pytest.raises(RuntimeError, match="not a permutation"):
        G.gptq_fwrd(to_gpu(W), torch.zeros((0, n), dtype=torch.float64, device="cuda"), q, to_gpu(bad))
    bad = perm.copy()
    bad[0] = n
    with pytest.raises(RuntimeError, match="not a permutation"):
        G.gptq_fwrd(to_gpu(W), torch.zeros((0, n), dtype=torch.float64, device="cuda"), q, to_gpu(bad))


def test_more_rows_than_a_grid_dimension(G):
    """m > 65535 (lm_head-sized Linears): the row loops of gather / tail / un-permute and the fused kernel's grid."""
    from gpu_common import to_gpu
    m, n = 66000, 256
    rng = np.random.RandomState(1)
    W = (rng.standard_normal((m, n)) * 0.02).astype(np.float16).astype(np.float32)
    X = O.make_activations(2048, n, seed=9, dist="llm").astype(np.float64)
    f = O.process_hessian_alt(X.T @ X / X.shape[0], 1e-3, "energy")
    res = G.gptq_quantize(to_gpu(W), to_gpu(f.R), G.Quantizer(4, 128, False), to_gpu(f.perm), block_size=1024)
    rows = np.r_[0:64, 65500:65600, m - 64:m]
    oq = O.Quantizer(4, 128, False)
    _, _, codes_o = O.gptq_fwrd(W[rows], f.R, oq, f.perm, block_size=1024, use_triton=True, fma=True, return_codes=True)
    codes = res.codes[torch.from_numpy(rows).cuda()].cpu().numpy().astype(np.int32) + res.min_q
    assert np.mean(codes == codes_o) >= 0.999


def test_full_size_properties(G):
    """Qwen3-8B o_proj shape (m = n = 4096): size-independent properties, no oracle."""
    torch.manual_seed(0)
    n = m = 4096
    X = torch.randn(16384, n, device="cuda", dtype=torch.float64) * torch.logspace(0, -2, n, device="cuda", dtype=torch.float64)
    H = X.T @ X / X.shape[0]
    Hinv = torch.linalg.inv(H + 1e-6 * torch.eye(n, device="cuda", dtype=torch.float64))
    R = torch.linalg.cholesky(Hinv, upper=True)            # R^T R = H^-1 (test scaffolding only)
    Rx = torch.linalg.cholesky(H, upper=True)
    perm = torch.arange(n, device="cuda")
    W = (torch.randn(m, n, device="cuda") * 0.02).half().float()
    q = G.Quantizer(4, 128, True)
    a = G.gptq_quantize(W, R, q, perm, block_size=1024, R_x=Rx)
    b = G.gptq_quantize(W, R, G.Quantizer(4, 128, True), perm, block_size=128, R_x=Rx)
    assert a.rank == n
    ca = a.codes.int() + a.min_q
    assert int(ca.min()) >= -7 and int(ca.max()) <= 7
    assert float((a.codes == b.codes).float().mean()) >= 0.999          # block-size independence
    s, z = q.get_expanded_params(m, n)
    rtn = (torch.clamp(torch.round(W / s + z), -7, 7) - z) * s
    e_rtn = G.log_quantization_error(W, rtn, Rx, perm)
    assert a.rel_error < e_rtn                                           # GPTQ beats RTN under H
    # deterministic
    c = G.gptq_quantize(W, R, G.Quantizer(4, 128, True), perm, block_size=1024)
    assert torch.equal(a.codes, c.codes)
    # tcgen05 3xTF32 trailing update vs strict-fp32 SIMT trailing update: >= 99.9 % identical codes
    d = G.gptq_quantize(W, R, G.Quantizer(4, 128, True), perm, block_size=1024, strict_fp32=True)
    same = float((a.codes == d.codes).float().mean())
    print(f"tcgen05 3xTF32 vs strict fp32 trailing update: {same:.5%} identical codes")
    assert same >= 0.999


@pytest.mark.parametrize("dense", [False, True], ids=["upper_trapezoidal", "dense_Rx"])
def test_metric_kernel_skips_only_structural_zeros(G, dense):
    """tq_quant_error starts each tile's reduction at the tile's first row when Rx is upper trapezoidal (the R of
    a QR, gptq_utils.py:124) and must NOT do so for a matrix with entries left of the diagonal.  Against an fp64
    torch reference of gptq_utils.py:275-291; tolerance 1e-3 relative (TF32 operands, bar 1 %)."""
    torch.manual_seed(5)
    m, n, k = 300, 1536, 1100
    W = torch.randn(m, n, device="cuda")
    Q = W + 0.05 * torch.randn(m, n, device="cuda")
    Rx = torch.randn(k, n, device="cuda", dtype=torch.float64)
    if not dense:
        Rx = torch.triu(Rx)
    else:
        Rx[700, 3] = 40.0                                  # a single entry far left of the diagonal must count
    perm = torch.randperm(n, device="cuda")
    got = G.log_quantization_error(W, Q, Rx, perm)
    Wp, Dp = W[:, perm].double(), (W - Q)[:, perm].double()
    ref = float(torch.linalg.norm(Dp @ Rx.T) / torch.linalg.norm(Wp @ Rx.T))
    assert abs(got - ref) <= 1e-3 * ref, (got, ref)
