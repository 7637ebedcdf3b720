"""Packed-weight output: the GPTQ / AutoGPTQ / vLLM checkpoint layout.  CPU: the oracle-side reader /
writer against itself and against the LSB-first bitstream; GPU: `export_gptq` read back by the independent reader
reproduces final_W bit for bit."""
import numpy as np
import pytest

from oracle import gptq_format as F
from oracle import truncgptq_oracle as O


@pytest.mark.parametrize("bits", [2, 3, 4, 8])
def test_reader_inverts_writer_and_is_the_lsb_first_bitstream(bits):
    rng = np.random.RandomState(bits)
    K, N = 256, 24
    vals = rng.randint(0, 1 << bits, size=(K, N))
    words = F.pack_rows(vals, bits)
    assert words.shape == (K * bits // 32, N) and words.dtype == np.uint32
    assert np.array_equal(F.unpack_rows(words, bits), vals)
    # AutoGPTQ's schemes (including the 3-bit 32-in-3-words one) are the little-endian bitstream of each column
    stream = O.pack_codes(vals.T.astype(np.int64), bits, 0)                 # [N, K * bits / 32]
    assert np.array_equal(words, stream.T)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,sym,group", [(4, True, 128), (4, False, 128), (3, False, 128), (2, False, 128),
                                            (8, False, 128), (3, True, 128), (4, False, -1)])
def test_export_round_trip_bit_exact(bits, sym, group):
    import torch
    import gptq_svd_b200 as G
    m, n = 256, 512
    X = O.make_activations(4096, n, seed=bits, dist="llm").astype(np.float64)
    f = O.process_hessian_alt(X.T @ X / X.shape[0], 1e-4, "energy")
    W = O.make_weight(m, n, seed=10 + bits)
    q = G.gptq_quantize(torch.from_numpy(W).cuda(), torch.from_numpy(f.R).cuda(), G.Quantizer(bits, group, sym),
                        torch.from_numpy(f.perm).cuda(), block_size=1024)
    ck = G.export_gptq(q, group, scales_dtype=torch.float32)
    g = group if group > 0 else n
    assert ck["qweight"].shape == (n * bits // 32, m) and ck["qweight"].dtype == torch.int32
    assert ck["qzeros"].shape == (n // g, m * bits // 32) and ck["scales"].shape == (n // g, m)
    assert np.array_equal(ck["g_idx"].cpu().numpy(), np.arange(n) // g)
    Wd = F.dequantize(ck["qweight"].cpu().numpy(), ck["qzeros"].cpu().numpy(), ck["scales"].cpu().numpy(),
                      ck["g_idx"].cpu().numpy(), bits)
    assert np.array_equal(Wd, q.final_W.cpu().numpy())                      # bit for bit with fp32 scales
    # the format's fp16 scales: same codes and zeros, values within fp16 rounding of the scale
    ck16 = G.export_gptq(q, group)
    assert ck16["scales"].dtype == torch.float16 and torch.equal(ck16["qweight"], ck["qweight"])
    W16 = F.dequantize(ck16["qweight"].cpu().numpy(), ck16["qzeros"].cpu().numpy(), ck16["scales"].cpu().numpy(),
                       ck16["g_idx"].cpu().numpy(), bits)
    assert np.abs(W16 - Wd).max() <= 1e-3 * np.abs(Wd).max() + 1e-8
    # the zero-offset-free variant (GPTQModel "v2") stores the zero itself
    ck2 = G.export_gptq(q, group, scales_dtype=torch.float32, v1_zero_offset=False)
    W2 = F.dequantize(ck2["qweight"].cpu().numpy(), ck2["qzeros"].cpu().numpy(), ck2["scales"].cpu().numpy(),
                      ck2["g_idx"].cpu().numpy(), bits, zero_offset=0)
    assert np.array_equal(W2, Wd)
