"""TEST INFRASTRUCTURE (oracle side): an independent reader / writer of the GPTQ checkpoint layout
(AutoGPTQ v1 `QuantLinear`, the layout vLLM's `gptq` loader consumes), restated from the published
algorithm - the reference repository has no packed output (README.md:133 lists it as a roadmap item), so this
file is what the product's `export_gptq` / `tq_pack_gptq` is checked against.  numpy only.

Layout (in_features = K, out_features = N, pack factor pf = 32 // bits for bits in {2, 4, 8}):
    qweight [K // pf, N] int32   qweight[r, j] = sum_t  w[r * pf + t, j] << (bits * t)
    qzeros  [G, N // pf] int32   qzeros[g, c]  = sum_t ((z[g, c * pf + t] - 1) & mask) << (bits * t)     (v1: zero - 1)
    scales  [G, N]       fp16
    g_idx   [K]          int32   group of input channel i
3 bits: 32 values in 3 words, AutoGPTQ's explicit scheme (`pack3` below): values 0-9 in bits 0-29 of word 0, value 10
in bits 30-31 of word 0 and bit 0 of word 1, values 11-20 in bits 1-30 of word 1, value 21 in bit 31 of word 1 and
bits 0-1 of word 2, values 22-31 in bits 2-31 of word 2.
Dequantisation: W[j, i] = scales[g_idx[i], j] * (w[i, j] - ((unpack(qzeros)[g_idx[i], j] + 1) & mask))."""
import numpy as np


def pack_rows(vals: np.ndarray, bits: int) -> np.ndarray:
    """vals [K, N] unsigned (< 2^bits) -> words [K * bits // 32, N] uint32, packed along axis 0."""
    vals = np.asarray(vals).astype(np.uint64)
    K, N = vals.shape
    if bits in (2, 4, 8):
        pf = 32 // bits
        assert K % pf == 0
        out = np.zeros((K // pf, N), dtype=np.uint64)
        for t in range(pf):
            out |= vals[t::pf] << np.uint64(bits * t)
        return out.astype(np.uint32)
    if bits == 3:
        assert K % 32 == 0
        out = np.zeros((K // 32 * 3, N), dtype=np.uint64)
        i, row = 0, 0
        while row < out.shape[0]:
            for j in range(i, i + 10):
                out[row] |= vals[j] << np.uint64(3 * (j - i))
            i += 10
            out[row] |= vals[i] << np.uint64(30)
            row += 1
            out[row] |= (vals[i] >> np.uint64(2)) & np.uint64(1)
            i += 1
            for j in range(i, i + 10):
                out[row] |= vals[j] << np.uint64(3 * (j - i) + 1)
            i += 10
            out[row] |= vals[i] << np.uint64(31)
            row += 1
            out[row] |= (vals[i] >> np.uint64(1)) & np.uint64(3)
            i += 1
            for j in range(i, i + 10):
                out[row] |= vals[j] << np.uint64(3 * (j - i) + 2)
            i += 10
            row += 1
        return (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    raise ValueError(f"bits={bits}")


def unpack_rows(words: np.ndarray, bits: int) -> np.ndarray:
    """Inverse of pack_rows: [R, N] uint32 -> [R * 32 // bits, N] int64."""
    w = np.asarray(words).astype(np.uint32).astype(np.uint64)
    R, N = w.shape
    mask = np.uint64((1 << bits) - 1)
    if bits in (2, 4, 8):
        pf = 32 // bits
        out = np.zeros((R * pf, N), dtype=np.int64)
        for t in range(pf):
            out[t::pf] = ((w >> np.uint64(bits * t)) & mask).astype(np.int64)
        return out
    if bits == 3:
        assert R % 3 == 0
        out = np.zeros((R // 3 * 32, N), dtype=np.int64)
        for blk in range(R // 3):
            w0, w1, w2 = w[3 * blk], w[3 * blk + 1], w[3 * blk + 2]
            o = out[32 * blk:32 * blk + 32]
            for t in range(10):
                o[t] = (w0 >> np.uint64(3 * t)) & mask
            o[10] = ((w0 >> np.uint64(30)) & np.uint64(3)) | ((w1 & np.uint64(1)) << np.uint64(2))
            for t in range(10):
                o[11 + t] = (w1 >> np.uint64(3 * t + 1)) & mask
            o[21] = ((w1 >> np.uint64(31)) & np.uint64(1)) | ((w2 & np.uint64(3)) << np.uint64(1))
            for t in range(10):
                o[22 + t] = (w2 >> np.uint64(3 * t + 2)) & mask
        return out
    raise ValueError(f"bits={bits}")


def dequantize(qweight, qzeros, scales, g_idx, bits: int, zero_offset: int = 1) -> np.ndarray:
    """W [out_features, in_features] fp32 from the checkpoint tensors (AutoGPTQ v1 QuantLinear.forward)."""
    mask = (1 << bits) - 1
    w = unpack_rows(np.asarray(qweight).view(np.uint32), bits)                              # [K, N]
    z = unpack_rows(np.ascontiguousarray(np.asarray(qzeros).view(np.uint32).T), bits).T      # [G, N]
    z = (z + zero_offset) & mask
    s = np.asarray(scales).astype(np.float32)                                               # [G, N]
    g = np.asarray(g_idx).astype(np.int64)
    W = (w.astype(np.float32) - z[g].astype(np.float32)) * s[g]                             # [K, N], fp32: (q - z) * s
    return np.ascontiguousarray(W.T)
