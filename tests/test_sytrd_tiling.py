"""CPU model of the index arithmetic of `sytrd_panel_sym_kernel` (gptq_svd_b200/csrc/eigh.cu): the lower
triangle of the trailing matrix is cut into 128 x 64 tiles whose row blocks start at a multiple of `align`
rows, tiles are dealt to CTAs in contiguous chunks of the row-block-major order, every CTA flushes one row
partial per row block into `rowpart[slot]` and one column partial per tile into `colpart[row block]`, and
the consumer adds, for row r, `nslots` row partials and the column partials of the row blocks from
`(r_local // 64) // 2` on.  The model replays exactly that and must reproduce A22 @ u for every column -
this is the check that was run before the kernel first touched a GPU."""
import math

import numpy as np
import pytest

TR, TC, ROW_SLOTS = 128, 64, 40


def product_by_tiles(A, c, G, history, align):
    n = A.shape[0]
    base, ln = c + 1, n - c - 1
    dl = base % align
    base_e = base - dl
    nrb = (ln + dl + TR - 1) // TR
    nstrips = (ln + TC - 1) // TC
    TA = nrb * (nrb - 1) + min(2 * nrb, nstrips)
    T = TA + (2 * nrb if history else 0)              # W^T u / V^T u tiles follow the matrix tiles
    ch = max(1, (T + G - 1) // G)
    u = np.random.RandomState(c).standard_normal(ln)
    Ap = np.zeros((n + 4 * TR, n + 2 * TC))
    Ap[:n, :n] = A                                    # TMA: out-of-bounds reads as zero
    rowpart = np.full((ROW_SLOTS, n), np.nan)
    colpart = np.full((nrb + 1, n), np.nan)
    for b in range(G):
        t0 = min(T, b * ch)
        t1 = min(T, t0 + ch)
        rb_rows, yrow = -1, None

        def flush():
            nonlocal rb_rows, yrow
            slot = b - (rb_rows * (rb_rows + 1)) // ch
            assert 0 <= slot < ROW_SLOTS
            rl = rb_rows * TR + np.arange(TR) - dl
            m = (rl >= 0) & (rl < ln)
            rowpart[slot, base + rl[m]] = yrow[m]
            rb_rows, yrow = -1, None

        for t in range(t0, t1):
            if t >= TA:
                if rb_rows >= 0:
                    flush()
                continue
            rb = int((math.sqrt(4.0 * t + 1.0) - 1.0) * 0.5)
            while rb * (rb + 1) > t:
                rb -= 1
            while (rb + 1) * (rb + 2) <= t:
                rb += 1
            cs = t - rb * (rb + 1)
            assert cs < min(2 * rb + 2, nstrips)
            if rb_rows >= 0 and rb_rows != rb:
                flush()
            if rb_rows < 0:
                rb_rows, yrow = rb, np.zeros(TR)
            x, y = base_e + rb * TR, base + cs * TC
            assert x % align == 0                      # 16-byte (align 2) / 256-byte (align 32) box origin
            tile = Ap[x:x + TR, y:y + TC]
            rl = rb * TR + np.arange(TR) - dl
            cl = cs * TC + np.arange(TC)
            ur = np.where((rl >= 0) & (rl < ln), u[np.clip(rl, 0, ln - 1)], 0.0)
            uc = np.where(cl < ln, u[np.clip(cl, 0, ln - 1)], 0.0)
            yrow += (tile * (rl[:, None] >= cl[None, :])) @ uc
            tcol = (tile * (rl[:, None] > cl[None, :])).T @ ur
            m = cl < ln
            colpart[rb, base + cl[m]] = tcol[m]
        if rb_rows >= 0:
            flush()
    y = np.zeros(ln)
    for r in range(base, n):
        rl = r - base
        rbr = (rl + dl) // TR
        ts, ncs = rbr * (rbr + 1), min(2 * rbr + 2, nstrips)
        nslots = (ts + ncs - 1) // ch - ts // ch + 1
        rbc = (rl // TC) // 2
        y[rl] = rowpart[:nslots, r].sum() + colpart[rbc:nrb, r].sum()
    ref = A[base:, base:] @ u
    return float(np.abs(y - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("n,align,G", [(300, 2, 148), (300, 32, 148), (700, 32, 148), (520, 32, 49), (1024, 32, 16)])
def test_tile_partials_reproduce_the_product(n, align, G):
    rng = np.random.RandomState(n)
    M = rng.standard_normal((n, n))
    A = M + M.T
    cols = list(range(0, n - 1, 7)) + [n - 3, n - 2]
    worst = max(product_by_tiles(A, c, G, history=(c % 64) > 0, align=align) for c in cols)
    assert worst <= 1e-12
