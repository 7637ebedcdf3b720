"""GPU parity of the two other front ends (rows f3 / f4) and of the layer-by-layer pipeline (f2),
through the C ABI, against golden vectors from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import truncgptq_oracle as O
from oracle.make_golden_frontends import CHOL_CASES, SKETCH_CASES, sketch_blocks

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gptq_svd_b200 as G
    return G


def _gpu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("case", CHOL_CASES, ids=[c[0] for c in CHOL_CASES])
def test_process_hessian_golden(G, case):
    name, n, m, T, seed, actorder, damp, shift, bits, sym = case
    g = _load(name)
    U, perm = G.process_hessian(_gpu(g["H"]), actorder=actorder, damp_percent=damp)
    assert U.dtype == torch.float64 and tuple(U.shape) == (n, n)
    assert np.array_equal(perm.cpu().numpy(), g["perm"])
    Un = U.cpu().numpy()
    assert np.abs(Un - g["U"]).max() <= 1e-9 * np.abs(g["U"]).max()
    assert np.abs(np.tril(Un, -1)).max() == 0.0 and np.all(np.diagonal(Un) > 0)
    # U^T U (H_perm + damp mean I) = I
    Hp = g["H"][np.ix_(g["perm"], g["perm"])]
    e = 1 if shift else 0
    Hd = Hp + (10 ** e * damp) * np.mean(np.diag(Hp)) * np.eye(n)
    assert np.linalg.norm(Un.T @ Un @ Hd - np.eye(n)) <= 1e-8 * n
    # the loop with torch-loop arithmetic, as --mode gptq calls it (quantize.py:221-229)
    W = O.make_weight(m, n, seed=seed + 1)
    q = G.Quantizer(bits, 128, sym)
    fw, k = G.gptq_fwrd(_gpu(W), U, q, perm, block_size=1024, use_triton=False)
    assert k == n
    oq = O.Quantizer(bits, 128, sym)
    oq.find_params(W)
    c0, c1 = O.recover_codes(fw.cpu().numpy(), oq), O.recover_codes(g["final_W"], oq)
    assert np.mean(c0 == c1) >= 0.999


def test_process_hessian_identity_fallback(G):
    H = -torch.eye(64, dtype=torch.float64, device="cuda")          # no rung of the ladder can fix this
    U, perm = G.process_hessian(H)
    assert torch.equal(U, torch.eye(64, dtype=torch.float64, device="cuda"))
    assert torch.equal(perm, torch.arange(64, device="cuda"))


@pytest.mark.parametrize("case", SKETCH_CASES, ids=[c[0] for c in SKETCH_CASES])
def test_sketch_golden(G, case):
    name, n, T, seed, rank, thr, method = case
    g = _load(name)
    X = O.make_activations(T, n, seed=seed, dist="llm")
    lin = torch.nn.Linear(n, 8, bias=False)
    sk = G.Sketcher(lin, rank, device="cuda")
    c = 0
    for i, Rb in enumerate(sketch_blocks(seed, rank, T)):
        xb = _gpu(X[c:c + Rb.shape[1]])
        sk.add_batch(xb.view(2, -1, n) if i == 0 else xb, _gpu(Rb))    # one 3-D batch, like a hook sees
        c += Rb.shape[1]
    assert sk.n_samples == T
    Y = sk.get_scaled_sketch()
    Yn = Y.cpu().numpy()
    assert np.linalg.norm(Yn - g["Y"]) <= 1e-5 * np.linalg.norm(g["Y"])
    R, perm = G.process_sketch(_gpu(g["Y"]), thr, method)
    k = int(g["k"])
    assert R.shape[0] == k and R.dtype == torch.float64 and perm.dtype == torch.int64
    p = perm.cpu().numpy()
    assert sorted(p.tolist()) == list(range(n))
    assert np.array_equal(p[:k], g["perm"][:k])
    Rn = R.cpu().numpy()
    S = np.linalg.svd(g["Y"].astype(np.float64), compute_uv=False)
    cond = (S[0] / S[k - 1]) ** 2
    assert np.abs(Rn - g["R"]).max() <= (1e-13 * cond + 1e-10) * np.abs(g["R"]).max()
    # the hook path draws its own Gaussian block and only has to run
    sk2 = G.Sketcher(lin, rank, device="cuda")
    sk2.hook_fn(lin, (_gpu(X[:256]),), None)
    assert sk2.n_samples == 256 and float(sk2.Y.abs().sum()) > 0


def _tiny_qwen3(seed=0):
    from transformers import Qwen3Config, Qwen3ForCausalLM
    torch.manual_seed(seed)
    cfg = Qwen3Config(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, head_dim=64, max_position_embeddings=256,
                      tie_word_embeddings=False)
    model = Qwen3ForCausalLM(cfg).to(torch.float16).cuda().eval()
    return model


@pytest.mark.parametrize("mode", ["eigh", "gptq", "svd", "eigh-parallel-groups"])
def test_pipeline_tiny_qwen3(G, mode):
    """quantize.main's loop on a random-init Qwen3: every decoder Linear ends on its grid, the
    model still runs, and GPTQ-style error compensation beats plain round-to-nearest."""
    from gptq_svd_b200.pipeline import get_layers, get_sequenced_groups, get_submodule, quantize_model
    model = _tiny_qwen3()
    ref = _tiny_qwen3()
    rtn = _tiny_qwen3()
    g = torch.Generator().manual_seed(1)
    ids = [torch.randint(0, 512, (1, 128), generator=g) for _ in range(16)]
    true_seq = mode != "eigh-parallel-groups"       # False: one capture pass per layer, solves side by side
    mode = mode.split("-")[0]
    out = quantize_model(model, ids, mode=mode, w_bits=4, group_size=128, sym=False, eps=1e-4,
                         threshold_method="energy", batch_size=8, device="cuda", keep_packed=True,
                         true_sequential=true_seq)
    assert len(out["layer_stats"]) == 2 * 7
    for li, (layer, layer0, layer_r) in enumerate(zip(get_layers(model), get_layers(ref), get_layers(rtn))):
        for group in get_sequenced_groups(layer):
            for name in group:
                Wq = get_submodule(layer, name).weight.data.float()
                W0 = get_submodule(layer0, name).weight.data.float()
                assert not torch.equal(Wq, W0)
                q = G.Quantizer(4, 128, False)
                q.find_params(W0)
                s, z = q.get_expanded_params(*W0.shape)
                codes = Wq / s + z
                assert float((codes - codes.round()).abs().max()) <= 2e-2          # fp16 storage of the grid values
                pk = out["packed"][f"layer_{li}.{name}"]
                assert pk["qweight"].dtype in (torch.int32, torch.uint32) and pk["scale"].shape[0] == W0.shape[0]
                # plain RTN on the same grid, for the comparison below
                Wr = ((W0 / s + z).round().clamp(0, 15) - z) * s
                get_submodule(layer_r, name).weight.data.copy_(Wr)
    x = torch.cat(ids[:4]).cuda()
    with torch.no_grad():
        y0 = ref(x).logits.float()
        yq = model(x).logits.float()
        yr = rtn(x).logits.float()
    assert torch.isfinite(yq).all()
    err_q = float(torch.linalg.norm(yq - y0) / torch.linalg.norm(y0))
    err_r = float(torch.linalg.norm(yr - y0) / torch.linalg.norm(y0))
    assert err_q < 0.5
    if mode != "svd":      # a rank-n Gaussian sketch of 2048 tokens is a noisy Hessian estimate: parity is checked above
        assert err_q < err_r * 1.05, (err_q, err_r)


def test_solver_pool_matches_sequential(G):
    """Three solves in flight (one thread, stream and SM budget each) agree with the one-at-a-time call:
    same k, same pivots, factors to rounding (the partial-sum order depends on the grid size)."""
    from gptq_svd_b200.concurrent import SolverPool
    Hs = []
    for seed in range(3):
        X = O.make_activations(2048, 512, seed=40 + seed, dist="llm").astype(np.float64)
        Hs.append(_gpu(X.T @ X / X.shape[0]))
    seq = [G.spectral_solve(H, 1e-4, "energy") for H in Hs]
    pool = SolverPool(workers=3)
    try:
        for budget in (None, 20):
            res = pool.spectral_solve_many(Hs, 1e-4, "energy", sm_budget=budget)
            for r, q in zip(res, seq):
                assert r.k == q.k
                assert torch.equal(r.perm[:r.k], q.perm[:q.k])
                assert float((r.R - q.R).abs().max() / q.R.abs().max()) <= 1e-9
                assert float((r.R_x - q.R_x).abs().max() / q.R_x.abs().max()) <= 1e-10
        again = pool.spectral_solve_many(Hs, 1e-4, "energy", sm_budget=20)
        for r, q in zip(again, res):
            assert torch.equal(r.R, q.R)                      # reproducible for a given budget
        out = pool.process_hessian_alt_many(Hs[:2], 1e-4, "energy")
        assert len(out) == 2 and out[0][0].shape == seq[0].R.shape
        with pytest.raises(RuntimeError):
            pool.spectral_solve_many([torch.eye(4, dtype=torch.float64)], 1e-4, "energy")   # CPU tensor: no fallback
    finally:
        pool.close()
    _ = G  # the main thread's budget is untouched
    f = G.spectral_solve(Hs[0], 1e-4, "energy")
    assert torch.equal(f.R, seq[0].R)


def test_stage_callback(G):
    """tq_set_stage_callback: the callback runs once per solve on the calling thread when the tridiagonal
    reduction has completed on the device; results are unchanged; NULL removes it."""
    from gptq_svd_b200 import _lib
    lib = _lib.load()
    X = O.make_activations(2048, 384, seed=77, dist="llm").astype(np.float64)
    H = _gpu(X.T @ X / X.shape[0])
    base = G.spectral_solve(H, 1e-4, "energy")
    seen = []
    cb = _lib.STAGE_CALLBACK(lambda stage, user: seen.append(stage))
    lib.tq_set_stage_callback(cb, None)
    try:
        f = G.spectral_solve(H, 1e-4, "energy")
    finally:
        lib.tq_set_stage_callback(_lib.STAGE_CALLBACK(0), None)
    assert seen == [_lib.TQ_STAGE_SYTRD_DONE]
    assert f.k == base.k and torch.equal(f.R, base.R) and torch.equal(f.perm, base.perm)
    G.spectral_solve(H, 1e-4, "energy")
    assert seen == [_lib.TQ_STAGE_SYTRD_DONE]              # removed
    # an SM budget changes the grid, not the answer (to rounding); 0 restores the whole GPU
    lib.tq_set_sm_budget(24)
    try:
        g = G.spectral_solve(H, 1e-4, "energy")
    finally:
        lib.tq_set_sm_budget(0)
    assert g.k == base.k and torch.equal(g.perm[:g.k], base.perm[:base.k])
    assert float((g.R - base.R).abs().max() / base.R.abs().max()) <= 1e-9
