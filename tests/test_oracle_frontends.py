"""CPU: the oracle's restatement of the two other front ends (process_hessian, Sketcher /
process_sketch) against golden vectors generated from the unmodified reference
(oracle/make_golden_frontends.py), plus the host-side helpers of the pipeline."""
import os

import numpy as np
import pytest

from oracle import truncgptq_oracle as O
from oracle.make_golden_frontends import CHOL_CASES, SKETCH_CASES, sketch_blocks

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("case", CHOL_CASES, ids=[c[0] for c in CHOL_CASES])
def test_oracle_process_hessian(case):
    name, n, m, T, seed, actorder, damp, shift, bits, sym = case
    g = _load(name)
    U, perm, e = O.process_hessian(g["H"], actorder=actorder, damp_percent=damp)
    assert np.array_equal(perm, g["perm"])
    assert e == (1 if shift else 0)                       # the shifted case needs the second rung
    assert np.abs(U - g["U"]).max() <= 1e-9 * np.abs(g["U"]).max()
    assert np.abs(np.tril(U, -1)).max() == 0.0
    W = O.make_weight(m, n, seed=seed + 1)
    q = O.Quantizer(bits, 128, sym)
    fw, k = O.gptq_fwrd(W, g["U"], q, g["perm"], block_size=1024, use_triton=False)
    assert k == n
    c0, c1 = O.recover_codes(fw, q), O.recover_codes(g["final_W"], q)
    assert np.mean(c0 == c1) >= 0.999


@pytest.mark.parametrize("case", SKETCH_CASES, ids=[c[0] for c in SKETCH_CASES])
def test_oracle_sketch(case):
    name, n, T, seed, rank, thr, method = case
    g = _load(name)
    X = O.make_activations(T, n, seed=seed, dist="llm")
    sk = O.Sketcher(n, rank)
    c = 0
    for Rb in sketch_blocks(seed, rank, T):
        sk.add_batch(X[c:c + Rb.shape[1]], Rb)
        c += Rb.shape[1]
    Y = sk.get_scaled_sketch()
    assert np.linalg.norm(Y - g["Y"]) <= 1e-5 * np.linalg.norm(g["Y"])       # fp32 GEMM, summation order differs
    R, perm, k = O.process_sketch(g["Y"], thr, method)
    assert k == int(g["k"])
    assert np.array_equal(perm[:k], g["perm"][:k])
    assert np.abs(R - g["R"]).max() <= 1e-8 * np.abs(g["R"]).max()


def test_pipeline_host_helpers():
    torch = pytest.importorskip("torch")
    from torch import nn
    from gptq_svd_b200 import pipeline as P

    assert P.get_adaptive_eps("mlp.down_proj", 1e-4) == pytest.approx(1e-5)
    assert P.get_adaptive_eps("self_attn.o_proj", 1e-4) == pytest.approx(1e-5)
    assert P.get_adaptive_eps("self_attn.q_proj", 1e-4) == 1e-4

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(8, 8) for _ in range(4))

    class Mlp(nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_proj, self.up_proj, self.down_proj = nn.Linear(8, 16), nn.Linear(8, 16), nn.Linear(16, 8)

    class Layer(nn.Module):
        def __init__(self):
            super().__init__()
            self.self_attn, self.mlp = Attn(), Mlp()

    layer = Layer()
    assert P.get_sequenced_groups(layer) == [["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"],
                                             ["self_attn.o_proj"], ["mlp.gate_proj", "mlp.up_proj"],
                                             ["mlp.down_proj"]]
    assert P.get_submodule(layer, "mlp.down_proj") is layer.mlp.down_proj
    del layer.self_attn.o_proj
    assert ["self_attn.o_proj"] not in P.get_sequenced_groups(layer)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = nn.Module()
            self.model.layers = nn.ModuleList([Layer()])

    assert P.get_layers(M())[0].mlp is not None
    with pytest.raises(ValueError):
        P.get_layers(nn.Linear(2, 2))
    # no CPU path: the front ends refuse CPU tensors
    import gptq_svd_b200 as G
    with pytest.raises(RuntimeError):
        G.process_hessian(torch.eye(4, dtype=torch.float64))
    with pytest.raises(RuntimeError):
        G.process_sketch(torch.ones(4, 4))


def test_solver_pool_fails_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from gptq_svd_b200.concurrent import SolverPool
    with pytest.raises(RuntimeError):
        pool = SolverPool(workers=2, device="cuda")
        pool.result(pool.submit(lambda: 1))
    with pytest.raises(RuntimeError):
        SolverPool(workers=1, device="cpu")


def test_pipeline_helpers_match_the_reference_on_a_real_layer():
    """Only where /root/reference exists (the build container): the group structure and layer lookup of
    pipeline.py equal the reference's model_utils on a real (random-init, tiny) Qwen3."""
    import sys
    ref_src = "/root/reference/src/TruncGPTQ"
    if not os.path.isfile(os.path.join(ref_src, "model_utils.py")):
        pytest.skip("reference sources not present")
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    sys.path.insert(0, ref_src)
    try:
        import model_utils as RM
    finally:
        sys.path.remove(ref_src)
    from gptq_svd_b200 import pipeline as P
    cfg = tr.Qwen3Config(vocab_size=64, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=2, num_key_value_heads=1, head_dim=32, max_position_embeddings=64)
    model = tr.Qwen3ForCausalLM(cfg)
    assert P.get_layers(model) is RM.get_layers(model)
    layer = P.get_layers(model)[0]
    assert P.get_sequenced_groups(layer) == RM.get_sequenced_groups(layer)
    ids = [torch.randint(0, 64, (1, 16)) for _ in range(3)]
    a, ka = P.capture_initial_inputs(model, ids, device="cpu", batch_size=2)
    b, kb = RM.capture_initial_inputs(model, ids, device="cpu", batch_size=2)
    assert torch.equal(a, b) and sorted(ka) == sorted(kb)
    assert RM.get_layers(model)[0] is layer            # the wrapper was removed again
