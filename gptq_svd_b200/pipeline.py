"""
Layer-by-layer quantization of a Hugging Face causal LM through the hot path - the caller side
of the reference (`quantize.main`, src/TruncGPTQ/quantize.py:103-252, with the helpers of
src/TruncGPTQ/model_utils.py:57-181): capture the inputs of the first decoder layer, then per
layer and per group hook the group's first Linear (one statistic per group, shared by its
Linears), re-run the layer forward over the calibration batches, solve, quantize every Linear
of the group in place, and finally propagate the layer's outputs to the next layer.

`mode` selects the front end like the reference's --mode: "eigh" (TruncGPTQ, the hot path),
"gptq" (damped Cholesky, torch-loop arithmetic) or "svd" (sketch).  Everything numerical runs
in libtruncgptq.so; the model forward is PyTorch / transformers as in the reference.
"""
from __future__ import annotations

import logging
import time
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from .frontends import Sketcher, process_hessian, process_sketch
from .gptq_utils import HessianAccumulator, Quantizer, gptq_fwrd, gptq_quantize, pack_codes, process_hessian_alt

GROUP_ORDER = (("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"), ("self_attn.o_proj",),
               ("mlp.gate_proj", "mlp.up_proj"), ("mlp.down_proj",))


def get_adaptive_eps(layer_name: str, base_eps: float) -> float:
    """eps x 0.1 for o_proj / down_proj (quantize.py:17-20)."""
    return base_eps * 0.1 if ("down_proj" in layer_name or "o_proj" in layer_name) else base_eps


def get_submodule(root: nn.Module, name: str) -> nn.Module:
    for part in name.split("."):
        root = getattr(root, part)
    return root


def get_layers(model: nn.Module) -> nn.ModuleList:
    """Decoder layers of a Llama / Qwen style model (model_utils.py:57-75)."""
    inner = getattr(model, "model", None)
    if inner is not None:
        if hasattr(inner, "layers"):
            return inner.layers
        if hasattr(inner, "decoder"):
            return inner.decoder.layers
    if hasattr(model, "layers"):
        return model.layers
    if hasattr(model, "transformer") and hasattr(model.transformer, "h"):
        return model.transformer.h
    raise ValueError("Could not find layers in model architecture")


def get_sequenced_groups(layer: nn.Module) -> List[List[str]]:
    """q/k/v | o | gate/up | down, restricted to what the layer has (model_utils.py:77-108)."""
    present = {name for name, _ in layer.named_modules()}
    groups = [[n for n in g if n in present] for g in GROUP_ORDER]
    return [g for g in groups if g]


def _to_device(v, device):
    if isinstance(v, torch.Tensor):
        return v.to(device)
    if isinstance(v, (list, tuple)):
        return type(v)(_to_device(x, device) for x in v)
    return v


class _StopForward(Exception):
    pass


def capture_initial_inputs(model: nn.Module, input_ids_list: List[torch.Tensor], device="cuda",
                           batch_size: int = 1) -> Tuple[torch.Tensor, Dict[str, Any]]:
    """Hidden states entering decoder layer 0 and the kwargs it is called with
    (model_utils.py:122-181): the first layer is wrapped, the forward is cut there."""
    layers = get_layers(model)
    ids = torch.cat(list(input_ids_list), dim=0)
    n_samples, seq_len = ids.shape
    dtype = next(model.parameters()).dtype
    inps = torch.zeros((n_samples, seq_len, model.config.hidden_size), dtype=dtype, device=device)
    state = {"filled": 0, "kwargs": None}

    class Catcher(nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, hidden, **kwargs):
            b = hidden.shape[0]
            inps[state["filled"]:state["filled"] + b] = hidden
            state["filled"] += b
            if state["kwargs"] is None:
                state["kwargs"] = kwargs
            raise _StopForward()

        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                return getattr(self.inner, name)

    layers[0] = Catcher(layers[0])
    try:
        model_device = next(model.parameters()).device
        for i in range(0, n_samples, batch_size):
            try:
                model(ids[i:i + batch_size].to(model_device))
            except _StopForward:
                pass
    finally:
        layers[0] = layers[0].inner
    return inps, state["kwargs"]


def quantize_model(model: nn.Module, input_ids_list: List[torch.Tensor], *, mode: str = "eigh", w_bits: int = 4,
                   group_size: int = 128, sym: bool = False, eps: float = 1e-4,
                   threshold_method: str = "energy", adaptive_eps: bool = False, batch_size: int = 32,
                   device="cuda", actorder: bool = False, damp_percent: float = 0.01,
                   sketch_ratio: float = 1.0, keep_packed: bool = False,
                   offload_layers: bool = False, true_sequential: bool = True) -> Dict[str, Any]:
    """Quantize every decoder Linear of `model` in place (quantize.py:103-252).

    `true_sequential=True` is the reference's order: the groups of a layer are captured and quantized one
    after another, each seeing the groups before it already quantized.  `true_sequential=False` (the usual
    GPTQ option of that name; mode "eigh" only) captures the four Hessians of a layer in ONE forward pass of
    the un-quantized layer and solves the narrow ones side by side on the GPU (`concurrent.SolverPool`).

    Returns {"layer_stats": [...], "total_time": s, "packed": {...}}; with `keep_packed` the
    integer codes / scales / zeros / packed words of every Linear are kept (new: the reference
    only stores dequantised weights, quantize.py:231)."""
    if mode not in ("eigh", "gptq", "svd"):
        raise ValueError(f"mode must be eigh, gptq or svd, got {mode!r}")
    torch.set_grad_enabled(False)
    model.config.use_cache = False
    n_samples = sum(int(t.shape[0]) for t in input_ids_list)
    inps, layer_kwargs = capture_initial_inputs(model, input_ids_list, device=device, batch_size=batch_size)
    outs = torch.zeros_like(inps)
    layers = get_layers(model)
    stats: List[Dict[str, Any]] = []
    packed: Dict[str, Any] = {}
    t_start = time.time()

    def run_layer(layer, lo):
        kwargs = {k: _to_device(v, device) for k, v in layer_kwargs.items()}
        kwargs["use_cache"] = False
        return layer(inps[lo:lo + batch_size], **kwargs)

    pool = None
    if not true_sequential:
        if mode != "eigh":
            raise ValueError("true_sequential=False is implemented for mode='eigh'")
        from .concurrent import SolverPool
        pool = SolverPool(workers=3, device=device)

    def quantize_group(li, layer, group, R, R_x, perm):
        for name in group:
            sub = get_submodule(layer, name)
            W = sub.weight.data.float()
            q = Quantizer(w_bits=w_bits, group_size=group_size, sym=sym)
            t0 = time.time()
            use_triton = mode != "gptq"                      # quantize.py:211-229
            if keep_packed:
                ql = gptq_quantize(W, R, q, perm, block_size=1024, use_triton=use_triton, R_x=R_x)
                final_W, rank = ql.final_W, ql.rank
                packed[f"layer_{li}.{name}"] = {"qweight": pack_codes(ql.codes, w_bits), "scale": ql.scale,
                                                "zero": ql.zero, "bits": w_bits, "group_size": group_size,
                                                "sym": sym}
            else:
                final_W, rank = gptq_fwrd(W, R, q, perm, block_size=1024, use_triton=use_triton, R_x=R_x)
            sub.weight.copy_(final_W)
            torch.cuda.synchronize(device)
            stats.append({"name": f"layer_{li}.{name}", "rank": rank if mode != "gptq" else "N/A",
                          "time": time.time() - t0})
            logging.info(f"   {name: <15} | Rank: {stats[-1]['rank']!s: <4} | Time: {stats[-1]['time']:.2f}s")

    for li, layer in enumerate(layers):
        layer = layer.to(device)
        if pool is not None:
            groups = get_sequenced_groups(layer)
            accs, handles = [], []
            for group in groups:
                first = get_submodule(layer, group[0])
                acc = HessianAccumulator(first.weight.shape[1], device=device)
                accs.append(acc)
                handles.append(first.register_forward_hook(
                    lambda mod, inp, out, acc=acc: acc.add_batch(inp[0].detach())))
            try:
                for lo in range(0, n_samples, batch_size):
                    run_layer(layer, lo)
            finally:
                for h in handles:
                    h.remove()
            Hs = [acc.get_hessian() for acc in accs]
            epss = [get_adaptive_eps(g[0], eps) if adaptive_eps else eps for g in groups]
            narrow = [gi for gi, H in enumerate(Hs) if H.shape[0] <= 8192]
            pending = {}
            if len(narrow) > 1:
                for gi in narrow:
                    pending[gi] = pool.submit(lambda H=Hs[gi], e=epss[gi]: process_hessian_alt(H, e, threshold_method),
                                              max(8, 148 // len(narrow)))
            facs = {gi: pool.result(h) for gi, h in pending.items()}
            for gi in range(len(groups)):
                if gi not in facs:
                    facs[gi] = process_hessian_alt(Hs[gi], epss[gi], threshold_method)
            for gi, group in enumerate(groups):
                quantize_group(li, layer, group, *facs[gi])
            del Hs, facs, accs
        for group in (get_sequenced_groups(layer) if pool is None else []):
            first = get_submodule(layer, group[0])
            in_features = first.weight.shape[1]
            cur_eps = get_adaptive_eps(group[0], eps) if adaptive_eps else eps
            if mode == "svd":
                acc = Sketcher(first, int(in_features * sketch_ratio), device=device)
                handle = first.register_forward_hook(acc.hook_fn)
            else:
                acc = HessianAccumulator(in_features, device=device)
                handle = first.register_forward_hook(lambda mod, inp, out, acc=acc: acc.add_batch(inp[0].detach()))
            try:
                for lo in range(0, n_samples, batch_size):
                    run_layer(layer, lo)
            finally:
                handle.remove()
            R_x = None
            if mode == "svd":
                R, perm = process_sketch(acc.get_scaled_sketch(), cur_eps, threshold_method)
            elif mode == "gptq":
                R, perm = process_hessian(acc.get_hessian(), actorder=actorder, damp_percent=damp_percent)
            else:
                R, R_x, perm = process_hessian_alt(acc.get_hessian(), cur_eps, threshold_method)
            del acc
            quantize_group(li, layer, group, R, R_x, perm)
            del R, R_x, perm
        for lo in range(0, n_samples, batch_size):                # propagate with the quantized layer
            out = run_layer(layer, lo)
            if isinstance(out, tuple):
                out = out[0]
            outs[lo:lo + out.shape[0]] = out
        inps, outs = outs, inps
        if offload_layers:
            layer.to("cpu")
    if pool is not None:
        pool.close()
    return {"layer_stats": stats, "total_time": time.time() - t_start, "packed": packed}
