#!/usr/bin/env python
"""
bench.py - TruncGPTQ solve-and-quantize hot path on B200 (driver contract in the task
statement; metric / config from BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
  python bench.py --impl reference --steps K --warmup W    # CPU reference arm (oracle port)

Workload (configs[1] of BASELINE.json): Qwen3-8B-shaped decoder block, random-init weights,
4-bit symmetric g128, eps 1e-4 'energy', 128 x 2048 synthetic calibration tokens fed in 4
batches of 32 x 2048 like the reference run (run_benchmark.py:37), block_size 1024.
One STEP = one decoder layer through the whole hot path:
  4 groups x [H += X^T X over 262144 tokens  ->  process_hessian_alt]  +  7 Linears x gptq_fwrd
(q/k/v share one H, gate/up share one H: quantize.py:110-219, model_utils.py:77-108).
All 36 layers of the model have the same shapes, so the headline
  value = "Qwen3-8B end-to-end quantize time (s)" = ceil(36 / n_gpus) x mean step time
(with N GPUs the independent layers are sharded across ranks, no data-path collective: weak scaling; a rank
quantises whole layers, hence the ceiling).  `--steps 36` times a whole model per rank.

With N > 1 the same JSON line carries a second block, "block_parallel" (strong scaling): ONE decoder block
through gptq_svd_b200.dist.quantize_block_parallel - calibration tokens sharded T / N per rank, one NCCL
all-reduce of each fp64 H, the four solves placed longest-first on different GPUs, factors handed to the GPUs
that own sibling Linears, loops on their owners.

Scheduling inside a step (one GPU): the four Hessians of a block are independent once accumulated, so
the wide (n = 12288) solve starts first with the whole GPU and, when its band reduction - the part that
wants every SM - is done, drops to an SM budget while the three narrow solves run next to its
bulge chase and tail (gptq_svd_b200/concurrent.py; `--overlap-tail 0` / `--concurrent-solves 0` switch this off).
stdout carries exactly one JSON line; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS = 36
TOKENS = 128 * 2048
CHUNK = 32 * 2048
# (in_features, [out_features of every Linear sharing this H])
GROUPS = [(4096, [4096, 1024, 1024]), (4096, [4096]), (4096, [12288, 12288]), (12288, [4096])]
SMALL_GROUPS = [(512, [512, 128, 128]), (512, [512]), (512, [1536, 1536]), (1536, [512])]   # --tiny (CI)
METRIC = "qwen3_8b_truncgptq_quantize_time"
UNIT = "s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = {}
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
    which = "measured" if d else "fallback (B200_PROFILING.md)"
    return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "tensor_tflops": float(d.get("bf16_tflops_sustained", 1500.0)),
            "source": which}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """Samples from here on count (the process is started during the warm-up: nvidia-smi's own start-up - NVML
        initialisation, device enumeration - disturbs the driver for a few hundred ms and used to land in the first
        timed step)."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("TQ_BENCH_SMI_MS", "500"), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baseline
def _cpu_threads():
    """Every host core for the BLAS behind numpy / scipy (OpenBLAS, pthreads): torchrun exports OMP_NUM_THREADS=1,
    which OpenBLAS would honour - must run BEFORE numpy is imported.  Returns the thread count in effect."""
    cores = os.cpu_count() or 1
    for var in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    try:
        import numpy  # noqa: F401
        import scipy.linalg  # noqa: F401
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(cores)
        used = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        vend = sorted({f"{i.get('internal_api')} {i.get('version')}" for i in threadpool_info() if i.get("user_api") == "blas"})
        return (min(used) if used else cores), ", ".join(vend)
    except Exception:
        return cores, "unknown"


def _lapack_qrcp(S):
    import numpy as np
    import scipy.linalg as sla
    r, p = sla.qr(S, mode="r", pivoting=True)            # LAPACK dgeqp3, what the reference's jax call runs on a CPU
    return r[:S.shape[0]], p.astype(np.int64)


def cpu_sample(seed: int = 0, keep: bool = False):
    """Bounded sample (10 - 30 s) of the same workload through the CPU oracle (numpy / LAPACK port of the
    reference; /root/reference itself is Python and does not travel to the GPU box).
    One n=4096 group: H from 8192 tokens, full solver, loop on 256 rows of a 4096-wide Linear with the
    arithmetic the reference itself uses on a CPU (use_triton=False: torch loop, gptq_utils.py:516-534).
    Returns per-stage seconds and the extrapolation to one model; keep=True also returns the inputs and the
    oracle's results, which bench.py feeds to the CUDA path for the "parity" block."""
    import numpy as np
    from oracle import truncgptq_oracle as O

    n, Ts, ms = 4096, 8192, 256
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((n, n)).astype(np.float32) * np.logspace(0, -1, n, dtype=np.float32)[None, :]
    X = (rng.standard_normal((Ts, n)).astype(np.float32) @ A.T / np.sqrt(n) * 3).astype(np.float16)
    W = O.make_weight(ms, n, seed + 1)
    t0 = time.perf_counter()
    acc = O.HessianAccumulator(n)
    acc.add_batch(X)
    H = acc.get_hessian()
    t1 = time.perf_counter()
    f = O.process_hessian_alt(H, 1e-4, "energy", qrcp=_lapack_qrcp)
    t2 = time.perf_counter()
    q = O.Quantizer(4, 128, True)
    fw, k, codes = O.gptq_fwrd(W, f.R, q, f.perm, block_size=1024, use_triton=False, return_codes=True)
    err = O.quantization_error(W, fw, f.R_x, f.perm)
    t3 = time.perf_counter()
    tH, tS, tL = t1 - t0, t2 - t1, t3 - t2
    per_layer = tH * (TOKENS / Ts) * (3 + 9) + tS * (3 + 27) + tL * (71680 / ms)
    out = {"t_hessian": tH, "t_solver": tS, "t_loop": tL, "k": int(k), "model_s": per_layer * LAYERS,
           "sample": (f"oracle port on one n=4096 group: H from {Ts} tokens, full eigh+dgeqp3+qr, loop on {ms} rows; "
                      "extrapolated by T*n^2 (H), n^3 (solver), m*n^2 (loop) to 36 Qwen3-8B layers")}
    if keep:
        out["data"] = {"X": X, "W": W, "H": H, "f": f, "codes": codes, "err": err, "q": q}
    return out


def cpu_full_sample(log=None):
    """The reference arm proper: ONE timed instance of every distinct shape of a
    Qwen3-8B decoder layer through the CPU oracle, multiplied by multiplicity only:
      H        n = 4096 over all 262144 tokens (4 batches of 65536, as the reference feeds them);
               n = 12288 over ONE batch of 65536 tokens, x 4 (the GEMM is linear in the token count);
      solver   process_hessian_alt at n = 4096 and at n = 12288, complete (eigh + dgeqp3 + qr);
      loop     gptq_fwrd at full size for (m, n) = (4096, 4096), (1024, 4096), (12288, 4096), (4096, 12288)
               with the reference's CPU arithmetic (use_triton=False) and its error metric.
    per layer = 3 H4096 + H12288 + 3 S4096 + S12288 + 2 L(4096,4096) + 2 L(1024,4096) + 2 L(12288,4096) + L(4096,12288)."""
    import numpy as np
    from oracle import truncgptq_oracle as O

    def say(msg):
        if log:
            log(msg)

    rng = np.random.RandomState(0)
    t = {}

    def make_x(rows, n):
        A = rng.standard_normal((n, n)).astype(np.float32) * np.logspace(0, -1, n, dtype=np.float32)[None, :]
        return (rng.standard_normal((rows, n)).astype(np.float32) @ A.T / np.sqrt(n) * 3).astype(np.float16)

    facs = {}
    for n, batches in ((4096, 4), (12288, 1)):
        X = make_x(CHUNK, n)
        acc = O.HessianAccumulator(n)
        t0 = time.perf_counter()
        for _ in range(batches):
            acc.add_batch(X)
        t[f"H{n}"] = (time.perf_counter() - t0) * (4 / batches)
        say(f"reference: H n={n}: {t[f'H{n}']:.1f} s (timed {batches} of 4 batches)")
        H = acc.get_hessian()
        del acc, X
        t0 = time.perf_counter()
        facs[n] = O.process_hessian_alt(H, 1e-4, "energy", qrcp=_lapack_qrcp)
        t[f"S{n}"] = time.perf_counter() - t0
        say(f"reference: solver n={n}: {t[f'S{n}']:.1f} s (k = {facs[n].k})")
        del H
    for m, n in ((4096, 4096), (1024, 4096), (12288, 4096), (4096, 12288)):
        W = O.make_weight(m, n, 7)
        f = facs[n]
        q = O.Quantizer(4, 128, True)
        t0 = time.perf_counter()
        fw, _ = O.gptq_fwrd(W, f.R, q, f.perm, block_size=1024, use_triton=False)
        O.quantization_error(W, fw, f.R_x, f.perm)
        t[f"L{m}x{n}"] = time.perf_counter() - t0
        say(f"reference: loop {m} x {n}: {t[f'L{m}x{n}']:.1f} s")
    per_layer = (3 * t["H4096"] + t["H12288"] + 3 * t["S4096"] + t["S12288"] + 2 * t["L4096x4096"] +
                 2 * t["L1024x4096"] + 2 * t["L12288x4096"] + t["L4096x12288"])
    return {"model_s": per_layer * LAYERS, "layer_s": per_layer, "stage_seconds": t,
            "ranks": {str(n): int(facs[n].k) for n in facs},
            "sample": ("oracle port, one timed instance of every distinct shape of a Qwen3-8B layer: H at n=4096 over all "
                       "262144 tokens and at n=12288 over 65536 tokens (x4); process_hessian_alt (eigh+dgeqp3+qr) at "
                       "n=4096 and n=12288 in full; gptq_fwrd (use_triton=False) + error metric at full size for "
                       "(4096,4096), (1024,4096), (12288,4096), (4096,12288); multiplied by multiplicity x 36 layers")}


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, blas = _cpu_threads()
    t0 = time.perf_counter()
    if args.reference_sample == "bounded":
        s = cpu_sample(0)
        steps = 1
    else:
        # The full pass takes ~10 minutes of host time and does not depend on --gpus: when the driver asks for the
        # reference arm again on the same box (N = 2, 4, 8 of the scaling run), the measurement of the first call is
        # reused and the line says so.  --reference-fresh measures again.
        import socket
        cache_path = os.path.join(ROOT, ".reference_arm_cache.json")
        key = {"host": socket.gethostname(), "cores": cores, "blas": blas, "bits": 4, "sym": True, "eps": 1e-4}
        s = None
        if not args.reference_fresh and os.path.exists(cache_path):
            try:
                c = json.load(open(cache_path))
                if c.get("key") == key and 0 <= time.time() - c.get("measured_at", 0) < 12 * 3600:
                    s = c["sample"]
                    s["sample"] += (f"; measured once on this box {int(time.time() - c['measured_at'])} s ago in "
                                    f"{c['wall_s']:.0f} s and reused (the figure does not depend on --gpus)")
                    cached_wall = c["wall_s"]
            except Exception:
                s = None
        if s is None:
            s = cpu_full_sample(lambda m: sys.stderr.write(m + "\n"))
            try:
                json.dump({"key": key, "measured_at": time.time(), "wall_s": time.perf_counter() - t0, "sample": s},
                          open(cache_path, "w"))
            except Exception:
                pass
            cached_wall = None
        steps = 1
    wall = time.perf_counter() - t0
    if args.reference_sample != "bounded" and cached_wall is not None:
        wall = cached_wall
    v = s["model_s"]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": 0, "ms_per_step": wall / steps * 1e3, "higher_is_better": False,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": _config(args, None),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": s["sample"], "blas": blas,
                            "stage_seconds": s.get("stage_seconds"), "layer_s": s.get("layer_s"),
                            "note": "one step = one pass over the sample described; steps / warmup requested "
                                    f"({args.steps} / {args.warmup}) are clamped to 1 / 0: a pass takes minutes"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()


def _config(args, k_list):
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    return {"workload": "Qwen3-8B-shaped random-init, 4-bit sym g128, 128x2048 synthetic tokens (BASELINE configs[1])"
            if not args.tiny else "tiny CI shapes",
            "step": "one decoder layer: 4 x (SYRK over 262144 tokens + spectral solve) + 7 x gptq_fwrd",
            "layers_per_model": LAYERS, "value_is": "ceil(36 / n_gpus) x mean step seconds",
            "bits": args.bits, "sym": bool(args.sym), "group_size": 128, "eps": args.eps, "threshold_method": "energy",
            "block_size": 1024, "activations": f"randn @ A^T, column scales logspace(0,{args.decay}), 8 outlier channels x30",
            "retained_rank": k_list, "l2": "inputs per step (12.9 GB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"layers sharded over {world} rank(s), no collective",
            "tridiagonal_reduction": {-1: "two-stage (band + bulge chase) for n >= 8192, one-stage below",
                                      0: "one-stage", 1: "two-stage (band + bulge chase)"}[args.two_stage],
            "solves": ((f"n=12288 first; after its band reduction it drops to {args.tail_budgets.split(',')[0]} SMs and "
                        f"the three n=4096 solves run next to its bulge chase and tail ({args.tail_budgets.split(',')[1]} SMs each)"
                        if args.overlap_tail else
                        "the three n=4096 Hessians of a layer side by side (SM budget 49 each), n=12288 alone")
                       if args.concurrent_solves else "one after another")}


# --------------------------------------------------------------------------- GPU arm
def make_x(torch, rows, n, seed, decay):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(n, n, device="cuda", generator=g) * torch.logspace(0, decay, n, device="cuda")[None, :]
    out = torch.empty(rows, n, device="cuda", dtype=torch.float16)
    for c in range(0, rows, 8192):
        z = torch.randn(min(8192, rows - c), n, device="cuda", generator=g)
        out[c:c + 8192] = (z @ A.T / (n ** 0.5) * 3).half()
    out[:, :8] *= 30
    return out


def _claim_stdout():
    """Route everything that writes to fd 1 (NCCL's version banner, library chatter) to stderr and
    return a file object on the real stdout, so that the bench prints exactly ONE JSON line there."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def _pin_to_gpu_numa(torch, local):
    """Run this process (and so first-touch its pinned staging buffers) on the NUMA node of its GPU: with all
    ranks on node 0 the host -> device staging of 8 ranks x 13.3 GB per step was the e2e limiter in round 1."""
    try:
        try:
            prop = torch.cuda.get_device_properties(local)
            bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = vis.split(",")[local] if vis else str(local)
            q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", idx],
                               capture_output=True, text=True, timeout=20).stdout.strip().lower()
            bdf = q[-12:] if len(q) >= 12 else q              # "00000000:1b:00.0" -> "0000:1b:00.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def gpu_parity(torch, G, sample):
    """The CPU sample's inputs through the CUDA path (the CPU run doubles as the parity oracle)."""
    import numpy as np
    d = sample["data"]
    fo, W = d["f"], d["W"]
    dev = torch.device("cuda", torch.cuda.current_device())
    X = torch.from_numpy(d["X"]).to(dev)
    acc = G.HessianAccumulator(X.shape[1], dev)
    acc.add_batch(X)
    Hg = acc.get_hessian()
    relH = float(np.linalg.norm(Hg.cpu().numpy() - d["H"]) / np.linalg.norm(d["H"]))
    f_e2e = G.spectral_solve(Hg, 1e-4, "energy")
    f = G.spectral_solve(torch.from_numpy(d["H"]).to(dev), 1e-4, "energy")            # stage-wise: the oracle's H
    k = fo.k
    perm = f.perm.cpu().numpy()
    e = f.eigvals.cpu().numpy()
    sig = fo.eigvals > 1e-10 * fo.eigvals[0]
    R = f.R.cpu().numpy()
    out = {"H_rel_fro": relH, "k_equal": bool(f.k == k), "k": int(k), "k_end_to_end": int(f_e2e.k),
           "eig_max_rel": float(np.abs(e[sig] / fo.eigvals[sig] - 1).max()),
           "perm_equal": bool(f.k == k and np.array_equal(perm[:k], fo.perm[:k]))}
    if f.k == k and out["perm_equal"]:
        out["R_max_rel"] = float(np.abs(R - fo.R).max() / np.abs(fo.R).max())
        out["Rx_max_rel"] = float(np.abs(f.R_x.cpu().numpy() - fo.R_x).max() / np.abs(fo.R_x).max())
    q = G.Quantizer(4, 128, True)
    res = G.gptq_quantize(torch.from_numpy(W).to(dev), torch.from_numpy(fo.R).to(dev), q, torch.from_numpy(fo.perm).to(dev),
                          block_size=1024, use_triton=False, R_x=torch.from_numpy(fo.R_x).to(dev))
    codes = res.codes.cpu().numpy().astype(np.int32) + res.min_q
    out["codes_frac"] = float(np.mean(codes == d["codes"]))
    out["scales_equal"] = bool(np.array_equal(q.scale.cpu().numpy(), d["q"].scale) and
                               np.array_equal(q.zero.cpu().numpy(), d["q"].zero))
    out["rel_err_ratio"] = float(res.rel_error / d["err"])
    out["what"] = ("CUDA path on the CPU sample's own X / W (n=4096, 8192 tokens, 256 rows): H end to end; solver on "
                   "the oracle's fp64 H (k, eigenvalues, perm[:k], R, R_x); loop with the oracle's factors, torch-loop "
                   "arithmetic on both sides (codes, scales/zeros, ||WX-QX|| ratio)")
    return out


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--sym", type=int, default=1)
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--decay", type=float, default=-1.0)
    ap.add_argument("--tiny", action="store_true", help="small shapes (functional check only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip solver_ms / kernels / block_parallel")
    ap.add_argument("--e2e-steps", type=int, default=-1,
                    help="steps of the end-to-end leg; -1: min(layers per rank, 8) - only the first layer's copies are "
                         "exposed, as in a whole-model run; 0 skips the leg (parameter sweeps)")
    ap.add_argument("--reference-fresh", action="store_true",
                    help="reference arm: measure again even if this box has a measurement less than 12 h old")
    ap.add_argument("--reference-sample", default="full", choices=["full", "bounded"],
                    help="--impl reference: every distinct shape once (minutes) or the 20 s sample of the ours arm")
    ap.add_argument("--overlap-tail", type=int, default=1,
                    help="start the narrow solves when the wide solve has finished its band reduction")
    ap.add_argument("--tail-budgets", default="76,24", help="SM budgets in the tail: wide,narrow")
    ap.add_argument("--no-priority", action="store_true",
                    help="run the wide solve on a normal-priority stream (default: high priority, it is the critical path)")
    ap.add_argument("--narrow-start", default="band", choices=["start", "band"],
                    help="release the narrow solves when the wide solve starts, or when its band reduction is done")
    ap.add_argument("--two-stage", type=int, default=-1,
                    help="tridiagonal reduction: -1 automatic (two-stage for n >= 8192), 0 one-stage, 1 two-stage everywhere")
    ap.add_argument("--concurrent-solves", type=int, default=1,
                    help="solve the n <= 8192 Hessians of a layer side by side on one GPU (0: one after another)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, real_stdout)

    import torch
    import torch.distributed as dist

    import gptq_svd_b200 as G
    from gptq_svd_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:      # not under torchrun: re-launch ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        real_stdout.flush()
        sys.exit(subprocess.call(cmd, stdout=real_stdout))      # the ranks write their JSON line to OUR stdout
    torch.cuda.set_device(local)
    numa_node = _pin_to_gpu_numa(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    import ctypes as C
    groups = SMALL_GROUPS if args.tiny else GROUPS
    tokens, chunk = (8192, 2048) if args.tiny else (TOKENS, CHUNK)
    dev = torch.device("cuda", local)

    # ---- synthetic inputs: resident in HBM for `value`, mirrored in pinned host memory for `e2e`
    Xs, Ws = [], []
    for gi, (n, outs) in enumerate(groups):
        Xs.append(make_x(torch, tokens, n, 1000 * rank + gi, args.decay))
        Ws.append([(torch.randn(m, n, device=dev, generator=torch.Generator(device="cuda").manual_seed(2000 * rank + 10 * gi + li))
                    * 0.02).half() for li, m in enumerate(outs)])
    ranks_seen = []

    copy_stream = torch.cuda.Stream(device=dev)
    prof = {"on": False, "acc": {}}       # sampled kernel timing, summed over the host threads that launched
    prof_lock = threading.Lock()
    pool = None
    if args.concurrent_solves:
        from gptq_svd_b200.concurrent import SolverPool
        pool = SolverPool(workers=4, device=dev)
    staging = {}          # device staging buffers of the e2e leg (allocated once, reused every step)

    def prof_collect():
        """Fold the calling thread's sampled launches into prof['acc'] and restart its sampling."""
        for kind, name in enumerate(_lib.PROF_KINDS):
            w, ms_, sa, to, wa, sm = (C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0), C.c_double(0),
                                      C.c_double(0))
            lib.tq_profile_kernel(kind, C.byref(w), C.byref(ms_), C.byref(sa), C.byref(to), C.byref(wa), C.byref(sm))
            if to.value:
                with prof_lock:
                    a = prof["acc"].setdefault(name, {"work": 0.0, "ms": 0.0, "sampled": 0, "total": 0, "work_all": 0.0,
                                                      "sm_ms": 0.0})
                    a["work"] += w.value; a["ms"] += ms_.value; a["sampled"] += sa.value
                    a["total"] += to.value; a["work_all"] += wa.value; a["sm_ms"] += sm.value
        lib.tq_profile_end(None, None, None, None)

    def solve(H):
        """process_hessian_alt on the calling (worker) thread, with that thread's kernel sampling."""
        H.record_stream(torch.cuda.current_stream(dev))
        lib.tq_set_eigh_two_stage(args.two_stage)          # per host thread
        if prof["on"]:
            lib.tq_profile_begin(4)
        try:
            return G.process_hessian_alt(H, args.eps, "energy")
        finally:
            if prof["on"]:
                prof_collect()

    e2e_state = {"step": 0, "total": 0, "pending": None, "done": [None, None]}

    def enqueue_copies(x_src, w_src, order, slot):
        """All host -> device copies of one layer into staging set `slot`, on the copy stream, in the order the groups
        are consumed; an event per tensor.  The set is free once the step that used it last has finished."""
        events = {}
        if e2e_state["done"][slot] is not None:
            copy_stream.wait_event(e2e_state["done"][slot])
        with torch.cuda.stream(copy_stream):
            for gi in order:
                n, outs = groups[gi]
                for c in range(0, tokens, chunk):
                    key = (slot, "x", gi, c)
                    if key not in staging:
                        staging[key] = torch.empty((min(chunk, tokens - c), n), dtype=torch.float16, device=dev)
                    staging[key].copy_(x_src[gi][c:c + chunk], non_blocking=True)
                    events[("x", gi, c)] = torch.cuda.Event()
                    events[("x", gi, c)].record(copy_stream)
                for li, m in enumerate(outs):
                    key = (slot, "w", gi, li)
                    if key not in staging:
                        staging[key] = torch.empty((m, n), dtype=torch.float16, device=dev)
                    staging[key].copy_(w_src[gi][li], non_blocking=True)
                    events[("w", gi, li)] = torch.cuda.Event()
                    events[("w", gi, li)].record(copy_stream)
        return events

    timeline = {"wide": []} if os.environ.get("TQ_BENCH_TIMELINE") else None

    def layer_step(x_src, w_src, host: bool, sink=None):
        """One decoder layer through the public API.  host=True: inputs come from pinned host
        buffers - every H2D copy of the step is enqueued on a side stream up front and the compute
        stream waits per chunk, so the copies of later groups overlap the solves of earlier ones -
        and the dequantised fp16 weights go back to pinned host buffers (D2H)."""
        ks = []
        events = {}
        small = [gi for gi, (n, _) in enumerate(groups) if n <= 8192] if pool is not None else []
        if len(small) < 2:
            small = []
        wides = [gi for gi in range(len(groups)) if gi not in small]
        tail = bool(args.overlap_tail) and len(small) > 0 and len(wides) == 1
        order = (wides + small) if tail else (small + wides)
        tail_wide, tail_narrow = (int(v) for v in args.tail_budgets.split(","))
        budget = tail_narrow if tail else max(8, 148 // max(1, len(small)))
        if host:
            # double-buffered staging: this step's copies were enqueued one step ago (or just now for the first
            # step); the NEXT step's copies are enqueued here, behind them on the copy stream, so that the PCIe
            # transfer of layer i + 1 runs under the solves of layer i
            slot = e2e_state["step"] % 2
            if e2e_state["pending"] is None:
                e2e_state["pending"] = enqueue_copies(x_src, w_src, order, slot)
            events = e2e_state["pending"]
            e2e_state["pending"] = (enqueue_copies(x_src, w_src, order, slot ^ 1)
                                    if e2e_state["step"] + 1 < e2e_state["total"] else None)
        facs = [None] * len(groups)
        pending = {}
        looped = set()
        deferred, wide_handle = [], None
        released = threading.Semaphore(0)

        def run_loops(gi):
            looped.add(gi)
            n, outs = groups[gi]
            R, R_x, perm = facs[gi]
            ks.append(int(R.shape[0]))
            for li, m in enumerate(outs):
                if host:
                    torch.cuda.current_stream(dev).wait_event(events[("w", gi, li)])
                    W = staging[(slot, "w", gi, li)]
                else:
                    W = w_src[gi][li]
                q = G.Quantizer(args.bits, 128, bool(args.sym))
                fw, k = G.gptq_fwrd(W, R, q, perm, block_size=1024, use_triton=True, R_x=R_x)
                if host:
                    sink[gi][li].copy_(fw, non_blocking=True)

        for gi in order:
            n, outs = groups[gi]
            acc = G.HessianAccumulator(n, dev)
            for c in range(0, tokens, chunk):
                if host:
                    torch.cuda.current_stream(dev).wait_event(events[("x", gi, c)])
                    xb = staging[(slot, "x", gi, c)]
                else:
                    xb = x_src[gi][c:c + chunk]
                acc.add_batch(xb.view(-1, 2048, n) if (xb.shape[0] % 2048 == 0) else xb)
            H = acc.get_hessian()
            del acc
            if tail and gi in wides:
                # the wide solve starts first on a worker with the whole GPU; when the part of it that wants every SM
                # (band reduction; the whole reduction on the one-stage path) is done it drops to `tail_wide` SMs
                # and the narrow solves start next to its bulge chase and its latency- / DGEMM-bound tail
                # (stage callback of the C ABI)
                def solve_wide(H=H, sem=released):
                    fired = []
                    if timeline is not None:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record(torch.cuda.current_stream(dev))

                    def on_stage(stage, user, sem=sem):
                        if stage in (_lib.TQ_STAGE_BAND_DONE, _lib.TQ_STAGE_SYTRD_DONE) and not fired:
                            fired.append(stage)
                            lib.tq_set_sm_budget(tail_wide)
                            sem.release()
                    cb = _lib.STAGE_CALLBACK(on_stage)
                    lib.tq_set_stage_callback(cb, None)
                    try:
                        return solve(H)
                    finally:
                        if timeline is not None:
                            ev1 = torch.cuda.Event(enable_timing=True)
                            ev1.record(torch.cuda.current_stream(dev))
                            timeline["wide"].append((ev0, ev1))
                        lib.tq_set_stage_callback(_lib.STAGE_CALLBACK(0), None)
                        if not fired:
                            sem.release()                    # never leave the main thread waiting
                wide_handle = pool.submit(solve_wide, 148, urgent=not args.no_priority)
                if args.narrow_start == "start":
                    released.release()
            elif gi in small:
                fn = (lambda H=H: solve(H))
                if tail:
                    deferred.append((gi, fn))
                else:
                    pending[gi] = pool.submit(fn, budget)
            else:
                for gj in list(pending):                 # the wide solve wants the GPU for itself
                    facs[gj] = pool.result(pending.pop(gj))
                facs[gi] = solve(H)
            del H
        if tail:
            released.acquire()
            for gj, fn in deferred:
                pending[gj] = pool.submit(fn, budget)
        # loops of a group as soon as its factors are there (narrow groups first: they finish first)
        for gj in list(pending):
            facs[gj] = pool.result(pending.pop(gj))
            run_loops(gj)
        if tail:
            facs[wides[0]] = pool.result(wide_handle)
        for gi in range(len(groups)):
            if gi not in looped:
                run_loops(gi)
        del facs
        if host:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            e2e_state["done"][slot] = ev
            e2e_state["step"] += 1
        return ks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if args.warmup > 0 and rank == 0 and not os.environ.get("TQ_BENCH_NO_SMI"):
        sampler.start()                   # started before the warm-up (its NVML start-up takes a second or two and
                                          # stalls driver calls meanwhile), counted from the timed region on
    for wi in range(args.warmup):
        if wi == args.warmup - 1:
            # the last warm-up step already samples kernels: the first events a host thread creates are slow (the
            # driver grows its pools) and used to land in the first timed step; the samples are thrown away below
            prof["on"] = True
            lib.tq_profile_begin(4)
        ranks_seen = layer_step(Xs, Ws, False)
    if prof["on"]:
        prof["on"] = False
        lib.tq_profile_end(None, None, None, None)
        with prof_lock:
            prof["acc"].clear()
    if args.warmup == 0 and rank == 0 and not os.environ.get("TQ_BENCH_NO_SMI"):
        sampler.start()
    barrier()
    sampler.mark()
    prof["on"] = True
    lib.tq_profile_begin(4)               # main thread: SYRK and the loop kernels
    l0 = lib.tq_launch_count() + (pool.launch_count() if pool else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    step_ev[0].record()
    for si in range(args.steps):
        layer_step(Xs, Ws, False)
        step_ev[si + 1].record()
    e1.record()
    barrier()
    if rank == 0:
        sys.stderr.write("per-step ms: " + ", ".join(f"{step_ev[i].elapsed_time(step_ev[i + 1]):.1f}"
                                                     for i in range(args.steps)) + "\n")
        if timeline is not None:        # where a step goes: start -> wide solve starts -> wide solve ends -> step ends
            for i, (a, b) in enumerate(timeline["wide"][-args.steps:]):
                sys.stderr.write(f"timeline step {i}: wide solve starts at {step_ev[i].elapsed_time(a):.1f} ms, runs "
                                 f"{a.elapsed_time(b):.1f} ms, step ends {b.elapsed_time(step_ev[i + 1]):.1f} ms later\n")
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = torch.tensor([lib.tq_launch_count() + (pool.launch_count() if pool else 0) - l0], device=dev,
                            dtype=torch.float64)
    prof_collect()
    prof["on"] = False
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_per_step = float(ms.item()) / args.steps
    layers_per_rank = math.ceil(LAYERS / world)
    value = layers_per_rank * ms_per_step / 1e3
    if args.e2e_steps < 0:
        args.e2e_steps = min(layers_per_rank, 8)

    # ---- e2e: same step through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    Xh = Wh = Oh = None
    pin_err = ""
    if args.e2e_steps <= 0:
        pin_err = "skipped (--e2e-steps 0)"
    else:
      try:
        def to_pinned(t):         # straight into pinned memory: no pageable copy of the 12.9 GB of activations
            return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        Xh = [to_pinned(x) for x in Xs]
        Wh = [[to_pinned(w) for w in ws] for ws in Ws]
        Oh = [[torch.empty(w.shape, dtype=w.dtype, pin_memory=True) for w in ws] for ws in Wh]
      except Exception as ex:       # pinned allocation can fail on a small host
        pin_err = str(ex)[:200]
    pinned_ok = torch.tensor([0 if pin_err else 1], device=dev, dtype=torch.int32)
    if world > 1:                 # every rank takes the same branch: the timed region below contains barriers
        dist.all_reduce(pinned_ok, op=dist.ReduceOp.MIN)
    try:
        if int(pinned_ok.item()) == 0:
            raise RuntimeError(pin_err or "pinned host allocation failed on another rank")
        e2e_state.update(step=0, total=1, pending=None)
        layer_step(Xh, Wh, True, Oh)                   # warm-up (staging allocation)
        barrier()
        e2e_state.update(step=0, total=args.e2e_steps, pending=None)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.e2e_steps):
            layer_step(Xh, Wh, True, Oh)
        t1.record()
        barrier()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = sum(x.numel() * 2 for x in Xh) + sum(w.numel() * 2 for ws in Wh for w in ws)
        d2h = sum(w.numel() * 2 for ws in Oh for w in ws)
        e2e = {"value": layers_per_rank * float(ems.item()) / args.e2e_steps / 1e3, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "numa_node_rank0": numa_node,
               "note": "pinned host X / W -> device inside the timed region on a side stream, double-buffered: the copies of "
                       "layer i + 1 run under the solves of layer i (the first layer's are exposed); dequantised fp16 "
                       "weights read back; " + ("every rank is pinned to the NUMA node of its GPU" if numa_node is not None
                                                  else "no NUMA pinning (the box exposes no node for the GPU)")}
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}
    del Xh, Wh, Oh

    # ---- extras (untimed above): solver ms per Hessian, block-parallel run, parity of the CPU sample
    extras = {}
    peaks = _peaks()
    if not args.no_extras and not args.tiny:
        if world == 1:
            extras["solver_ms"] = solver_ms(torch, G, Xs, args)
        else:
            extras["block_parallel"] = block_parallel(torch, dist, G, Xs, Ws, groups, args, rank, world, dev, tokens, chunk)
    if pool is not None:
        pool.close()

    if rank == 0:
        kernels = {}
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        for name, a in prof["acc"].items():
            if not a["sampled"] or a["ms"] <= 0:
                continue
            unit = _lib.PROF_UNIT[_lib.PROF_KINDS.index(name)]
            rate = a["work"] / (a["ms"] / 1e3)
            est_ms_per_step = a["ms"] / a["sampled"] * a["total"] / args.steps
            kernels[name] = {"launches_per_step": a["total"] / args.steps, "sampled": a["sampled"],
                             "avg_launch_ms": a["ms"] / a["sampled"], "ms_per_step_est": est_ms_per_step,
                             ("GB/s" if unit == "B" else "TFLOP/s"): rate / (1e9 if unit == "B" else 1e12),
                             "frac_of_peak": rate / ((peaks["hbm_gbs"] * 1e9) if unit == "B" else (peaks["tensor_tflops"] * 1e12)),
                             "sm_share": min(1.0, a["sm_ms"] / a["ms"] / sm_count),
                             "bound": "hbm" if unit == "B" else "tensor"}
        roof = roofline_block(kernels, peaks)
        cpu = parity = None
        if world == 1 and not args.no_cpu_baseline and not args.tiny:
            try:
                cores, blas = _cpu_threads()
                s = cpu_sample(0, keep=True)
                cpu = {"value": s["model_s"], "unit": UNIT, "cores": cores, "kind": "port", "sample": s["sample"], "blas": blas,
                       "stage_seconds": {"hessian": s["t_hessian"], "solver": s["t_solver"], "loop": s["t_loop"]}}
                try:
                    parity = gpu_parity(torch, G, s)
                except Exception as ex:
                    parity = {"error": str(ex)[:300]}
            except Exception as ex:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"[:200]}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64 solver / f32 loop / f16 SYRK inputs", "data": "synthetic", "config": _config(args, ranks_seen),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()), "roofline": roof, "cpu_baseline": cpu,
               "kernels": kernels, "parity": parity}
        out.update(extras)
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def roofline_block(kernels, peaks):
    """The dominant own kernel of the step (largest estimated ms per step among the sampled kernels)."""
    if not kernels:
        return None
    name = max(kernels, key=lambda k: kernels[k]["ms_per_step_est"])
    kk = kernels[name]
    hbm = kk["bound"] == "hbm"
    achieved = kk["GB/s"] if hbm else kk["TFLOP/s"]
    peak = peaks["hbm_gbs"] if hbm else peaks["tensor_tflops"]
    traffic, note = None, None
    try:        # DRAM bytes of ONE captured launch (ncu --set full), committed with the profile
        with open(os.path.join(ROOT, "profiles", "r02_ncu_dominant.json")) as f:
            cap = json.load(f)
        if cap.get("kernel") == name:
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
            note = cap.get("note")
    except Exception:
        pass
    doc = {"sytrd_panel_sym_kernel": "one-stage tridiagonal reduction of the three n = 4096 Hessians of a layer (lower "
                                     "triangle of the trailing matrix x reflector per column, TMA-staged; the 67 MB triangle "
                                     "is L2-resident at this size, so the kernel is latency-bound, not DRAM-bound)",
           "sb2st_chase_kernel": "bulge chase of the two-stage reduction at n = 12288 (persistent, one sweep per CTA; the "
                                 "12.6 MB band is L2-resident: bytes are L2 traffic, the kernel is bound by its "
                                 "sweep-to-sweep dependency chain)",
           "pchol_panel_kernel": "pivoted Cholesky panel (one grid barrier per pivot; latency-bound)"}
    return {"bound": kk["bound"], "kernel": name + (": " + doc[name] if name in doc else ""),
            "achieved": achieved, "peak": peak, "peak_source": peaks["source"], "unit": "GB/s" if hbm else "TFLOP/s",
            "frac": achieved / peak, "sm_share": kk["sm_share"], "frac_within_sm_share": achieved / peak / kk["sm_share"],
            "sm_share_note": "launches of this kernel were confined to sm_share of the SMs (several solves share the GPU "
                             "under SM budgets): frac is against the whole GPU's peak, frac_within_sm_share against "
                             "that share of it",
            "traffic": traffic, "traffic_note": note,
            "algorithmic_work": "see include/truncgptq.h (tq_profile_kernel table) and DESIGN.md 3",
            "sampled_launches": kk["sampled"], "avg_launch_ms": kk["avg_launch_ms"],
            "share_of_step": kk["ms_per_step_est"], "share_unit": "estimated ms per step (kernels of concurrent solves overlap)"}


def solver_ms(torch, G, Xs, args):
    """Second half of BASELINE.json's metric: tq_spectral_solve ms per n x n Hessian (CUDA events, second call)."""
    out = {}
    for n in (4096, 8192, 12288):
        src = {4096: 0, 12288: 3}.get(n)
        X = Xs[src][:65536] if src is not None else make_x(torch, 65536, n, 77, args.decay)
        acc = G.HessianAccumulator(n, X.device)
        acc.add_batch(X)
        H = acc.get_hessian()
        del acc
        G.spectral_solve(H, args.eps, "energy")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f = G.spectral_solve(H, args.eps, "energy")
        e1.record()
        torch.cuda.synchronize()
        out[str(n)] = {"ms": e0.elapsed_time(e1), "k": int(f.k)}
        del H, f
        torch.cuda.empty_cache()
    return out


def block_parallel(torch, dist, G, Xs, Ws, groups, args, rank, world, dev, tokens, chunk):
    """ONE decoder block on all N GPUs (strong scaling; north_star's multi-GPU design): rank r accumulates the
    tokens [r T / N, (r + 1) T / N) of every group, fp64 H all-reduced over NCCL, solves placed longest-first,
    factors handed to the owners of sibling Linears, loops on their owners."""
    from gptq_svd_b200 import dist as D
    # every rank must see the SAME block: regenerate rank 0's inputs everywhere
    Xb = [make_x(torch, tokens, n, gi, args.decay) for gi, (n, _) in enumerate(groups)] if rank != 0 else Xs
    Wb = [[(torch.randn(m, n, device=dev, generator=torch.Generator(device="cuda").manual_seed(10 * gi + li)) * 0.02).half()
           for li, m in enumerate(outs)] for gi, (n, outs) in enumerate(groups)] if rank != 0 else Ws
    shards = []
    for gi, (n, _) in enumerate(groups):
        b, e = D.shard_range(tokens, world, rank, 2048)
        shards.append([Xb[gi][c:min(c + chunk, e)] for c in range(b, e, chunk)])
    plan = D.plan_block(groups, world)

    def run(timers=None):
        return D.quantize_block_parallel(shards, Wb, groups, bits=args.bits, sym=bool(args.sym), eps=args.eps,
                                         block_size=1024, timers=timers, plan=plan)
    run()
    dist.barrier(); torch.cuda.synchronize()
    reps = 2
    timers = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res = run(timers)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ar = timers.get("allreduce", [])
    ar_ms = sum(a.elapsed_time(b) for _, a, b in ar) / reps        # includes waiting for the slowest rank to arrive
    per_group = len(ar) // reps if reps else 0
    ar_best = (sum(min(ar[r * per_group + g][1].elapsed_time(ar[r * per_group + g][2]) for r in range(reps))
                   for g in range(per_group)) if per_group else 0.0)
    ar_bytes = sum(nb for nb, _, _ in ar) / reps
    wide = max(nb for nb, _, _ in ar) if ar else 0
    wide_ms = min(a.elapsed_time(b) for nb, a, b in ar if nb == wide) if ar else 0.0
    arr = torch.tensor([ar_ms, wide_ms, ar_best], device=dev, dtype=torch.float64)
    dist.all_reduce(arr, op=dist.ReduceOp.MAX)
    owned = torch.tensor([len(res)], device=dev, dtype=torch.int64)
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    busbw = (2 * (world - 1) / world) * wide / (float(arr[1]) / 1e3) / 1e9 if wide_ms > 0 else None
    return {"scaling": "strong", "unit": "s", "block_s": float(t.item()) / 1e3, "model_s_if_sequential_blocks": LAYERS * float(t.item()) / 1e3,
            "allreduce_ms_per_block": float(arr[2]), "allreduce_ms_per_block_incl_rank_skew": float(arr[0]),
            "allreduce_bytes_per_block": int(ar_bytes),
            "allreduce_widest": {"bytes": int(wide), "ms": float(arr[1]), "busbw_GBps": busbw},
            "linears_quantized": int(owned.item()), "solve_owner": plan.solve_owner, "loop_owner": plan.loop_owner,
            "tokens_per_rank": tokens // world,
            "note": "one decoder block through dist.quantize_block_parallel: token-sharded SYRK, NCCL all-reduce of fp64 H "
                    "(the only collective), LPT placement of the four solves, point-to-point hand-off of (R, R_x, perm) to "
                    "sibling-Linear owners; the block time flattens at the widest solve (n = 12288), which does not split"}


if __name__ == "__main__":
    main()
