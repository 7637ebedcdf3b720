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
  value = "Qwen3-8B end-to-end quantize time (s)" = 36 x mean step time / n_gpus
(with N GPUs the independent layers are sharded across ranks, no data-path collective:
weak scaling).  `--steps 36` times a whole model per rank.

Scheduling inside a step (one GPU): the four Hessians of a block are independent once accumulated, so
the wide (n = 12288) solve starts first with the whole GPU and, when its tridiagonal reduction - the
bandwidth-bound part - is done, drops to an SM budget while the three narrow solves run next to its
tail (gptq_svd_b200/concurrent.py; `--overlap-tail 0` / `--concurrent-solves 0` switch this off).
stdout carries exactly one JSON line; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
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
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
        for r in self.rows:
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
def cpu_sample(seed: int = 0):
    """Bounded sample of the same workload through the CPU oracle (numpy / LAPACK port of the
    reference; /root/reference itself is Python and does not travel to the GPU box).
    One n=4096 group: H from 8192 tokens, full solver, loop on 256 rows of a 4096-wide Linear.
    Returns per-stage seconds and the extrapolation to one model."""
    import numpy as np
    import scipy.linalg as sla
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

    def lapack_qrcp(S):
        _, r, p = sla.qr(S, mode="economic", pivoting=True)       # LAPACK dgeqp3, as the reference's stub
        return r, p.astype(np.int64)

    f = O.process_hessian_alt(H, 1e-4, "energy", qrcp=lapack_qrcp)
    t2 = time.perf_counter()
    q = O.Quantizer(4, 128, True)
    fw, k = O.gptq_fwrd(W, f.R, q, f.perm, block_size=1024, use_triton=False)
    O.quantization_error(W, fw, f.R_x, f.perm)
    t3 = time.perf_counter()
    tH, tS, tL = t1 - t0, t2 - t1, t3 - t2
    per_layer = tH * (TOKENS / Ts) * (3 + 9) + tS * (3 + 27) + tL * (71680 / ms)
    return {"t_hessian": tH, "t_solver": tS, "t_loop": tL, "k": int(k), "model_s": per_layer * LAYERS,
            "sample": (f"oracle port on one n=4096 group: H from {Ts} tokens, full eigh+dgeqp3+qr, loop on {ms} rows; "
                       "extrapolated by T*n^2 (H), n^3 (solver), m*n^2 (loop) to 36 Qwen3-8B layers")}


def run_reference(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(0)
    vals = []
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    for s in range(steps):
        vals.append(cpu_sample(s))
    wall = time.perf_counter() - t0
    v = sum(x["model_s"] for x in vals) / len(vals)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": min(args.warmup, 1), "ms_per_step": wall / steps * 1e3, "higher_is_better": False,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": _config(args, None),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[0]["sample"]},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    real_stdout.write(json.dumps(out) + "\n")
    real_stdout.flush()


def _config(args, k_list):
    return {"workload": "Qwen3-8B-shaped random-init, 4-bit sym g128, 128x2048 synthetic tokens (BASELINE configs[1])"
            if not args.tiny else "tiny CI shapes",
            "step": "one decoder layer: 4 x (SYRK over 262144 tokens + spectral solve) + 7 x gptq_fwrd",
            "layers_per_model": LAYERS, "value_is": "36 x mean step seconds / n_gpus",
            "bits": args.bits, "sym": bool(args.sym), "group_size": 128, "eps": args.eps, "threshold_method": "energy",
            "block_size": 1024, "activations": f"randn @ A^T, column scales logspace(0,{args.decay}), 8 outlier channels x30",
            "retained_rank": k_list, "l2": "inputs per step (12.9 GB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"layers sharded over {args.gpus} rank(s), no collective",
            "tridiagonal_reduction": ("two-stage (experimental, TQ_EIGH_TWO_STAGE=1: the roofline block below still "
                                      "describes the one-stage panel kernel and has no samples)"
                                      if os.environ.get("TQ_EIGH_TWO_STAGE", "0") not in ("", "0") else "one-stage"),
            "solves": ((f"n=12288 first; after its tridiagonal reduction it drops to {args.tail_budgets.split(',')[0]} SMs and "
                        f"the three n=4096 solves run next to its tail ({args.tail_budgets.split(',')[1]} SMs each)"
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


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--sym", type=int, default=1)
    ap.add_argument("--eps", type=float, default=1e-4)
    ap.add_argument("--decay", type=float, default=-1.0)
    ap.add_argument("--tiny", action="store_true", help="small shapes (functional check only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--overlap-tail", type=int, default=1,
                    help="start the narrow solves when the wide solve has finished its tridiagonal reduction")
    ap.add_argument("--tail-budgets", default="100,16",
                    help="SM budgets in the tail: wide,narrow (measured: 48,32 1.61 s; 64,28 1.59; 100,16 1.58; 124,8 1.62)")
    ap.add_argument("--overlap-loops", type=int, default=0,
                    help="run the loops of the narrow groups while the wide solve is in flight")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
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
    prof_acc = {"on": False, "b": 0.0, "ms": 0.0, "samp": 0, "tot": 0}   # wide solves timed on worker threads
    pool = None
    if args.concurrent_solves:
        from gptq_svd_b200.concurrent import SolverPool
        pool = SolverPool(workers=4, device=dev)
    staging = {}          # device staging buffers of the e2e leg (allocated once, reused every step)

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
            cur = torch.cuda.current_stream(dev)
            copy_stream.wait_stream(cur)          # staging buffers are free once the previous step is done
            with torch.cuda.stream(copy_stream):
                for gi in order:                      # copies in the order the groups are consumed
                    n, outs = groups[gi]
                    for c in range(0, tokens, chunk):
                        key = ("x", gi, c)
                        if key not in staging:
                            staging[key] = torch.empty((min(chunk, tokens - c), n), dtype=torch.float16, device=dev)
                        staging[key].copy_(x_src[gi][c:c + chunk], non_blocking=True)
                        events[key] = torch.cuda.Event()
                        events[key].record(copy_stream)
                    for li, m in enumerate(outs):
                        key = ("w", gi, li)
                        if key not in staging:
                            staging[key] = torch.empty((m, n), dtype=torch.float16, device=dev)
                        staging[key].copy_(w_src[gi][li], non_blocking=True)
                        events[key] = torch.cuda.Event()
                        events[key].record(copy_stream)
        # 1 + 2. Hessians and spectral solves (one host thread, stream and SM budget per solve,
        #    gptq_svd_b200/concurrent.py).  Default (`tail`): the bandwidth-bound wide solve starts first with the
        #    whole GPU; the narrow, latency-bound ones are released next to its tail.  `--overlap-tail 0`: the
        #    narrow ones first, side by side, then the wide one alone.
        facs = [None] * len(groups)
        pending = {}
        looped = set()
        deferred, wide_handle = [], None

        def run_loops(gi):
            n, outs = groups[gi]
            R, R_x, perm = facs[gi]
            ks.append(int(R.shape[0]))
            for li, m in enumerate(outs):
                if host:
                    torch.cuda.current_stream(dev).wait_event(events[("w", gi, li)])
                    W = staging[("w", gi, li)]
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
                    xb = staging[("x", gi, c)]
                else:
                    xb = x_src[gi][c:c + chunk]
                acc.add_batch(xb.view(-1, 2048, n) if (xb.shape[0] % 2048 == 0) else xb)
            H = acc.get_hessian()
            del acc
            if tail and gi in wides:
                # the wide solve starts first on a worker with the whole GPU; when its tridiagonal reduction (the
                # bandwidth-bound part) is done it drops to `tail_wide` SMs and the narrow solves start next to
                # its latency- and DGEMM-bound stages (stage callback of the C ABI)
                sytrd_done = threading.Semaphore(0)

                def solve_wide(H=H, sem=sytrd_done):
                    H.record_stream(torch.cuda.current_stream(dev))

                    def on_stage(stage, user, sem=sem):
                        if stage == _lib.TQ_STAGE_SYTRD_DONE:
                            lib.tq_set_sm_budget(tail_wide)
                            sem.release()
                    cb = _lib.STAGE_CALLBACK(on_stage)
                    lib.tq_set_stage_callback(cb, None)
                    if prof_acc["on"]:                   # the sampled launch timing is per host thread
                        lib.tq_profile_begin(4)
                    try:
                        return G.process_hessian_alt(H, args.eps, "energy")
                    finally:
                        if prof_acc["on"]:
                            import ctypes as C
                            b, ms_, sa, to = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
                            lib.tq_profile_end(C.byref(b), C.byref(ms_), C.byref(sa), C.byref(to))
                            prof_acc["b"] += b.value
                            prof_acc["ms"] += ms_.value
                            prof_acc["samp"] += sa.value
                            prof_acc["tot"] += to.value
                        lib.tq_set_stage_callback(_lib.STAGE_CALLBACK(0), None)
                        sem.release()                    # never leave the main thread waiting
                wide_handle = pool.submit(solve_wide, 148)
            elif gi in small:
                def solve(H=H):
                    H.record_stream(torch.cuda.current_stream(dev))     # read on the worker's stream
                    return G.process_hessian_alt(H, args.eps, "energy")
                if tail:
                    deferred.append((gi, solve))
                else:
                    pending[gi] = pool.submit(solve, budget)
            elif pending and args.overlap_loops:
                # the wide solve goes to a worker with the whole GPU as its budget; this thread meanwhile runs
                # the loops of the narrow groups (ordinary kernels that fit between the solver's launches)
                for gj in list(pending):
                    facs[gj] = pool.result(pending.pop(gj))
                def solve_wide(H=H):
                    H.record_stream(torch.cuda.current_stream(dev))
                    return G.process_hessian_alt(H, args.eps, "energy")
                wide = pool.submit(solve_wide, 148)
                for gj in small:
                    run_loops(gj)
                    looped.add(gj)
                facs[gi] = pool.result(wide)
            else:
                for gj in list(pending):                 # the wide solve wants the GPU for itself
                    facs[gj] = pool.result(pending.pop(gj))
                facs[gi] = G.process_hessian_alt(H, args.eps, "energy")
            del H
        if tail:
            sytrd_done.acquire()
            for gj, fn in deferred:
                pending[gj] = pool.submit(fn, budget)
            facs[wides[0]] = pool.result(wide_handle)
        for gj in list(pending):
            facs[gj] = pool.result(pending.pop(gj))
        # 3. grid + loop of the seven Linears
        for gi in range(len(groups)):
            if gi not in looped:
                run_loops(gi)
        del facs
        return ks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ranks_seen = layer_step(Xs, Ws, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("TQ_BENCH_NO_SMI"):
        sampler.start()
    if not os.environ.get("TQ_BENCH_NO_PROF"):
        lib.tq_profile_begin(4)
        prof_acc["on"] = True
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
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = torch.tensor([lib.tq_launch_count() + (pool.launch_count() if pool else 0) - l0], device=dev,
                            dtype=torch.float64)
    import ctypes as C
    pb, pms, psamp, ptot = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
    lib.tq_profile_end(C.byref(pb), C.byref(pms), C.byref(psamp), C.byref(ptot))
    prof_acc["on"] = False
    pb.value += prof_acc["b"]
    pms.value += prof_acc["ms"]
    psamp.value += prof_acc["samp"]
    ptot.value += prof_acc["tot"]
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_per_step = float(ms.item()) / args.steps
    value = LAYERS * ms_per_step / 1e3 / world

    # ---- e2e: same step through the public API with HOST buffers (pinned), copies inside the timed region
    e2e = None
    Xh = Wh = Oh = None
    pin_err = ""
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
        layer_step(Xh, Wh, True, Oh)
        barrier()
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
        e2e = {"value": LAYERS * float(ems.item()) / args.e2e_steps / 1e3 / world, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "note": "pinned host X / W -> device inside the timed region (copies of later groups overlap the solves of earlier ones on a side stream), dequantised fp16 weights read back"}
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(ex)[:200]}
    del Xh, Wh, Oh

    if rank == 0:
        peak, which = _peaks()
        achieved = (pb.value / 1e9) / (pms.value / 1e3) if pms.value > 0 else None
        traffic, traffic_note = None, None
        try:        # DRAM bytes of ONE captured launch (ncu --set full), committed with the profile
            with open(os.path.join(ROOT, "profiles", "r01_ncu_sytrd_sym.json")) as f:
                cap = json.load(f)
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
            traffic_note = (f"ncu capture of panel {cap['panel']} at n={cap['n']}: {traffic / 1e9:.1f} GB DRAM for "
                            f"{cap['alg_bytes'] / 1e9:.1f} GB algorithmic in that launch")
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "sytrd_panel_sym_kernel (tridiagonal reduction: lower triangle of the trailing "
                                          "matrix x reflector per column, TMA-staged, one cooperative launch per 64 columns)",
                "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes": "sum over the panel's columns of len * (len / 2 + 2 i) * 8 (DESIGN.md 3.2)",
                "sampled_launches": int(psamp.value), "total_launches": int(ptot.value),
                "avg_launch_ms": (pms.value / psamp.value) if psamp.value else None,
                "avg_alg_bytes": (pb.value / psamp.value) if psamp.value else None}
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not args.tiny:
            try:
                s = cpu_sample(0)
                cpu = {"value": s["model_s"], "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": s["sample"],
                       "stage_seconds": {"hessian": s["t_hessian"], "solver": s["t_solver"], "loop": s["t_loop"]}}
            except Exception as ex:
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"[:200]}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64 solver / f32 loop / f16 SYRK inputs", "data": "synthetic", "config": _config(args, ranks_seen),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()), "roofline": roof, "cpu_baseline": cpu}
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    if pool is not None:
        pool.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
