"""Probe: SM partitions (CUDA green contexts) for solves that share one GPU.

Splits the SMs into two groups, makes a green context + stream for each, and checks that the library's kernels
and its cuBLAS calls run correctly on those streams (one spectral solve per partition, results against the same
solve on the default stream), then times a wide and a narrow solve side by side with and without partitions.
Usage: python scripts/green_ctx_probe.py [sms_for_partition_b]"""
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cuda.bindings import driver as cu

import gptq_svd_b200 as G
from gptq_svd_b200 import _lib


def ck(r):
    err = r[0]
    if int(err) != 0:
        raise RuntimeError(f"driver error {err}")
    return r[1] if len(r) == 2 else r[1:]


def make_partitions(dev_index, count_b):
    """-> [(sm_count, torch stream)] for the two groups: `count_b` SMs (rounded by the driver) and the rest."""
    dev = ck(cu.cuDeviceGet(dev_index))
    res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    groups, n, remaining = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, count_b))
    out = []
    keep = []
    for r in (groups[0], remaining):
        desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
        gctx = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        st = ck(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append((int(r.sm.smCount), torch.cuda.ExternalStream(int(st))))
        keep.append((gctx, st))
    return out, keep


def make_h(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(4 * n, n, device="cuda", generator=g, dtype=torch.float64)
    X *= torch.logspace(0, -2.5, n, device="cuda", dtype=torch.float64)
    return X.T @ X / X.shape[0]


def main():
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    count_b = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    parts, keep = make_partitions(0, count_b)
    print("partitions:", [(c, hex(s.cuda_stream)) for c, s in parts])
    lib = _lib.load()
    Hn, Hw = make_h(4096, 1), make_h(12288, 2)
    ref_n = G.spectral_solve(Hn, 1e-4, "energy")
    ref_w = G.spectral_solve(Hw, 1e-4, "energy")
    torch.cuda.synchronize()

    def solve_on(H, stream, budget, out, key):
        torch.cuda.set_device(0)
        lib.tq_set_sm_budget(int(budget))
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            f = G.spectral_solve(H, 1e-4, "energy")
            e1.record(stream)
        out[key] = (f, e0, e1)

    # correctness on partition streams
    (cb, sb), (ca, sa) = parts
    out = {}
    for H, (c, s), ref, name in ((Hn, parts[0], ref_n, "narrow on B"), (Hw, parts[1], ref_w, "wide on A")):
        s.wait_stream(torch.cuda.current_stream())
        solve_on(H, s, c, out, name)
        torch.cuda.synchronize()
        f, e0, e1 = out[name]
        print(f"{name} ({c} SMs): {e0.elapsed_time(e1):.1f} ms, k {f.k} vs {ref.k}, perm equal "
              f"{bool(torch.equal(f.perm, ref.perm))}, R rel diff {float((f.R - ref.R).abs().max() / ref.R.abs().max()):.2e}")

    # side by side: wide + 3 narrow, (a) plain streams with budgets, (b) partitions
    def side_by_side(stream_w, budget_w, streams_n, budget_n, label):
        out = {}
        for s in [stream_w] + streams_n:
            s.wait_stream(torch.cuda.current_stream())
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        ths = [threading.Thread(target=solve_on, args=(Hw, stream_w, budget_w, out, "w"))]
        ths += [threading.Thread(target=solve_on, args=(Hn, s, budget_n, out, f"n{i}")) for i, s in enumerate(streams_n)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for s in [stream_w] + streams_n:
            torch.cuda.current_stream().wait_stream(s)
        t1.record()
        torch.cuda.synchronize()
        ms = {k: v[1].elapsed_time(v[2]) for k, v in out.items()}
        print(f"{label}: total {t0.elapsed_time(t1):.1f} ms; " + ", ".join(f"{k} {v:.0f}" for k, v in sorted(ms.items())))

    plain = [torch.cuda.Stream() for _ in range(4)]
    for rep in range(2):
        side_by_side(plain[0], ca, plain[1:], cb // 3, f"plain streams, budgets {ca}/{cb // 3}")
        # three narrow solves share partition B: one stream each inside the same green context
        nb = [parts[0][1]]
        for _ in range(2):
            gctx = keep[0][0]
            nb.append(torch.cuda.ExternalStream(int(ck(cu.cuGreenCtxStreamCreate(gctx, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0)))))
        side_by_side(parts[1][1], ca, nb, cb // 3, f"green contexts {ca} + {cb} SMs")
        side_by_side(plain[0], 148, nb, cb // 3, f"wide on the whole device, narrow solves confined to {cb} SMs")
        side_by_side(plain[0], ca, nb, cb // 3, f"wide on the whole device with budget {ca}, narrow solves confined to {cb} SMs")


if __name__ == "__main__":
    main()
