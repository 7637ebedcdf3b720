"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count,
total and share.  Usage: python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    try:
        v = float(r[vi].replace(",", ""))
    except Exception:
        continue
    u = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1)
    name = r[ki].split("(")[0]
    c, t = agg.get(name, (0, 0.0))
    agg[name] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
print(f"# {sys.argv[1]}: {sum(c for c, _ in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time (cold-cache, serialised)")
print(f"{'kernel':60s} {'launches':>9s} {'total ms':>10s} {'share':>7s}")
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name[:60]:60s} {c:9d} {t / 1e6:10.3f} {100 * t / tot:6.1f}%")
