"""Summarise an .ncu-rep (read on the CPU box): headline metrics + top stall instructions.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
print(f"# {rep}")
for k in want:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:75s} {vals[i]} {units[i]}")
for i, h in enumerate(hdr):
    if "tensor" in h and "peak_sustained" not in h and h not in want and vals[i] not in ("0", "", "n/a"):
        print(f"{h:75s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci, si = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
stall_cols = [(j, n) for j, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data, tot_by = [], {n: 0.0 for _, n in stall_cols}
for r in rows[2:]:
    try:
        s = float(r[si])
    except Exception:
        continue
    data.append((s, r[ci].strip(), r))
    for j, n in stall_cols:
        try:
            tot_by[n] += float(r[j])
        except Exception:
            pass
tot = sum(d[0] for d in data) or 1.0
print("\n## stall reasons (share of all warp-stall samples)")
for n, v in sorted(tot_by.items(), key=lambda x: -x[1])[:8]:
    print(f"{n:28s} {100 * v / tot:5.1f}%")
print("\n## top instructions by stall samples")
for s, ins, r in sorted(data, key=lambda x: -x[0])[:18]:
    top = max(stall_cols, key=lambda jn: float(r[jn[0]] or 0))
    print(f"{100 * s / tot:5.1f}%  {ins[:70]:70s} {top[1]}")
