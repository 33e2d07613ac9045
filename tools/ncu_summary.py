"""summarise an .ncu-rep: key metrics + hottest instructions (run on the build box, no GPU)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.15):
        print(f"{h:95s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c for c in r):
        hdr, start = r, i + 1
        break
si, j = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(k, c) for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[start:]:
    try:
        v = float(r[j])
    except Exception:
        continue
    st = sorted(((float(r[k] or 0), c) for k, c in stall_cols), reverse=True)[:2]
    data.append((v, r[si], st))
tot = sum(d[0] for d in data)
for v, s, st in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{100 * v / tot:5.1f}%  {s[:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
