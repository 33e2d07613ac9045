"""per-source-line and per-opcode summary of one kernel in an .ncu-rep (run where ncu is installed, no GPU needed)
usage: ncu_lines.py report.ncu-rep [kernel-name-regex] [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ksel = ["--kernel-name", "regex:" + kname] if kname else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]
for h, v in zip(hdr, vals):
    short = h.split(".", 2)[-1] if h.startswith(("SM_", "TPC.")) else h
    if h in keys or short in keys or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.15):
        print(f"{h:100s} {v:>18s}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, line, fname = None, None, ""
lines, ops = {}, collections.defaultdict(lambda: [0, 0])
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if not hdr or len(r) < 7:
        continue
    if r[0].isdigit():
        line = (fname, int(r[0]))
        lines.setdefault(line, [r[1], 0, 0, {}])
        continue
    if r[0] == "" and line is not None:
        try:
            s, ie = int(r[si]), int(r[ii])
        except Exception:
            continue
        d = lines[line]
        d[1] += s
        d[2] += ie
        for k, c in enumerate(hdr):
            if c.startswith("stall_") and "Not Issued" not in c and k < len(r):
                try:
                    d[3][c] = d[3].get(c, 0) + int(r[k])
                except ValueError:
                    pass
        sass = r[3].strip().split()
        op = sass[1] if sass[0].startswith("@") else sass[0]
        parts = op.split(".")
        op = parts[0] + ("." + parts[1] if parts[0] in ("LDS", "STS", "LDG", "DMMA", "BAR", "RED", "REDG") and len(parts) > 1 else "")
        ops[op][0] += ie
        ops[op][1] += s
tot = sum(d[1] for d in lines.values()) or 1
allst = collections.Counter()
for d in lines.values():
    allst.update(d[3])
print("stall totals:", [(k[6:], v) for k, v in allst.most_common(8)])
for ln, d in sorted(lines.items(), key=lambda x: -x[1][1])[:top]:
    st = sorted(d[3].items(), key=lambda x: -x[1])[:2]
    print(f"{100 * d[1] / tot:5.1f}% inst {d[2] / 1e6:8.1f}M {ln[0][:14]}:{ln[1]}: {d[0].strip()[:72]:72s} {[(k[6:], v) for k, v in st]}")
ti = sum(v[0] for v in ops.values()) or 1
print("opcodes:")
for k, v in sorted(ops.items(), key=lambda x: -x[1][0])[:16]:
    print(f"  {k:12s} {v[0] / 1e6:9.1f}M {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / tot:5.1f}%")
