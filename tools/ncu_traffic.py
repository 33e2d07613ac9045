"""
profiles/r02_traffic.json from .ncu-rep captures: dram__bytes_read.sum + dram__bytes_write.sum PER LAUNCH of the kernels
bench.py reports a roofline for (bench.py reads this file instead of carrying literals).
usage: python tools/ncu_traffic.py key=report.ncu-rep[:kernel-regex] ...   (run where ncu is installed; no GPU needed)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    rep, _, kre = rest.partition(":")
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else [])
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    hdr, units = rows[0], rows[1]
    vals = rows[2:]
    tot = []
    for v in vals:
        d = dict(zip(hdr, v))
        u = dict(zip(hdr, units))
        b = sum(float(d[k]) * UNIT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tot.append(b)
    avg = sum(tot) / len(tot)
    out[key] = None if avg != avg else avg  # NaN: a replay pass of the capture failed
    out.setdefault("_source", {})[key] = f"{os.path.basename(rep)} ({len(tot)} launch(es) of {kre or 'all kernels'})"
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
