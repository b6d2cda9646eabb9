"""Top stall-sample instructions of one kernel from an ncu report's source page (SASS view).

    python tools/ncu_hot_sass.py report.ncu-rep kernel_name [top]
"""
import csv
import io
import subprocess
import sys

path, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for s in start[:1]:
    hdr = rows[s]
    ix = {h: i for i, h in enumerate(hdr)}
    si = ix["Warp Stall Sampling (All Samples)"]
    body = []
    for r in rows[s + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        try:
            body.append((float(r[si]), r))
        except ValueError:
            pass
    tot = sum(v for v, _ in body) or 1.0
    print("total samples", tot, "instructions", len(body))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for n, (v, r) in enumerate(body):
        r.append(n)
    for v, r in sorted(body, key=lambda t: -t[0])[:top]:
        reasons = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{v:7.0f} {100 * v / tot:5.1f}%  #{r[-1]:4d} {r[ix['Source']][:90]:90s} {reasons}")
