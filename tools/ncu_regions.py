#!/usr/bin/env python
"""Developer tool: summarise an `ncu --page source --csv` export by code region (stall samples per 250-instruction block).
Usage: ncu_regions.py <source.csv> [block]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
block = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("instructions", len(data), "code bytes", len(data) * 16, "samples", tot)
reg = collections.OrderedDict()
keys = ["# Samples", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier", "stall_branch_resolving", "Instructions Executed"]
for i, r in enumerate(data):
    d = reg.setdefault(i // block, [0.0] * len(keys))
    for j, k in enumerate(keys):
        d[j] += f(r, k)
print("start  " + " ".join(k.replace("stall_", "")[:9].rjust(9) for k in keys))
for k, d in reg.items():
    ops = [r[ix["Source"]].split()[0] if r[ix["Source"]] else "" for r in data[k * block : (k + 1) * block]]
    ops = [o if not o.startswith("@") else (r.split()[1] if len(r.split()) > 1 else "") for o, r in zip(ops, (r[ix["Source"]] for r in data[k * block : (k + 1) * block]))]
    tag = ",".join(sorted(set(o.split(".")[0] for o in ops if o.startswith(("LDTM", "UTCHMMA", "UTMALDG", "UTMASTG", "LDG", "STG", "STS", "LDL", "STL", "UBLKPF", "USETMAXREG")))))
    print(f"{k * block:6d} " + " ".join(f"{x:9.0f}" for x in d), tag)
