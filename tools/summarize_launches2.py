"""Per-kernel mean duration of an `ncu --metrics gpu__time_duration.sum --csv` launch list, keyed by kernel name + grid."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) < 15:
        continue
    name = r[4].split("(")[0][-60:] + " grid=" + r[8]
    agg.setdefault(name, []).append(float(r[-1]) / 1000)
for k, v in agg.items():
    print(f"{k:90s} n={len(v):3d} mean={sum(v) / len(v):8.2f} us")
