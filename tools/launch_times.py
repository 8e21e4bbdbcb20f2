"""per-kernel average duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    d[r[ik]].append(float(r[iv].replace(',', '')))
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print("%-14s n=%3d avg %9.1f us  share %.4f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
