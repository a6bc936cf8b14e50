"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: launch_summary.py csv [title]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print(f"{k[:100]:100s} launches {a[0]:4d}  total {a[1]:10.3f} ms  share {a[1] / tot * 100:6.2f}%")
