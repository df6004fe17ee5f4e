"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import collections
import csv
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    a = agg[r[ik]]
    a[0] += 1
    a[1] += v * scale
tot = sum(a[1] for a in agg.values())
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:100]:100s} launches={n:6d} total_ms={ms:10.3f} share={100 * ms / tot:5.1f}%")
print(f"total_ms={tot:.3f} launches={sum(a[0] for a in agg.values())}")
