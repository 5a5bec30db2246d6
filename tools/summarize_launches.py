"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
un = h.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[un], 1e-6)
    t = tot[r[kn][:90]]
    t[0] += 1
    t[1] += v * scale
total = sum(v[1] for v in tot.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {ms:.3f} | {100 * ms / total:.2f}% |")
