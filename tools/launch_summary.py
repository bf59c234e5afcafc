"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X ...): per kernel total time, launches, share.
usage: python tools/launch_summary.py launches.csv [passes]   (passes = forward passes captured, to print per-pass numbers)"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
passes = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    tot[r[ik]] += float(r[iv].replace(",", "")) * 1e-6   # ns -> ms
    cnt[r[ik]] += 1
total = sum(tot.values())
print(f"# per forward pass ({passes:g} passes captured); cold-cache, serialised launches: compare SHARES, not absolute times")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v / passes:9.3f} ms {cnt[k] / passes:7.1f}x {100 * v / total:5.1f}%  {k[:150]}")
print(f"total {total / passes:.3f} ms per forward pass")
