"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): python scripts/launch_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    k = r[ik].split("(")[0]
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {len(data)} launches, {tot / 1e3:.1f} ms of kernel time (cold-cache, serialised: shares matter, not absolutes)")
print(f"{'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>6s}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:8d} {t:12.1f} {t / c:10.1f} {100 * t / tot:5.1f}%  {k[:100]}")
