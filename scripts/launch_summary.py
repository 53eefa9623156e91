"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_summary.py <csv> [steps]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if d.get("Metric Unit", "ns") in ("us", "usecond"):
            v *= 1e3
        a = agg.setdefault(d["Kernel Name"][:56], [0, 0.0])
        a[0] += 1
        a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:56s} {a[0]:4d} {a[1] / 1e6 / steps:9.3f} ms/step {100 * a[1] / tot:5.1f}%")
print(f"total {tot / 1e6 / steps:.3f} ms/step")
