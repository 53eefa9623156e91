"""Aggregate per kernel the launches of the LAST complete step of an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`bench.py --device-pass-only` (a step = everything from the last `k_paste_plan` launch on, plus the fill kernels right before it):
   python scripts/launch_step.py <csv> "<comment>" > profiles/<name>.csv"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if r[0] == "ID")
data = [dict(zip(hdr, r)) for r in rows if r[0] != "ID" and len(r) == len(hdr)]
last = max(i for i, d in enumerate(data) if d["Kernel Name"].startswith("k_paste_plan"))
start = last
while start > 0 and data[start - 1]["Kernel Name"].startswith("void at::") and last - start < 3:
    start -= 1
agg = collections.OrderedDict()
for d in data[start:]:
    v = float(d["Metric Value"].replace(",", ""))
    if d.get("Metric Unit", "ns") in ("us", "usecond"):
        v *= 1e3
    name = d["Kernel Name"].split("(")[0][:44]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v / 1e3
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("kernel,launches,us,share")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k},{a[0]},{a[1]:.1f},{a[1] / tot:.4f}")
print(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.1f},1.0")
