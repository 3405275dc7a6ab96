"""Dev tool: per-kernel totals of an `ncu --csv` launch list (gpu__time_duration.sum + dram bytes)."""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(sys.argv[1])) if r and r[0] == "ID")
col = {n: i for i, n in enumerate(hdr)}
acc = defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    name = r[col["Kernel Name"]].split("(")[0][:70]
    metric, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
    a = acc[name]
    if metric == "gpu__time_duration.sum":
        a[0] += 1
        a[1] += val * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    elif metric.startswith("dram__bytes"):
        a[2] += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
tot = sum(a[1] for a in acc.values())
print(f"{'kernel':70s} {'n':>4s} {'us':>9s} {'share':>6s} {'MB':>8s} {'GB/s':>7s}")
for name, (n, us, by) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} {n:4d} {us:9.1f} {us / tot * 100:5.1f}% {by / 1e6:8.1f} {by / max(us, 1e-9) / 1e3:7.0f}")
print(f"{'total':70s} {sum(a[0] for a in acc.values()):4d} {tot:9.1f}")
