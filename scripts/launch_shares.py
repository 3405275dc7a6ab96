"""Dev tool: per-kernel share of an ncu gpu__time_duration launch list (csv)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ik][:90], []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"# {sum(len(v) for v in agg.values())} launches, total {tot/1000:.1f} us (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{sum(v)/tot*100:5.1f}%  n={len(v):3d}  avg {sum(v)/len(v)/1000:8.1f} us  {k}")
