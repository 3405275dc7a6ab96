"""Dev tool: attribute the executed instructions and stall samples of an `ncu --page source --csv` dump to CUDA source
lines, through the line table of the cubin (`nvdisasm -g x.cubin`, cut to the kernel's .text section).

    python scripts/ncu_lines.py src.csv kernel.dis [min_share_pct]
"""
import csv, re, sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
col = {n: i for i, n in enumerate(rows[hdr_i])}
body = rows[hdr_i + 1:]
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
loc, order = None, []
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and ".dword" not in ln and ".byte" not in ln:
        order.append(loc)
print(f"sass lines: listing {len(order)}, profile {len(body)}")
n = min(len(order), len(body))
inst, samp = defaultdict(float), defaultdict(float)
for r, l in zip(body[:n], order[:n]):
    inst[l] += float(r[col["Instructions Executed"]] or 0)
    samp[l] += float(r[col["# Samples"]] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
byfile_i, byfile_s = defaultdict(float), defaultdict(float)
for l in inst:
    byfile_i[l[0] if l else None] += inst[l]; byfile_s[l[0] if l else None] += samp[l]
for f in sorted(byfile_i, key=lambda k: -byfile_i[k]):
    print(f"{f}: inst {byfile_i[f] / ti * 100:.1f}%  samples {byfile_s[f] / ts * 100:.1f}%")
for l in sorted(inst, key=lambda k: (k or ("", 0))):
    if inst[l] / ti * 100 >= thresh or samp[l] / ts * 100 >= thresh:
        print(f"{l}: inst {inst[l] / ti * 100:5.1f}%  samples {samp[l] / ts * 100:5.1f}%")
