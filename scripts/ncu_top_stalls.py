"""Dev tool: SASS instructions with the most stall samples in an .ncu-rep (first kernel)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; body = rows[2:]
isamp = hdr.index("# Samples"); isrc = hdr.index("Source"); ia = hdr.index("Instructions Executed")
cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_branch_resolving", "stall_lg", "stall_mio", "stall_math", "stall_not_selected", "stall_membar", "stall_dispatch", "stall_no_inst")}
tot = sum(int(r[isamp]) for r in body)
print("totals:", {k: sum(int(r[c]) for r in body) for k, c in cols.items()}, "samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:topn]
for i in sorted(order):
    r = body[i]
    st = {k[6:]: int(r[c]) for k, c in cols.items() if int(r[c]) > 0}
    prev = body[i-1][isrc].strip()[:40] if i > 0 else ""
    print(f"{i:5d} {int(r[isamp])/tot*100:5.2f}%  {r[isrc].strip()[:60]:60s} {st}   <- {prev}")
