"""Dev tool: aggregate an `ncu --page source --csv` dump into code regions.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python scripts/ncu_regions.py src.csv [bucket]

Prints, per bucket of SASS lines, the share of executed warp instructions, of stall samples and the
split of the samples over the main stall reasons, plus the global-load wavefront proxy
(L1 tag requests) of the bucket."""
import csv, sys
from collections import Counter

path = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
body = rows[hdr_i + 1:]
reasons = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_lg", "stall_not_selected",
           "stall_selected", "stall_barrier", "stall_branch_resolving", "stall_no_inst", "stall_dispatch"]
def f(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
tot_tag = sum(f(r, "L1 Tag Requests Global") for r in body)
print(f"total inst {tot_inst:.0f}  samples {tot_samp:.0f}  L1 tag requests (global) {tot_tag:.0f}  sass lines {len(body)}")
for b0 in range(0, len(body), bucket):
    part = body[b0:b0 + bucket]
    inst = sum(f(r, "Instructions Executed") for r in part)
    samp = sum(f(r, "# Samples") for r in part)
    tag = sum(f(r, "L1 Tag Requests Global") for r in part)
    rs = {n: sum(f(r, n) for r in part) for n in reasons}
    ops = Counter()
    for r in part:
        op = r[col["Source"]].split()
        if op:
            o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            ops[o.split(".")[0]] += f(r, "Instructions Executed")
    top = " ".join(f"{k}:{v / tot_inst * 100:.1f}" for k, v in ops.most_common(5))
    rtxt = " ".join(f"{n[6:]}:{v / max(tot_samp, 1) * 100:.1f}" for n, v in sorted(rs.items(), key=lambda kv: -kv[1])[:5] if v > 0)
    print(f"[{b0:5d}-{b0 + len(part):5d}] inst {inst / tot_inst * 100:5.1f}%  samples {samp / max(tot_samp, 1) * 100:5.1f}%  tags {tag / max(tot_tag, 1) * 100:5.1f}% | {rtxt} | {top}")
