"""Dev tool: aggregate the SASS source page of an .ncu-rep into regions between labelled markers.
Prints instructions executed and stall samples for consecutive address ranges of a given size."""
import csv, subprocess, sys
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
ilsb = hdr.index("stall_long_sb"); iavg = hdr.index("Avg. Threads Executed")
body = rows[2:]
tot_i = sum(int(r[ia]) for r in body); tot_s = sum(int(r[isamp]) for r in body)
print(f"total inst {tot_i}, samples {tot_s}, sass lines {len(body)}")
for c0 in range(0, len(body), chunk):
    blk = body[c0:c0 + chunk]
    ni = sum(int(r[ia]) for r in blk); ns = sum(int(r[isamp]) for r in blk); nl = sum(int(r[ilsb]) for r in blk)
    ops = {}
    for r in blk:
        op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ia])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print(f"[{c0:5d}-{c0+len(blk):5d}] inst {ni/tot_i*100:5.1f}%  samples {ns/tot_s*100:5.1f}%  long_sb {nl/max(tot_s,1)*100:5.1f}%  " + " ".join(f"{k}:{v/tot_i*100:.1f}" for k, v in top))
