"""Opcode mix + hottest SASS lines of one kernel from an ncu report with source info.
usage: ncu_opmix.py report.ncu-rep kernel_regex [n_top_lines]"""
import csv, subprocess, sys
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may match: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append((r[1], cur))
    elif cur is not None:
        cur.append(r)
name, blk = blocks[int(sys.argv[4]) if len(sys.argv) > 4 else 0]
hdr, body = blk[0], [r for r in blk[1:] if len(r) == len(blk[0])]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in body); ts = sum(int(r[isamp]) for r in body)
print(name[:100]); print("total warp-inst", tot, "sass lines", len(body), "samples", ts, "kernels matched", len(blocks))
c, s = Counter(), Counter()
for r in body:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[ia]); s[op] += int(r[isamp])
for op, n in c.most_common(28):
    print(f"  {op:10s} {100*n/tot:5.1f}% inst   {100*s[op]/max(ts,1):5.1f}% samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
top = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:ntop]
for i in sorted(top):
    r = body[i]
    st = {hdr[k][6:]: int(r[k]) for k in stall_cols if r[k] and int(r[k]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"  [{i:4d}] samp {int(r[isamp]):5d} exec {int(r[ia]):9d}  {r[isrc][:64]:64s} {st}")
