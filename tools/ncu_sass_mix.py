"""Opcode mix (executed warp instructions) and top stall lines from an .ncu-rep source page."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[2:], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
H = None; tot = collections.Counter(); lines = []
for r in rows:
    if r and r[0] == "Address":
        H = r; continue
    if H is None or len(r) < len(H) // 2:
        continue
    try:
        n = int(r[H.index("Instructions Executed")].replace(",", ""))
        smp = int(r[H.index("# Samples")].replace(",", "") or 0)
    except ValueError:
        continue
    src = r[H.index("Source")].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
    tot[op.split(".")[0]] += n
    lines.append((smp, n, src))
T = sum(tot.values())
print("executed warp instructions:", T)
for k, v in tot.most_common(22):
    print(f"  {k:12s} {v:12d} {100*v/T:5.1f}%")
print("top sampled instructions:")
for smp, n, src in sorted(lines, reverse=True)[:18]:
    print(f"  {smp:7d} samples  {n:10d} exec  {src[:90]}")
