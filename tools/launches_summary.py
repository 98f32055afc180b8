"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
per kernel name: launches, total/avg device time, share of the whole list, DRAM bytes read+written.
usage: launches_summary.py launches.csv [images] -> markdown table on stdout, JSON of DRAM bytes per image with --json"""
import csv, json, re, sys
from collections import OrderedDict
path = sys.argv[1]
images = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
iname, imet, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
per = OrderedDict()
seen = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[iname]).replace("void ", "").replace("vo::", "")
    d = per.setdefault(name, dict(launches=set(), t=0.0, rd=0.0, wr=0.0))
    d["launches"].add(r[iid])
    v = float(r[ival].replace(",", ""))
    u = r[iunit]
    if r[imet] == "gpu__time_duration.sum":
        d["t"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    else:
        b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["rd" if "read" in r[imet] else "wr"] += b
tot = sum(d["t"] for d in per.values())
if "--json" in sys.argv:
    print(json.dumps({k: (d["rd"] + d["wr"]) / (images or 1.0) for k, d in per.items()}, indent=1))
    sys.exit(0)
print("| kernel | launches | total us | avg us | share | DRAM read MB | DRAM write MB |")
print("|---|---|---|---|---|---|---|")
for k, d in sorted(per.items(), key=lambda kv: -kv[1]["t"]):
    n = len(d["launches"])
    print(f"| `{k}` | {n} | {d['t']:.1f} | {d['t']/n:.2f} | {100*d['t']/tot:.1f} % | {d['rd']/1e6:.1f} | {d['wr']/1e6:.1f} |")
print(f"\ntotal device time in the list: {tot/1e3:.3f} ms")
